"""Steady-state solves on the device vs the reference's sparse direct solver.

tests/golden/stationary.npz holds every ``spsolve(lhs, rhs)`` the unmodified reference issued while
running its own known-answer tests for ``LPSystem.solve`` / ``FESystem.stationary_solution``
(oracle/make_golden_stationary.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "stationary.npz")


def _records():
    z = np.load(GOLDEN)
    for k in range(int(z["n"])):
        m = sp.csr_matrix((z[f"{k}_data"], z[f"{k}_indices"], z[f"{k}_indptr"]), shape=tuple(z[f"{k}_shape"]))
        yield str(z[f"{k}_label"]), m, z[f"{k}_rhs"], z[f"{k}_x"]


def test_device_pcg_matches_reference_spsolve():
    from thermca_b200.stationary import spd_solve
    seen = 0
    for label, m, b, x_ref in _records():
        x, info = spd_solve(m, b, rtol=1e-13, return_info=True)
        scale = max(np.abs(x_ref).max(), 1e-300)
        err = np.abs(x - x_ref).max() / scale
        assert err < 1e-9, (label, m.shape, err, info)
        assert info["relres"] < 1e-11
        seen += 1
    assert seen == 14


def test_zero_rhs_and_bad_matrix():
    from thermca_b200.stationary import spd_solve
    m = sp.diags([2.0, 3.0, 4.0], format="csr")
    assert np.array_equal(spd_solve(m, np.zeros(3)), np.zeros(3))
    with pytest.raises(ValueError):
        spd_solve(sp.diags([2.0, 0.0, 4.0], format="csr"), np.ones(3))
    # a non-symmetric matrix (CG diverges) must not return silently
    from thermca_b200.stationary import NotConvergedError
    bad = sp.csr_matrix(np.eye(30) + 5.0 * np.diag(np.ones(29), 1))
    with pytest.raises(NotConvergedError):
        spd_solve(bad, np.ones(30))


def test_large_poisson_like_system():
    """10^5 unknowns: 3-D 7-point Laplacian + film diagonal; residual-checked against the operator."""
    from thermca_b200.stationary import spd_solve
    n1 = 46
    e = np.ones(n1)
    T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1])
    I = sp.identity(n1)
    K = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()
    K = K + sp.diags(np.full(K.shape[0], 1e-3))
    rng = np.random.default_rng(0)
    b = rng.standard_normal(K.shape[0])
    x, info = spd_solve(K, b, rtol=1e-10, return_info=True)
    assert np.linalg.norm(K @ x - b) / np.linalg.norm(b) < 1e-9
    assert info["iterations"] < 2000


def test_multi_rhs_pcg_matches_single_solves():
    """csrc/pcg_multi.cu: K right-hand sides with one matrix sweep per iteration == K separate solves."""
    import torch
    from thermca_b200.linsolve import DeviceOperator
    n1 = 24
    e = np.ones(n1)
    T = sp.diags([-e[:-1], 2.2 * e, -e[:-1]], [-1, 0, 1])
    I = sp.identity(n1)
    K3 = (sp.kron(sp.kron(T, I), I) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(I, I), T)).tocsr()
    rng = np.random.default_rng(5)
    dev = torch.device("cuda")
    op = DeviceOperator(K3, dev)
    for K in (1, 2, 5, 8):
        B = rng.standard_normal((K, K3.shape[0]))
        B[K - 1] *= 1e-3                                       # different scales converge at different iterations
        if K > 2:
            B[1] = 0.0                                         # a zero right-hand side stays zero
        Bd = torch.from_numpy(B).to(dev)
        X = op.solve_multi(Bd, 1e-11).cpu().numpy()
        for k in range(K):
            xs = op.solve(Bd[k], 1e-11).cpu().numpy()
            nb = np.linalg.norm(B[k])
            assert np.linalg.norm(K3 @ X[k] - B[k]) <= 1e-10 * max(nb, 1e-300) + (0 if nb else 0)
            assert np.abs(X[k] - xs).max() <= 1e-9 * max(np.abs(xs).max(), 1e-300) + (0.0 if nb else 1e-300)
