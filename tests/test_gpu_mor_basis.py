"""MOR basis generation on the device vs the reference's LU-based
``thermca.fem.mor_io.transform`` (fixtures: tests/golden/mor_basis.npz, oracle/make_golden_morbasis.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "mor_basis.npz")


def _case(name):
    z = np.load(GOLDEN)
    n = len(z[name + "_M_is"])
    csr = lambda k: sp.csr_matrix((z[f"{name}_{k}_data"], z[f"{name}_{k}_indices"], z[f"{name}_{k}_indptr"]), shape=(n, n))
    args = (sp.diags(z[name + "_M_is"], format="csr"), csr("A_body"),
            [sp.diags(d, format="csr") for d in z[name + "_A_film"]], csr("A"), z[name + "_B"], z[name + "_C"],
            int(z[name + "_dim"]))
    ref = {k: z[f"{name}_{k}"] for k in ("Ar_body", "Ar_film", "Br_T", "Crs", "VT1")}
    if name + "_VT0" in z:
        ref["VT0"] = z[name + "_VT0"]
    return args, ref


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("name", ["plate_mor", "plate_mor_var"])
def test_reduced_operators_match_reference(name):
    from thermca_b200.mor_basis import transform_device
    args, ref = _case(name)
    (Ar_body, Ar_film, Br_T, Crs, VT0, VT1), info = transform_device(*args, return_info=True)
    S, dim = Br_T.shape
    assert info["pcg_solves"] == S * (dim - 1)
    # Krylov vectors amplify solver rounding by ~3e3 (thermca_b200/mor_basis.py): 1e-6 on the operators
    assert _rel(Ar_body, ref["Ar_body"]) < 1e-6
    assert _rel(Ar_film, ref["Ar_film"]) < 1e-6
    assert _rel(Br_T, ref["Br_T"]) < 1e-9            # only the first basis vector contributes
    assert _rel(Crs, ref["Crs"]) < 1e-6
    assert _rel(VT1, ref["VT1"]) < 1e-6
    if "VT0" in ref:
        assert _rel(VT0, ref["VT0"]) < 1e-6
    # exact structural properties, independent of the reference
    M_is = args[0].diagonal()
    for s in range(S):
        V = VT1[s] / M_is[:, None]
        assert np.abs(V.T @ V - np.eye(dim)).max() < 1e-12          # orthonormal basis
        assert np.abs(VT0[s] @ VT1[s] - np.eye(dim)).max() < 1e-10   # left inverse
        assert np.abs(Ar_body[s] - Ar_body[s].T).max() < 1e-12 * np.abs(Ar_body[s]).max()
        assert np.abs(Br_T[s][1:]).max() < 1e-12 * abs(Br_T[s][0])   # b spans the first basis vector


def test_reduced_model_reproduces_full_model_gain():
    """Steady-state gain of the reduced model (input on surface s -> mean temperature of surface s) equals the
    full model's: the inverse-Krylov space contains A^-1 b, so the match is to solver accuracy."""
    from thermca_b200.mor_basis import transform_device
    from thermca_b200.stationary import spd_solve
    args, _ = _case("plate_mor")
    M_is, A_body, A_film, A, B, C, dim = args
    Ar_body, Ar_film, Br_T, Crs, VT0, VT1 = transform_device(*args)
    for s in range(B.shape[1]):
        x_full = spd_solve(A, B[:, s], rtol=1e-13)
        gain_full = C[:, s] @ x_full
        Ar = Ar_body[s] + Ar_film[:, s].sum(axis=0)                  # films of 1 on every linked surface
        gain_red = Crs[s][:, s] @ np.linalg.solve(Ar, Br_T[s])
        assert abs(gain_red - gain_full) < 1e-7 * abs(gain_full)
