"""Host-side logic: index maps, tracer, SELL layout, spec IO, C ABI exports (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN_MODELS, ROOT
from thermca_b200 import _cabi
from thermca_b200.device import csr_to_sell, csr_diag_positions
from thermca_b200.lowering import csr_data_idxs, csr_sub_matrix_idxs
from thermca_b200.spec import save_spec, load_spec
from thermca_b200.trace import trace, TableRegistry, run_program, UntraceableError


@pytest.mark.parametrize("name", GOLDEN_MODELS)
def test_index_maps_bit_exact(golden, name):
    """diag / sub-matrix back-maps equal the reference's own (thermca/solver.py:116-122)."""
    spec, ex = golden(name)
    for key in ("diag_didx", "uu_didx", "uk_didx"):
        assert np.array_equal(np.asarray(spec[key], dtype=np.int64), np.asarray(ex["ref_maps"][key], dtype=np.int64))


def test_csr_data_idxs_first_match_and_error():
    m = sp.csr_matrix(np.array([[1.0, 0, 2], [0, 3, 0], [4, 0, 5]]))
    assert list(csr_data_idxs(m.indptr, m.indices, [0, 2, 1], [2, 0, 1])) == [1, 3, 2]
    with pytest.raises(IndexError):
        csr_data_idxs(m.indptr, m.indices, [1], [0])
    ip, ix, orig = csr_sub_matrix_idxs(m.indptr, m.indices, 0, 2, 1, 3)
    assert list(ip) == [0, 1, 2] and list(ix) == [1, 0] and list(orig) == [1, 2]


def test_sell_layout_roundtrip():
    rng = np.random.default_rng(1)
    m = sp.random(70, 70, density=0.1, random_state=3, format="csr") + sp.eye(70, format="csr")
    m = sp.csr_matrix(m)
    s = csr_to_sell(m.indptr, m.indices, m.data)
    assert s["nslices"] == 3 and s["slice_ptr"][-1] == len(s["val"])
    x = rng.standard_normal(70)
    y = np.zeros(70)
    for sl in range(s["nslices"]):
        base, w = s["slice_ptr"][sl], (s["slice_ptr"][sl + 1] - s["slice_ptr"][sl]) // 32
        for lane in range(32):
            r = sl * 32 + lane
            if r < 70:
                for k in range(w):
                    y[r] += s["val"][base + 32 * k + lane] * x[s["col"][base + 32 * k + lane]]
    assert np.allclose(y, m @ x, rtol=1e-14)
    assert np.array_equal(s["val"][s["pos"]], m.data)
    d = csr_diag_positions(m.indptr, m.indices)
    assert np.allclose(m.data[d], m.diagonal())


def test_tracer_where_interp_and_inputs():
    tabs = TableRegistry()
    inp = np.array(2.5)
    xp, fp = np.array([0.0, 10.0, 20.0]), np.array([1.0, 3.0, 2.0])

    def f(a, b, scale=inp):
        k = np.interp((a + b) / 2.0, xp, fp)
        return np.where(a > b, k * abs(a - b) ** 0.25, 0.5 * k) * scale
    p = trace(f, 2, [inp], tabs)
    code, consts, res = p.arrays()
    a, b = np.array([5.0, 25.0, -3.0]), np.array([7.0, 1.0, -3.0])
    got = run_program(code, consts, res, [a, b], [4.0], tabs.tables)
    inp[...] = 4.0
    assert np.array_equal(got, f(a, b))


def test_tracer_follows_python_branches_and_rejects_data_dependent_loops():
    """`if` on data (thermca/lib/bearing.py:131, free_conv.py:317-323) is traced path by path and joined with
    WHERE; a loop whose trip count depends on data has no finite set of paths and is rejected."""
    tabs = TableRegistry()

    def f(a, b):
        if a > b:
            out = a * 2.0
        elif a > 0.0 and b < -1.0:
            out = b - a
        else:
            out = np.sqrt(abs(a))
        return out + 1.0
    p = trace(f, 2, [], tabs)
    code, consts, res = p.arrays()
    a, b = np.array([3.0, -1.0, 0.5, 2.0, -4.0]), np.array([1.0, 2.0, -3.0, 2.0, -2.0])
    got = run_program(code, consts, res, [a, b], [], tabs.tables)
    want = np.array([f(x, y) for x, y in zip(a, b)])
    assert np.array_equal(got, want)

    def buf(a, b):                       # numpy.empty_like + per-element assignment (free_conv.vert_cyl_gap)
        n = a - b
        c = np.empty_like(n)
        for i in range(len(n)):
            if n[i] < 0.2:
                c[i] = 0.48
            else:
                c[i] = 0.93
        return c * n
    p = trace(buf, 2, [], tabs)
    code, consts, res = p.arrays()
    assert np.array_equal(run_program(code, consts, res, [a, b], [], tabs.tables), buf(a.copy(), b.copy()))

    def loop(a, b):
        x = a - b
        while x > 1.0:
            x = x / 2.0
        return x
    with pytest.raises(UntraceableError):
        trace(loop, 2, [], tabs)


def test_spec_roundtrip(tmp_path, golden):
    spec, ex = golden("plate_mor_var")
    path = tmp_path / "s.npz"
    save_spec(path, spec, extra={"a": np.arange(3)})
    spec2, ex2 = load_spec(path)
    assert np.array_equal(spec2["rc_indices"], spec["rc_indices"])
    assert np.array_equal(spec2["mor"][0]["A_films"], spec["mor"][0]["A_films"])
    assert spec2["ffe"] == [] and spec2["mor"][0]["var_film_rc_idx"] is not None
    assert np.array_equal(ex2["extra"]["a"], np.arange(3))


def test_cabi_library_exports_every_declared_symbol():
    """The .so loads and exports every function include/thermca_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "thermca_b200.h")).read()
    declared = set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", header)) - {"tc_solver"}
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    if not os.path.exists(_cabi.LIB_PATH):
        from thermca_b200.build import build
        build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.load_library().tc_version() == 100


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "thermca_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "rhs_oracle" not in src and "import oracle" not in src and "from oracle" not in src, fn


def test_stationary_install_swaps_spsolve_and_fails_loudly_without_gpu():
    """thermca_b200.stationary routes the reference's steady-state solves (fe_system.py:467-469,
    lp_system.py:250) through the device PCG; without a CUDA device it must raise, not fall back."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refenv
    if not refenv.available():
        pytest.skip("reference tree not present")
    refenv.activate()
    import thermca.fem.fe_system as fes
    import thermca.lpm.lp_system as lps
    from thermca_b200 import stationary
    orig = (fes.spsolve, lps.spsolve)
    stationary.install()
    try:
        assert fes.spsolve is stationary.spd_solve and lps.spsolve is stationary.spd_solve
        import torch
        if not torch.cuda.is_available():
            import scipy.sparse as sp
            with pytest.raises(RuntimeError):
                fes.spsolve(sp.identity(3, format="csr"), np.ones(3))
    finally:
        stationary.uninstall()
    assert (fes.spsolve, lps.spsolve) == orig


def test_assembly_install_swaps_fem_matrices():
    """thermca_b200.assembly.install() routes FESystem's operator build (fe_system.py:114-125) to the device."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refenv
    if not refenv.available():
        pytest.skip("reference tree not present")
    refenv.activate()
    import thermca.fem.fe_system as fes
    from thermca_b200 import assembly
    orig = fes.FESystem.__dict__["fem_matrices_from_skfem"]
    assembly.install()
    try:
        assert fes.FESystem.fem_matrices_from_skfem is assembly.fem_matrices_device
        import torch
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError):
                assembly.p1_assemble_device(np.zeros((4, 3)), np.array([[0, 1, 2, 3]]), [])
    finally:
        assembly.uninstall()
    assert fes.FESystem.__dict__["fem_matrices_from_skfem"] is orig


def test_mor_basis_install_keeps_disk_cache_key():
    """thermca_b200.mor_basis.install() swaps mor_io.transform for a device build under the SAME cache name
    (thermca/_utils/func_tools.py:79-80: '<func.__name__>_<hash(args)>.pickle')."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refenv
    if not refenv.available():
        pytest.skip("reference tree not present")
    refenv.activate()
    import thermca.fem.mor_io as mio
    from thermca_b200 import mor_basis
    orig = mio.transform
    mor_basis.install()
    try:
        assert mio.transform is not orig
        assert mio.transform.__name__ == "transform" == orig.__name__
        assert mio.transform.__wrapped__.__name__ == "transform"
    finally:
        mor_basis.uninstall()
    assert mio.transform is orig


def test_package_level_install_and_uninstall():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import refenv
    if not refenv.available():
        pytest.skip("reference tree not present")
    refenv.activate()
    import thermca.solver as rs
    import thermca.fem.fe_system as fes
    import thermca.fem.mor_io as mio
    import thermca_b200
    before = (rs.transient_solution, fes.spsolve, fes.FESystem.__dict__["fem_matrices_from_skfem"], mio.transform)
    thermca_b200.install()
    try:
        assert rs.transient_solution is thermca_b200.solver.transient_solution
        assert fes.spsolve is thermca_b200.stationary.spd_solve
        assert fes.FESystem.fem_matrices_from_skfem is thermca_b200.assembly.fem_matrices_device
        assert mio.transform is not before[3]
    finally:
        thermca_b200.uninstall()
    assert (rs.transient_solution, fes.spsolve, fes.FESystem.__dict__["fem_matrices_from_skfem"], mio.transform) == before


def test_row3w_coefficients_satisfy_the_w_order_conditions():
    """The constants of the third-order W-method (oracle/row3_oracle.py == csrc/solver.cu:enqueue_row3_attempt), mapped
    back from the implementation form, satisfy all eight order conditions of a W-method of order 3 (Hairer & Wanner,
    Solving ODE II, sec. IV.7), R(inf) = 0 for the exact Jacobian, and the embedded result is of order 2."""
    import row3_oracle as R
    g = R.GAMMA
    C = np.zeros((4, 4)); C[1, 0] = R.C21; C[2, 0], C[2, 1] = R.C31, R.C32; C[3, 0], C[3, 1], C[3, 2] = R.C41, R.C42, R.C43
    At = np.zeros((4, 4)); At[1, 0] = R.AT21; At[2, 0], At[2, 1] = R.AT31, R.AT32; At[3] = At[2]
    Gam = np.linalg.inv(np.eye(4) / g - C)                    # Gamma^-1 = diag(1/gamma) - C
    A = At @ Gam
    b = np.array(R.M) @ Gam
    bh = (np.array(R.M) - np.array(R.E)) @ Gam
    al, gp = A.sum(1), Gam.sum(1)
    assert np.allclose(np.diag(Gam), g) and np.allclose(np.triu(Gam, 1), 0) and np.allclose(np.triu(A), 0, atol=1e-14)
    assert np.allclose(al, R.ALPHA, atol=1e-14)
    conds = [b.sum() - 1, b @ al - 0.5, b @ gp, b @ al ** 2 - 1 / 3, b @ (A @ al) - 1 / 6, b @ (A @ gp), b @ (Gam @ al),
             b @ (Gam @ gp)]
    assert np.max(np.abs(conds)) < 1e-13
    assert abs(1 - b @ np.linalg.solve(A + Gam, np.ones(4))) < 1e-12          # R(inf) = 0
    assert np.max(np.abs([bh.sum() - 1, bh @ al - 0.5, bh @ gp])) < 1e-13      # embedded: order 2
    for z in list(1j * np.logspace(-2, 4, 60)) + list(-np.logspace(-2, 6, 60)):
        assert abs(1 + z * b @ np.linalg.solve(np.eye(4) - z * (A + Gam), np.ones(4))) <= 1 + 1e-9


@pytest.mark.parametrize("world", [2, 3])
def test_localised_part_reassembles_the_global_operator(golden, world):
    """thermca_b200/partition.py: the rows every rank owns of a lowered FE body (RCM order, even blocks) — local CSR +
    ghost side lists — give back the global product; film-diagonal lists point at diagonal entries of the local CSR;
    surface load / mean lists are the global ones restricted to the owned rows (golden with variable films)."""
    import scipy.sparse as sp
    from thermca_b200.partition import localize_ffe_part
    spec, _ = golden("plate_ffe_var")
    part = spec["ffe"][0]
    n = int(part["dof"])
    A = sp.csr_matrix((part["A_data"], part["A_body"]["indices"], part["A_body"]["indptr"]), shape=(n, n))
    x = np.random.default_rng(0).standard_normal(n)
    y = np.zeros(n)
    seen_rows = 0
    n_film_entries = 0
    for rank in range(world):
        loc, pre, info = localize_ffe_part(part, rank, world)
        r0, r1, perm = info["r0"], info["r1"], info["perm"]
        xp = x[perm]
        Al = sp.csr_matrix((loc["A_data"], loc["A_body"]["indices"], loc["A_body"]["indptr"]), shape=(r1 - r0, r1 - r0))
        yl = Al @ xp[r0:r1]
        g = xp[pre["ghosts"]]
        assert np.all((pre["ghosts"] < r0) | (pre["ghosts"] >= r1))
        for sl in np.nonzero(pre["bidx"] >= 0)[0]:
            b = pre["bidx"][sl]
            w = (pre["gptr"][b + 1] - pre["gptr"][b]) // 32
            for lane in range(32):
                row = 32 * sl + lane
                if row < r1 - r0:
                    for k in range(w):
                        q = pre["gptr"][b] + 32 * k + lane
                        yl[row] += pre["gval"][q] * g[pre["gslot"][q]]
        y[perm[r0:r1]] = yl
        seen_rows += r1 - r0
        for ix, a in zip(loc["film_didx"], loc["film_adata"]):
            rows = np.searchsorted(loc["A_body"]["indptr"], ix, side="right") - 1
            assert np.array_equal(loc["A_body"]["indices"][ix], rows) and len(ix) == len(a)
            n_film_entries += len(ix)
        for s_, (ix, v) in enumerate(zip(loc["B_idx"], loc["B_val"])):
            full = dict(zip(np.asarray(part["B_idx"][s_]).tolist(), np.asarray(part["B_val"][s_]).tolist()))
            assert all(full[int(gi)] == bv for gi, bv in zip(perm[r0:r1][ix], v))
        for s_, (ix, v) in enumerate(zip(loc["C_idx"], loc["C_val"])):
            full = dict(zip(np.asarray(part["C_idx"][s_]).tolist(), np.asarray(part["C_val"][s_]).tolist()))
            assert all(full[int(gi)] == cv for gi, cv in zip(perm[r0:r1][ix], v))
    assert seen_rows == n
    assert n_film_entries == sum(len(ix) for ix in part["film_didx"])
    assert np.max(np.abs(y - A @ x)) <= 1e-13 * np.max(np.abs(A @ x))
