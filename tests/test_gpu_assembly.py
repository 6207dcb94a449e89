"""P1 operator build on the device vs the reference's
``FESystem.fem_matrices_from_skfem`` outputs (tests/golden/mesh_ops.npz, oracle/make_golden_meshops.py)
and vs the host restatement ``thermca_b200.synth.p1_operators``."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _mesh(name):
    z = np.load(os.path.join(GOLDEN, "mesh_ops.npz"))
    S = int(z[name + "_nsurf"])
    n = len(z[name + "_points"])
    ref = {
        "Lx": sp.csr_matrix((z[name + "_Lx_data"], z[name + "_Lx_indices"], z[name + "_Lx_indptr"]), shape=(n, n)),
        "Mx": z[name + "_Mx"], "Qs": z[name + "_Qs"], "areas": z[name + "_areas"],
        "sdofs": [z[f"{name}_sdofs{s}"] for s in range(S)], "names": [str(x) for x in z[name + "_surf_names"]],
    }
    return z[name + "_points"], z[name + "_tets"], [z[f"{name}_tris{s}"] for s in range(S)], ref


def _to_scipy(asm):
    n = asm["n"]
    return sp.csr_matrix((asm["data"].cpu().numpy(), asm["indices"].cpu().numpy(), asm["indptr"].cpu().numpy()),
                         shape=(n, n))


@pytest.mark.parametrize("name", ["plate", "constr_cyl"])
def test_assembly_matches_reference_operators(name):
    from thermca_b200.assembly import p1_assemble_device
    pts, tets, tris, ref = _mesh(name)
    asm = p1_assemble_device(pts, tets, tris)
    Lx = _to_scipy(asm)
    assert np.array_equal(Lx.indptr, ref["Lx"].indptr)                       # same sparsity pattern, bit-exact
    assert np.array_equal(Lx.indices, ref["Lx"].indices)
    scale = np.abs(ref["Lx"].data).max()
    assert np.abs(Lx.data - ref["Lx"].data).max() <= 1e-13 * scale
    assert np.allclose(asm["Mx_diag"].cpu().numpy(), ref["Mx"], rtol=1e-13, atol=0)
    for s in range(len(tris)):
        idx = asm["Qs_idx"][s].cpu().numpy()
        assert np.array_equal(idx, ref["sdofs"][s])                          # surface DOFs: sorted unique vertices
        assert np.array_equal(idx, np.nonzero(ref["Qs"][:, s])[0])
        assert np.allclose(asm["Qs_val"][s].cpu().numpy(), ref["Qs"][idx, s], rtol=1e-13, atol=0)
    assert np.allclose(asm["surf_areas"], ref["areas"], rtol=1e-13)
    # rows of a stiffness matrix sum to zero
    assert np.abs(np.asarray(Lx.sum(axis=1)).ravel()).max() < 1e-12 * scale


def test_assembly_matches_host_restatement_on_jittered_cuboid():
    from thermca_b200.assembly import p1_assemble_device
    from thermca_b200.synth import kuhn_cuboid_mesh, p1_operators
    pts, tets, tb, tu = kuhn_cuboid_mesh(9, 14, 4, 0.5, 1.0, 0.05, jitter=0.15)
    Lx_h, Mx_h, Qi_h, Qv_h, areas_h = p1_operators(pts, tets, [tb, tu])
    asm = p1_assemble_device(pts, tets, [tb, tu])
    Lx = _to_scipy(asm)
    assert np.array_equal(Lx.indptr, Lx_h.indptr) and np.array_equal(Lx.indices, Lx_h.indices)
    assert np.abs(Lx.data - Lx_h.data).max() <= 1e-12 * np.abs(Lx_h.data).max()
    assert np.allclose(asm["Mx_diag"].cpu().numpy(), Mx_h, rtol=1e-13, atol=0)
    for s in range(2):
        assert np.array_equal(asm["Qs_idx"][s].cpu().numpy(), Qi_h[s])
        assert np.allclose(asm["Qs_val"][s].cpu().numpy(), Qv_h[s], rtol=1e-12, atol=0)
    # run-to-run identical (no atomics)
    asm2 = p1_assemble_device(pts, tets, [tb, tu])
    assert np.array_equal(asm2["data"].cpu().numpy(), asm["data"].cpu().numpy())


def test_hollow_cylinder_conductance_known_answers():
    """thermca/tests/test_fe_sys.py:71-133 on operators assembled AND solved on the device:
    axial conductance 2.0455, radial 6.0733 (LP model at 512x512), 7 % tolerance as in the reference test."""
    from thermca_b200.assembly import p1_assemble_device
    from thermca_b200.stationary import spd_solve
    pts, tets, tris, ref = _mesh("constr_cyl")
    asm = p1_assemble_device(pts, tets, tris)
    L = _to_scipy(asm)                     # conductivity 1
    n = asm["n"]
    names = ref["names"]

    def solve(heat_surf, temp_surf):
        q = np.zeros(n)
        s = names.index(heat_surf)
        idx = asm["Qs_idx"][s].cpu().numpy()
        val = asm["Qs_val"][s].cpu().numpy()
        q[idx] = val / val.sum()                                    # 1 W spread by node area
        bound = asm["Qs_idx"][names.index(temp_surf)].cpu().numpy()
        free = np.setdiff1d(np.arange(n), bound)
        temps = np.zeros(n)
        temps[free] = spd_solve(L[free][:, free], q[free], rtol=1e-12)
        c = val / val.mean() / len(val)                             # C_surf column, fe_system.py:152-160
        return 1.0 / float(c @ temps[idx])

    assert np.isclose(solve("base", "half_end"), 2.0455, rtol=0.07)
    assert np.isclose(solve("outer", "half_inner"), 6.0733, rtol=0.07)


def test_device_mesh_spec_runs_and_matches_host_spec():
    """Operator built on the device -> SELL -> solver: RHS equals the host-assembled lowered network."""
    from thermca_b200.device import DeviceSolver
    from thermca_b200.synth import kuhn_cuboid_mesh, cuboid_spec
    from thermca_b200.spec import state_length
    from thermca_b200.workloads import device_mesh_spec
    dims = (7, 11, 3)
    pts, tets, tb, tu = kuhn_cuboid_mesh(*dims, 0.5, 1.0, 0.05, jitter=0.15)
    spec_d = device_mesh_spec(pts, tets, tb, tu)
    spec_h = cuboid_spec(*dims, jitter=0.15)
    rng = np.random.default_rng(1)
    y = 20.0 + rng.standard_normal(state_length(spec_h))
    f = []
    for spec in (spec_d, spec_h):
        dev = DeviceSolver(spec)
        f.append(dev.rhs(0.0, y))
        dev.close()
    assert np.abs(f[0] - f[1]).max() <= 1e-11 * np.abs(f[1]).max()
    # and an implicit step on it
    dev = DeviceSolver(spec_d)
    dev.set_state(np.full(state_length(spec_d), 20.0))
    dev.theta_steps(3, 1.0, theta=0.5, pcg_rtol=1e-10)
    yd = dev.get_state()
    dev.close()
    dev = DeviceSolver(spec_h)
    dev.set_state(np.full(state_length(spec_h), 20.0))
    dev.theta_steps(3, 1.0, theta=0.5, pcg_rtol=1e-10)
    yh = dev.get_state()
    dev.close()
    assert np.abs(yd - yh).max() < 1e-8


def test_fem_matrices_drop_in_returns_the_reference_tuple():
    """`fem_matrices_device(body_mesh, surf_meshes)` = the tuple FESystem.__init__ unpacks (fe_system.py:114-125),
    compared element by element with what the reference's own function returned for plate.msh."""
    from types import SimpleNamespace
    from thermca_b200.assembly import fem_matrices_device
    pts, tets, tris, ref = _mesh("plate")
    body = SimpleNamespace(points=pts, cell_blocks=[tets])
    surfs = SimpleNamespace(cell_blocks=tris, block_names=ref["names"])
    Mx, Lx, Qs, v2d, d2v, dof, sdofs, areas = fem_matrices_device(body, surfs)
    assert dof == len(pts) and np.array_equal(v2d, np.arange(dof)) and np.array_equal(d2v, np.arange(dof))
    assert sp.issparse(Lx) and Lx.format == "csr" and Lx.shape == (dof, dof)
    assert np.abs((Lx - ref["Lx"]).data).max() <= 1e-13 * np.abs(ref["Lx"].data).max()
    assert np.allclose(Mx, ref["Mx"], rtol=1e-13, atol=0)
    assert Qs.shape == ref["Qs"].shape and np.allclose(Qs, ref["Qs"], rtol=1e-13, atol=0)
    assert all(np.array_equal(a, b) for a, b in zip(sdofs, ref["sdofs"]))
    assert np.allclose(areas, ref["areas"], rtol=1e-13)
