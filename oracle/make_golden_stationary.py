"""Generate tests/golden/stationary.npz from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Records every ``spsolve(lhs, rhs)`` the reference's steady-state
solvers issue (thermca/fem/fe_system.py:467-469, thermca/lpm/lp_system.py:250) while the
reference's own known-answer tests run (thermca/tests/test_fe_sys.py:71-133 hollow cylinder,
thermca/tests/test_lp_sys.py) plus the plate mesh with a heat and a film boundary condition.
The tests' own assertions pin the recorded solutions.

Usage:  python oracle/make_golden_stationary.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import refenv  # noqa: E402

refenv.activate()
import thermca.fem.fe_system as fes  # noqa: E402
import thermca.lpm.lp_system as lps  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
records = []
label = ["?"]


def recorder(orig):
    def spsolve(lhs, rhs):
        x = orig(lhs, rhs)
        m = sp.csr_matrix(lhs)
        m.sort_indices()
        records.append((label[0], m, np.array(rhs, dtype=np.float64).ravel(), np.array(x, dtype=np.float64).ravel()))
        return x
    return spsolve


def main():
    os.makedirs("/tmp/thermca_golden_cwd", exist_ok=True)
    os.chdir("/root/reference/thermca/tests")            # the tests read their meshes from the working directory
    fes.spsolve = recorder(fes.spsolve)
    lps.spsolve = recorder(lps.spsolve)
    import thermca.tests.test_fe_sys as tfe
    import thermca.tests.test_lp_sys as tlp
    label[0] = "fe_hollow_cylinder"
    tfe.test_stationary_solution_of_asymmetrical_heat_flow_in_hollow_cylinder()
    for name in sorted(dir(tlp)):
        if name.startswith("test_"):
            label[0] = "lp_" + name[5:]
            n0 = len(records)
            try:
                getattr(tlp, name)()
            except Exception as e:                          # a test that cannot run under the shims is skipped
                del records[n0:]
                print("skipped", name, type(e).__name__, e)
    # plate mesh: heat on the bottom, film on every other face (the transient goldens' model, at equilibrium)
    from thermca import Mesh, Solid
    from thermca.static_bcs import HeatBC, FilmBC
    from dataclasses import dataclass

    @dataclass
    class Surf:
        name: str
    mesh = Mesh.read("plate.msh", file_format="gmsh")
    plate = fes.FESystem(body_mesh=mesh.extract_type("tetra"), surf_meshes=mesh.extract_type("triangle"),
                         init_temp=20.0, matl=Solid(dens=8000.0, spec_heat=500.0, condy=50.0), lump_cond_meth=None)
    label[0] = "fe_plate_heat_film"
    bcs = [HeatBC(Surf("bottom"), 1000.0)] + [FilmBC(Surf(s), film=10.0, env_temp=20.0)
                                              for s in ("top", "left", "right", "front", "back")]
    temps = plate.stationary_solution(bcs)
    mean = plate.C_surf.T @ temps
    print("plate mean surface temps", mean)

    out = {"n": np.int64(len(records))}
    for k, (lab, m, b, x) in enumerate(records):
        out[f"{k}_label"] = np.array(lab)
        out[f"{k}_indptr"] = m.indptr.astype(np.int64)
        out[f"{k}_indices"] = m.indices.astype(np.int64)
        out[f"{k}_data"] = m.data
        out[f"{k}_shape"] = np.array(m.shape, dtype=np.int64)
        out[f"{k}_rhs"] = b
        out[f"{k}_x"] = x
        print(k, lab, m.shape, m.nnz)
    np.savez_compressed(os.path.join(GOLDEN, "stationary.npz"), **out)


if __name__ == "__main__":
    main()
