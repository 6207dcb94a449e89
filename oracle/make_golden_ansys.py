"""Generate tests/golden/plate_ansys.npz (build container only; TEST INFRASTRUCTURE ONLY).

The sample points the reference's own validation test compares against
(thermca/tests/test_fe_part.py:272-303: ANSYS, 12800 quadratic hexahedra, file 'cuboid flat ansys.csv'):
every ceil(n/10)-th row, first two dropped, columns Time / Bottom / Top vertex / Bottom vertex — plus the
DOF numbers of the two probed mesh vertices of plate.msh.

Usage:  python oracle/make_golden_ansys.py
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import refenv  # noqa: E402

refenv.activate()
from thermca import Mesh  # noqa: E402

TESTS = "/root/reference/thermca/tests"
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    res = np.genfromtxt(os.path.join(TESTS, "cuboid flat ansys.csv"), delimiter=";", names=True, case_sensitive=True)
    skip = math.ceil(len(res["Time"]) / 10)
    res = res[::skip][2:]
    mesh = Mesh.read(os.path.join(TESTS, "plate.msh"), file_format="gmsh")
    pts = np.asarray(mesh.extract_type("tetra").points)
    def vertex(p):
        d = np.linalg.norm(pts - np.asarray(p), axis=1)
        assert d.min() < 1e-9
        return int(d.argmin())
    out = {"time": res["Time"], "bottom": res["Bottom"], "top_vertex": res["Top_vertex"],
           "bottom_vertex": res["Bottom_vertex"], "dof_bottom_vertex": np.int64(vertex((0.0, 0.0, 0.05))),
           "dof_top_vertex": np.int64(vertex((0.5, 1.0, 0.05))),
           "surf_names": np.array(list(mesh.extract_type("triangle").block_names))}
    print({k: (v if np.ndim(v) == 0 else np.asarray(v)) for k, v in out.items()})
    np.savez_compressed(os.path.join(GOLDEN, "plate_ansys.npz"), **out)


if __name__ == "__main__":
    main()
