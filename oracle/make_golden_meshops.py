"""Generate tests/golden/mesh_ops.npz (build container only; TEST INFRASTRUCTURE ONLY).

For the two Gmsh meshes of the reference's test-suite that are readable here (plate.msh,
constr_cyl.msh) store the mesh arrays and what the reference's own
``FESystem.fem_matrices_from_skfem`` (thermca/fem/fe_system.py:322-400) returns for them —
through the scikit-fem stand-in of oracle/shims (scikit-fem 10.0.2 is not installed; the stand-in
is pinned by the reference tests that pass on top of it).

Usage:  python oracle/make_golden_meshops.py
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
from thermca import Mesh  # noqa: E402
from thermca.fem.fe_system import FESystem  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
TESTS = "/root/reference/thermca/tests"


def main():
    out = {}
    for name in ("plate", "constr_cyl"):
        mesh = Mesh.read(os.path.join(TESTS, name + ".msh"), file_format="gmsh")
        body, surfs = mesh.extract_type("tetra"), mesh.extract_type("triangle")
        Mx, Lx, Qs, v2d, d2v, dof, sdofs, areas = FESystem.fem_matrices_from_skfem(body, surfs)
        assert np.array_equal(v2d, np.arange(dof))
        Lx = sp.csr_matrix(Lx)
        Lx.sort_indices()
        out[name + "_points"] = np.asarray(body.points, dtype=np.float64)
        out[name + "_tets"] = np.asarray(body.cell_blocks[0], dtype=np.int32)
        out[name + "_nsurf"] = np.int64(len(surfs.cell_blocks))
        out[name + "_surf_names"] = np.array(list(surfs.block_names))
        for s, cells in enumerate(surfs.cell_blocks):
            out[f"{name}_tris{s}"] = np.asarray(cells, dtype=np.int32)
            out[f"{name}_sdofs{s}"] = np.asarray(sdofs[s], dtype=np.int64)
        out[name + "_Lx_indptr"] = Lx.indptr.astype(np.int64)
        out[name + "_Lx_indices"] = Lx.indices.astype(np.int64)
        out[name + "_Lx_data"] = Lx.data
        out[name + "_Mx"] = np.asarray(Mx, dtype=np.float64)
        out[name + "_Qs"] = np.asarray(Qs, dtype=np.float64)
        out[name + "_areas"] = np.asarray(areas, dtype=np.float64)
        print(name, dof, Lx.nnz, list(surfs.block_names), areas)
    np.savez_compressed(os.path.join(GOLDEN, "mesh_ops.npz"), **out)


if __name__ == "__main__":
    main()
