"""Generate tests/golden/mor_basis.npz (build container only; TEST INFRASTRUCTURE ONLY).

Records the arguments and results of the reference's ``thermca.fem.mor_io.transform``
(thermca/fem/mor_io.py:139-166, LU-based inverse-Krylov Arnoldi) while ``Network(model)`` is built for the
MOR goldens ``plate_mor`` (dim 10) and ``plate_mor_var`` (dim 12).

Usage:  python oracle/make_golden_morbasis.py
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
import thermca.fem.mor_io as mio  # noqa: E402

import ref_models  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    os.makedirs("/tmp/thermca_golden_cwd", exist_ok=True)
    os.chdir("/tmp/thermca_golden_cwd")
    orig = mio.transform.__wrapped__                 # bypass the disk cache: always run the reference code
    rec = {}

    def spy(*a):
        out = orig(*a)
        rec["args"], rec["out"] = a, out
        return out
    mio.transform = spy
    out = {}
    for name in ("plate_mor", "plate_mor_var"):
        ref_models.MODELS[name]()
        M_is, A_body, A_film, A, B, C, dim = rec["args"]
        Ar_body, Ar_film, Br_T, Crs, VT0, VT1 = rec["out"]
        for key, m in (("A", A), ("A_body", A_body)):
            m = sp.csr_matrix(m)
            m.sort_indices()
            out[f"{name}_{key}_indptr"] = m.indptr.astype(np.int64)
            out[f"{name}_{key}_indices"] = m.indices.astype(np.int64)
            out[f"{name}_{key}_data"] = m.data
        out[name + "_M_is"] = M_is.diagonal()
        out[name + "_A_film"] = np.array([f.diagonal() for f in A_film])
        out[name + "_B"] = np.asarray(B)
        out[name + "_C"] = np.asarray(C)
        out[name + "_dim"] = np.int64(dim)
        out[name + "_Ar_body"] = Ar_body
        out[name + "_Ar_film"] = Ar_film
        out[name + "_Br_T"] = Br_T
        out[name + "_Crs"] = Crs
        out[name + "_VT1"] = VT1
        if name == "plate_mor":
            out[name + "_VT0"] = VT0
        print(name, A.shape, "dim", dim, "films", len(A_film), "S", B.shape[1])
    np.savez_compressed(os.path.join(GOLDEN, "mor_basis.npz"), **out)


if __name__ == "__main__":
    main()
