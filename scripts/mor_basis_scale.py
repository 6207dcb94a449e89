"""Scale check of the device MOR basis build: ~1M-DOF cuboid (2 coupling surfaces), dim = 30."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

from thermca_b200.assembly import p1_assemble_device
from thermca_b200.mor_basis import transform_device
from thermca_b200.synth import kuhn_cuboid_mesh, c_surf_columns
from thermca_b200.workloads import cuboid_dims

target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 30
nx, ny, nz = cuboid_dims(target)
pts, tets, tb, tu = kuhn_cuboid_mesh(nx, ny, nz, 0.5, 1.0, 0.05, jitter=0.15)
asm = p1_assemble_device(pts, tets, [tb, tu])
n = asm["n"]
Lx = sp.csr_matrix((asm["data"].cpu().numpy(), asm["indices"].cpu().numpy(), asm["indptr"].cpu().numpy()), shape=(n, n))
M_diag = 8000.0 * 500.0 * asm["Mx_diag"].cpu().numpy()
L = 50.0 * Lx
Qs = np.zeros((n, 2))
for s in range(2):
    Qs[asm["Qs_idx"][s].cpu().numpy(), s] = asm["Qs_val"][s].cpu().numpy()
C_surf = np.zeros((n, 2))
for s, c in enumerate(c_surf_columns([asm["Qs_val"][k].cpu().numpy() for k in range(2)])):
    C_surf[asm["Qs_idx"][s].cpu().numpy(), s] = c
# to_mor_state_space (thermca/fem/mor_io.py:95-136) with one film (value 1) on surface 1
M_is = sp.diags(1.0 / np.sqrt(M_diag), format="csr")
A_body = (M_is @ L @ M_is).tocsr()
A_film = [(M_is @ sp.diags(Qs[:, 1], format="csr") @ M_is).tocsr()]
A = (A_body + A_film[0]).tocsr()
B = M_is @ Qs
C = M_is @ C_surf
print(f"n = {n}, nnz = {A.nnz}, dim = {dim}", flush=True)
torch.cuda.synchronize()
t0 = time.time()
(Ar_body, Ar_film, Br_T, Crs, VT0, VT1), info = transform_device(M_is, A_body, A_film, A, B, C, dim, pcg_rtol=1e-12,
                                                                 return_info=True)
t1 = time.time() - t0
print(f"device basis build: {t1:.1f} s, {info['pcg_solves']} solves, {info['pcg_iterations']} PCG iterations "
      f"({info['pcg_iterations'] / info['pcg_solves']:.0f} per solve, "
      f"{1e6 * t1 / info['pcg_iterations']:.1f} us per iteration incl. everything else)", flush=True)
mis = M_is.diagonal()
for s in range(2):
    V = VT1[s] / mis[:, None]
    print("surface", s, "orthonormality", np.abs(V.T @ V - np.eye(dim)).max(), "left inverse",
          np.abs(VT0[s] @ VT1[s] - np.eye(dim)).max())
