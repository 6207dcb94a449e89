"""Bandwidth of the multi-right-hand-side PCG (csrc/pcg_multi.cu) on the ~1M-DOF cuboid operator: time per matrix sweep
for K = 1..8 (fixed 300 iterations) against the algorithmic bytes 12*nnz_row + 4 + K*113 per row."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

from thermca_b200.assembly import p1_assemble_device
from thermca_b200.linsolve import DeviceOperator
from thermca_b200.synth import kuhn_cuboid_mesh
from thermca_b200.workloads import cuboid_dims

nx, ny, nz = cuboid_dims(int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000)
pts, tets, tb, tu = kuhn_cuboid_mesh(nx, ny, nz, 0.5, 1.0, 0.05, jitter=0.15)
asm = p1_assemble_device(pts, tets, [tb, tu])
n = asm["n"]
L = 50.0 * sp.csr_matrix((asm["data"].cpu().numpy(), asm["indices"].cpu().numpy(), asm["indptr"].cpu().numpy()), shape=(n, n))
K_mat = (L + sp.diags(np.full(n, 1e-3 * L.diagonal().mean()))).tocsr()
dev = torch.device("cuda")
op = DeviceOperator(K_mat, dev)
nnz_row = K_mat.nnz / n
rng = np.random.default_rng(0)
iters = 300
for K in (1, 2, 4, 6, 8):
    B = torch.from_numpy(rng.standard_normal((max(K, 2), n))).to(dev)[:K] if K > 1 else torch.from_numpy(rng.standard_normal((2, n))).to(dev)
    Kk = B.shape[0]
    op.solve_multi(B, 1e-30, maxit=20)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    op.solve_multi(B, 1e-30, maxit=iters)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    per = dt / iters
    bytes_iter = n * (12 * nnz_row + 4 + Kk * 113)
    print(f"K = {Kk}: {1e6 * per:.1f} us per sweep, algorithmic {bytes_iter / per / 1e9:.0f} GB/s, "
          f"{1e6 * per / Kk:.1f} us per right-hand side and iteration", flush=True)
