"""Scale check of the device steady-state solve: ~1M-DOF jittered cuboid, 1 kW on the bottom, film 10 W/m2K to 20 C on the
other faces (equilibrium of BASELINE config #2)."""
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

from thermca_b200.assembly import p1_assemble_device
from thermca_b200.stationary import spd_solve
from thermca_b200.synth import kuhn_cuboid_mesh
from thermca_b200.workloads import cuboid_dims

nx, ny, nz = cuboid_dims(int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000)
pts, tets, tb, tu = kuhn_cuboid_mesh(nx, ny, nz, 0.5, 1.0, 0.05, jitter=0.15)
asm = p1_assemble_device(pts, tets, [tb, tu])
n = asm["n"]
L = 50.0 * sp.csr_matrix((asm["data"].cpu().numpy(), asm["indices"].cpu().numpy(), asm["indptr"].cpu().numpy()), shape=(n, n))
q = [np.zeros(n), np.zeros(n)]
for s in range(2):
    q[s][asm["Qs_idx"][s].cpu().numpy()] = asm["Qs_val"][s].cpu().numpy()
K = (L + sp.diags(10.0 * q[1])).tocsr()                       # FilmBC folded into the diagonal (fe_system.py:436-442)
b = q[0] * (1000.0 / q[0].sum()) + 10.0 * 20.0 * q[1]          # HeatBC + film * env_temp
torch.cuda.synchronize()
t0 = time.time()
x, info = spd_solve(K, b, rtol=1e-10, return_info=True)
dt = time.time() - t0
area = q[1].sum()
print(f"n = {n}: {info['iterations']} PCG iterations, relative residual {info['relres']:.2e}, {dt:.2f} s including the host-side "
      f"SELL conversion and upload; mean film-surface temperature {float(q[1] @ x / area):.4f} C "
      f"(energy balance: 20 + 1000 / (10 * {area:.4f}) = {20 + 1000 / (10 * area):.4f})")
