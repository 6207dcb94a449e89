"""Scale check of the device P1 assembly: ~1M-DOF jittered Kuhn cuboid assembled from mesh arrays vs the
closed-form device generator (csrc/synth.cu) of the same mesh; prints timings."""
import sys
import time

import numpy as np
import torch

from thermca_b200.assembly import device_fe_operator
from thermca_b200.synth import kuhn_cuboid_mesh
from thermca_b200.workloads import cuboid_dims, device_cuboid_operator

target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nx, ny, nz = cuboid_dims(target)
t0 = time.time()
pts, tets, tb, tu = kuhn_cuboid_mesh(nx, ny, nz, 0.5, 1.0, 0.05, jitter=0.15)
t_mesh = time.time() - t0
print(f"mesh {nx}x{ny}x{nz}: {len(pts)} vertices, {len(tets)} tets, host mesh generation {t_mesh:.1f} s", flush=True)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    op = device_fe_operator(pts, tets, [tb, tu], 50.0, 8000.0 * 500.0)
    torch.cuda.synchronize()
    print(f"device assembly + SELL layout (incl. upload of the mesh): {time.time() - t0:.3f} s", flush=True)
ref = device_cuboid_operator(nx, ny, nz, jitter=0.15)
torch.cuda.synchronize()
assert op["n"] == ref["n"]
# compare through the action on a vector and the diagonals / masses (layouts differ: generator has fixed width 15)
print("mass  max rel diff", float(((op["mass"] - ref["mass"]).abs() / ref["mass"]).max()))
print("adiag max rel diff", float(((op["adiag"] - ref["adiag"]).abs() / ref["adiag"].abs()).max()))
qb = torch.zeros(op["n"], dtype=torch.float64, device="cuda"); qb[op["Qs_idx"][0]] = op["Qs_val"][0]
print("q_bottom max abs diff", float((qb - ref["q_bottom"]).abs().max()), "of", float(ref["q_bottom"].max()))
import ctypes as C
from thermca_b200 import _cabi
lib = _cabi.load_library()
x = torch.randn(op["n"], dtype=torch.float64, device="cuda")
ys = []
for o in (op, ref):
    y = torch.empty_like(x)
    p = lambda a: C.c_void_p(a.data_ptr()) if a is not None else C.c_void_p(0)
    _cabi.check(lib.tc_sell_spmv(C.c_void_p(torch.cuda.current_stream().cuda_stream), 0, o["n"], o["nslices"],
                                 p(o["slice_ptr"]), p(o["col"]), p(None), p(o["val"]), p(x), p(y), p(None), p(None), 0.0, 0))
    ys.append(y)
torch.cuda.synchronize()
print("A x max rel diff", float((ys[0] - ys[1]).abs().max() / ys[1].abs().max()))
