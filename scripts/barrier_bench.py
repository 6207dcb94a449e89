"""ns per grid barrier on this GPU: reduce-barrier (csrc/gridbar.cuh) vs cooperative_groups grid.sync()."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from thermca_b200 import _cabi
lib = _cabi.load_library()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
out = []
for threads, bps in ((256, 8), (256, 4), (512, 4), (512, 2), (1024, 2), (1024, 1), (256, 1)):
    row = {"threads": threads, "blocks_per_sm": bps}
    for name, K, mode in (("grid.sync", 0, 1), ("barrier", 0, 0), ("reduce1", 1, 0), ("reduce2", 2, 0),
                          ("last1", 1, 2), ("last2", 2, 2)):
        ns = C.c_double()
        _cabi.check(lib.tc_bar_bench(st, 148 * bps, threads, 2000, K, mode, C.byref(ns)))
        row[name] = round(ns.value, 1)
    print(row, flush=True)
    out.append(row)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
