"""Wall time of complete small-model simulations (BASELINE config #1 coffee cup and the cfg#4-like assembly) on the
device next to the CPU restatement of the reference path (oracle/rhs_oracle.py, scipy RK45 as shipped)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from thermca_b200.device import DeviceSolver
from thermca_b200.spec import load_spec
from rhs_oracle import integrate

for name in ("coffeecup", "steel_sheet", "assembly", "plate_mor_var"):
    spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sim = ex["sim"]
    args = (float(sim["t0"]), float(sim["t1"]), float(sim["rel_tol"]), float(sim["abs_tol"]), str(sim["method"]))
    dev = DeviceSolver(spec)
    dev.run_rk(*args)                       # warm-up (graph capture, allocations)
    dev.close()
    t0 = time.perf_counter()
    dev = DeviceSolver(spec)
    t1 = time.perf_counter()
    out = dev.run_rk(*args)
    t2 = time.perf_counter()
    dev.close()
    c0 = time.perf_counter()
    ref = integrate(spec, [args[0], args[1]], args[2], args[3], args[4], 0)
    c1 = time.perf_counter()
    print(f"{name}: {len(out['times'])} stored steps, nfev {out['nfev']}: device setup {1e3 * (t1 - t0):.1f} ms + run "
          f"{1e3 * (t2 - t1):.1f} ms ({1e6 * (t2 - t1) / out['nfev']:.1f} us per RHS evaluation incl. stage algebra and "
          f"step control); CPU restatement {1e3 * (c1 - c0):.1f} ms", flush=True)
