"""Developer probe: bounded number of ROS2 attempts on a golden model, printing controller state."""
import os, sys, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from thermca_b200.spec import load_spec
from thermca_b200.device import DeviceSolver
from thermca_b200 import _cabi
name = sys.argv[1] if len(sys.argv) > 1 else "coffeecup"
spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
dev = DeviceSolver(spec)
dev._alloc_history(0)
_cabi.check(dev.lib.tc_ros2_begin(dev._handle, 0.0, 600.0, 1e-5, 1e-7, 0.0, 0.0, 1e-10, 2000, 0))
for k in range(12):
    _cabi.check(dev.lib.tc_ros2_advance(dev._handle, 50, 50))
    i = dev.info()
    print(k, {a: i[a] for a in ("done", "accepted", "rejected", "status", "stored", "t", "h_next")}, dev.pcg_stats(), flush=True)
    if i["done"]:
        _cabi.check(dev.lib.tc_reset_history(dev._handle))
print("state", dev.get_state()[:8])
