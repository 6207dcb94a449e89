import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from thermca_b200.device import DeviceSolver
from thermca_b200.spec import load_spec
for name in ("plate_mor", "plate_mor_var", "assembly"):
    spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    t_end = 32428.0 if name == "plate_mor" else 3000.0
    for method in ("RK45", "ROS2"):
        dev = DeviceSolver(spec)
        t0 = time.perf_counter()
        if method == "RK45":
            out = dev.run_rk(0.0, t_end, 1e-4, 1e-6, "RK45", num_skip=50)
        else:
            out = dev.run_ros2(0.0, t_end, 1e-4, 1e-6, num_skip=50)
        dt = time.perf_counter() - t0
        print(name, method, "accepted", out["accepted"], "rejected", out["rejected"], f"{dt:.2f} s", "T_end", out["net_temps"][-1][:4], flush=True)
        dev.close()
