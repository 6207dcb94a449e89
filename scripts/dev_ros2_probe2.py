import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from thermca_b200.spec import load_spec
from thermca_b200.device import DeviceSolver
from rhs_oracle import integrate
spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", "coffeecup.npz"))
t = time.time()
tight = integrate(spec, [0.0, 600.0], 1e-9, 1e-10, "RK45")
print("oracle", time.time() - t, len(tight["times"]), flush=True)
dev = DeviceSolver(spec)
t = time.time()
out = dev.run_ros2(0.0, 600.0, rtol=1e-5, atol=1e-7, pcg_rtol=1e-10, progress=lambda i: print(i, flush=True) if i["accepted"] % 200 < 4 else None)
print("ros2", time.time() - t, out["accepted"], out["rejected"], out["pcg"], flush=True)
print(np.max(np.abs(out["net_temps"][-1] - tight["net_temps"][-1]) / np.abs(tight["net_temps"][-1])))
