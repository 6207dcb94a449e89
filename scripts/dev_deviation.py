"""Developer probe: print deviations of device runs from the golden reference runs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from thermca_b200.spec import load_spec
from thermca_b200.device import DeviceSolver

def rel(a, b):
    return np.max(np.abs(a - b) / (np.abs(b) + 1e-300))

for name in ["plate_mor_var", "plate_ffe_var", "plate_mor"]:
    spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sim, ref = ex["sim"], ex["result"]
    dev = DeviceSolver(spec)
    out = dev.run_rk(float(sim["t0"]), float(sim["t1"]), float(sim["rel_tol"]), float(sim["abs_tol"]), str(sim["method"]))
    dev.close()
    io = out["io_temps"][ref["io_snap_idx"]]
    d = np.abs(io - ref["io_temps"])
    print(name, "steps", len(out["times"]), "max abs io", d.max(), "max |io|", np.abs(ref["io_temps"]).max(),
          "rel times", rel(out["times"][1:], ref["times"][1:]), "rel net", rel(out["net_temps"], ref["net_temps"]))

spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", "plate_ffe.npz"))
ref = ex["result"]
for rtol in (1e-3, 1e-4, 1e-5):
    dev = DeviceSolver(spec)
    out = dev.run_ros2(0.0, float(ex["sim"]["t1"]), rtol=rtol, atol=1e-6, pcg_rtol=1e-10)
    dev.close()
    d = np.abs(out["io_temps"][-1] - ref["io_temps"][-1])
    print("ros2 rtol", rtol, "steps", out["accepted"], "rej", out["rejected"], "max abs dev", d.max(), "at", d.argmax(),
          out["io_temps"][-1][d.argmax()], ref["io_temps"][-1][d.argmax()], "pcg", out["pcg"])
