"""Step counts of the implicit schemes on the device (ROS2: order 2, ROW3: order 3) next to the reference's own LSODA
(staged reference, CPU) and explicit RK45: attempts, linear solves and the error against a tight RK45 run.
  python scripts/implicit_scheme_steps.py [out.json]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from thermca_b200.device import DeviceSolver
from thermca_b200.spec import load_spec

out_path = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else None
rows = []
cases = [("plate_ffe", None, None), ("plate_ffe", 0.0, 32428.0), ("plate_mor", 0.0, 32428.0), ("coffeecup", 0.0, 300.0),
         ("assembly", 0.0, 45.0), ("lp_tdep", 0.01, 3.0)]
for name, t0, t1 in cases:
    spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    t0 = float(ex["sim"]["t0"]) if t0 is None else t0
    t1 = float(ex["sim"]["t1"]) if t1 is None else t1
    dev = DeviceSolver(spec)
    tight = dev.run_rk(t0, t1, 1e-10, 1e-12, "RK45", num_skip=10 ** 6)
    dev.close()
    yref = np.concatenate([tight["net_temps"][-1], tight["io_temps"][-1]])
    for rtol in (1e-3, 1e-4, 1e-5):
        rec = {"model": name, "t_span": [t0, t1], "rtol": rtol}
        for scheme in ("ROS2", "ROW3"):
            dev = DeviceSolver(spec)
            tt = time.perf_counter()
            out = dev.run_ros2(t0, t1, rtol=rtol, atol=1e-9, pcg_rtol=min(1e-8, 1e-3 * rtol), num_skip=10 ** 6, scheme=scheme)
            sec = time.perf_counter() - tt
            dev.close()
            y = np.concatenate([out["net_temps"][-1], out["io_temps"][-1]])
            rec[scheme] = {"accepted": out["accepted"], "rejected": out["rejected"], "pcg_solves": out["pcg"]["solves"],
                           "err_vs_tight": float(np.max(np.abs(y - yref)) / np.max(np.abs(yref))), "seconds": sec}
        dev = DeviceSolver(spec)
        out = dev.run_rk(t0, t1, rtol, 1e-9, "RK45", num_skip=10 ** 6)
        dev.close()
        y = np.concatenate([out["net_temps"][-1], out["io_temps"][-1]])
        rec["RK45"] = {"accepted": out["accepted"], "rejected": out["rejected"],
                       "err_vs_tight": float(np.max(np.abs(y - yref)) / np.max(np.abs(yref)))}
        rows.append(rec)
        print(json.dumps(rec), flush=True)
# the reference's own LSODA on the same models (CPU, staged reference), where it is available
try:
    import refenv
    refenv.activate()
    os.chdir(tempfile.mkdtemp())
    import ref_models
    for name, t0, t1 in cases:
        for rtol in (1e-3, 1e-4, 1e-5):
            net, kw = ref_models.MODELS[name]()
            span = [kw["time_span"][0] if t0 is None else t0, kw["time_span"][1] if t1 is None else t1]
            tt = time.perf_counter()
            res = net.sim(span, method="LSODA", rel_tol=rtol, abs_tol=1e-9, progress_bar=False)
            rec = {"model": name, "rtol": rtol, "reference_LSODA_steps": len(res._data.times) - 1,
                   "seconds": time.perf_counter() - tt}
            rows.append(rec)
            print(json.dumps(rec), flush=True)
except Exception as e:
    print("reference LSODA leg skipped:", type(e).__name__, str(e)[:200])
if out_path:
    json.dump(rows, open(out_path, "w"), indent=1)
