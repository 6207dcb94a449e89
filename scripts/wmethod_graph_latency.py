"""W-method attempts (ROS2 / ROW3w) replayed as one CUDA graph vs launched kernel by kernel: wall time of complete
simulations of the small reference-built models, and proof that both ways give bit-identical trajectories.

    python scripts/wmethod_graph_latency.py [--json out.json]

The script runs itself twice (TC_W_GRAPH=1 / 0 is read once per process by csrc/solver.cu)."""
import hashlib
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MODELS = ("coffeecup", "steel_sheet", "lp_tdep", "plate_ffe_var", "plate_mor_var", "assembly")


def child():
    from thermca_b200.device import DeviceSolver
    from thermca_b200.spec import load_spec
    rows = {}
    for name in MODELS:
        spec, ex = load_spec(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        sim = ex["sim"]
        t0, t1 = float(sim["t0"]), float(sim["t1"])
        for scheme in ("ROS2", "ROW3"):
            best, out = None, None
            for rep in range(3):                     # first repetition warms up (module load, allocations)
                dev = DeviceSolver(spec)
                a = time.perf_counter()
                out = dev.run_ros2(t0, t1, rtol=1e-4, atol=1e-6, scheme=scheme)
                b = time.perf_counter()
                dev.close()
                if rep and (best is None or b - a < best):
                    best = b - a
            attempts = int(out["accepted"] + out["rejected"])
            h = hashlib.sha256(np.ascontiguousarray(out["times"]).tobytes()
                               + np.ascontiguousarray(out["net_temps"]).tobytes()).hexdigest()[:16]
            rows[f"{name}/{scheme}"] = {"ms": 1e3 * best, "attempts": attempts, "us_per_attempt": 1e6 * best / attempts,
                                        "sha": h}
    print("RESULT " + json.dumps(rows))


def main():
    if os.environ.get("TC_W_GRAPH_CHILD"):
        return child()
    res = {}
    for mode in ("1", "0"):
        env = dict(os.environ, TC_W_GRAPH=mode, TC_W_GRAPH_CHILD="1")
        txt = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=1200)
        line = [l for l in txt.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(txt.stdout[-2000:], txt.stderr[-2000:])
            raise SystemExit("child failed")
        res["graph" if mode == "1" else "direct"] = json.loads(line[-1][7:])
    table = {}
    for k, g in res["graph"].items():
        d = res["direct"][k]
        table[k] = {"attempts": g["attempts"], "graph_us_per_attempt": round(g["us_per_attempt"], 1),
                    "direct_us_per_attempt": round(d["us_per_attempt"], 1),
                    "speedup": round(d["us_per_attempt"] / g["us_per_attempt"], 2),
                    "bit_identical": g["sha"] == d["sha"] and g["attempts"] == d["attempts"]}
        print(k, table[k], flush=True)
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(table, f, indent=1)
    assert all(r["bit_identical"] for r in table.values()), "graph replay and direct launches differ"


if __name__ == "__main__":
    main()
