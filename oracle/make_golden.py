"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  For every model in ``oracle/ref_models.py``:
1. build ``Network(model)`` through the real thermca API,
2. lower it (``thermca_b200.lowering.lower_network``) BEFORE simulating,
3. run the reference ``Network.sim`` while recording the (t, y, f(t,y)) triples its own
   ``update_time_derivative`` (thermca/solver.py:193-302) sees, via a wrapper around
   ``thermca.solver.simple_solve_ivp``,
4. store the lowered network, the recorded right-hand-side calls, the reference's own
   index maps (thermca/solver.py:116-122) and the collected results.

Usage:  python oracle/make_golden.py [model ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import refenv  # noqa: E402

refenv.activate()
import thermca.solver as ref_solver  # noqa: E402
from thermca._utils.sparse import to_csr_data_idxs, csr_sub_matrix  # noqa: E402

import ref_models  # noqa: E402
from thermca_b200.lowering import lower_network, initial_state  # noqa: E402
from thermca_b200.spec import save_spec  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAX_RHS = {"default": 48, "plate_ffe": 14, "plate_ffe_var": 14, "plate_ffe_stat": 8, "plate_ffe_statc": 8, "kuhn_cuboid": 20, "tdep_rod": 20, "assembly": 16}
MAX_IO_SNAPSHOTS = 8


def run(name):
    os.makedirs("/tmp/thermca_golden_cwd", exist_ok=True)
    os.chdir("/tmp/thermca_golden_cwd")          # disk_cache writes ./.thermca_cache
    net, kw = ref_models.MODELS[name]()
    spec = lower_network(net)
    y0 = initial_state(spec)

    calls = []
    orig = ref_solver.simple_solve_ivp
    stats = {}

    def recording(func, t_span, y0_, method="RK45", func_args=None, step_end_func=None, **options):
        def wrapped(t, y):
            out = func(t, y)
            if len(calls) < MAX_RHS.get(name, MAX_RHS["default"]):
                calls.append((float(t), np.array(y, dtype=np.float64), np.array(out, dtype=np.float64),
                              net.run_cond.data.copy(), np.array(net.heat_src), np.array(net.capy)))
            return out
        res = orig(wrapped, t_span, y0_, method, func_args, step_end_func, **options)
        stats["nfev"] = int(res.nfev)
        return res

    ref_solver.simple_solve_ivp = recording
    t0 = time.time()
    try:
        coll = ref_solver.transient_solution(
            tuple(float(t) for t in kw["time_span"]), net, kw.get("rel_tol", 1e-3),
            kw.get("abs_tol", 1e-6), kw.get("method", "RK45"), 0)
    finally:
        ref_solver.simple_solve_ivp = orig
    wall = time.time() - t0

    # the reference's own index maps (bit-exact parity targets)
    u, N = net._free_temp_lgth, net._node_count
    rc = net.run_cond
    diag = np.arange(u)
    ref_maps = {
        "diag_didx": np.asarray(to_csr_data_idxs(rc, (diag, diag)), dtype=np.int64),
        "uu_didx": np.asarray(csr_sub_matrix(rc, 0, u, 0, u)[1], dtype=np.int64),
        "uk_didx": np.asarray(csr_sub_matrix(rc, 0, u, u, N)[1], dtype=np.int64),
    }
    times = np.array(coll.times)
    nsteps = len(times)
    snap = np.unique(np.linspace(0, nsteps - 1, min(MAX_IO_SNAPSHOTS, nsteps)).round().astype(int))
    io_temps = np.array([coll.io_temps[i] for i in snap]) if coll.io_temps else np.zeros((len(snap), 0))
    result = {
        "times": times, "net_temps": np.array(coll.net_temps),
        "net_capys": np.array(coll.net_capys), "net_heat_srcs": np.array(coll.net_heat_srcs),
        "io_snap_idx": snap.astype(np.int64), "io_temps": io_temps,
        "clean_conds_last": np.asarray(coll.net_clean_conds[-1]),
        "nfev": np.int64(stats.get("nfev", 0)), "wall_s": np.float64(wall),
    }
    rhs = {"t": np.array([c[0] for c in calls]), "y": np.array([c[1] for c in calls]),
           "f": np.array([c[2] for c in calls]), "cond": np.array([c[3] for c in calls]),
           "heat": np.array([c[4] for c in calls]), "capy": np.array([c[5] for c in calls])}
    sim = {"t0": np.float64(kw["time_span"][0]), "t1": np.float64(kw["time_span"][1]),
           "rel_tol": np.float64(kw.get("rel_tol", 1e-3)), "abs_tol": np.float64(kw.get("abs_tol", 1e-6)),
           "method": kw.get("method", "RK45")}
    probe = getattr(net, "_golden_probe", {})
    path = os.path.join(GOLDEN, name + ".npz")
    save_spec(path, spec, y0=y0, rhs=rhs, result=result, ref_maps=ref_maps, sim=sim,
              probe={k: np.int64(v) for k, v in probe.items()} or {"none": np.int64(0)})
    print(f"{name}: N={N} u={u} state={len(y0)} steps={nsteps - 1} nfev={stats.get('nfev')} "
          f"wall={wall:.2f}s -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    for name in (sys.argv[1:] or list(ref_models.MODELS)):
        run(name)
