#!/usr/bin/env python
"""Headline benchmark: DOF*steps/s of the implicit transient step on a synthetic FE cuboid.

A "step" is one implicit theta step (theta = 1/2) of the stacked thermal system: one
right-hand-side evaluation (surface means -> fused network kernel -> SELL SpMV -> surface
loads) and one Jacobi-PCG solve of (M + theta*h*K) — north_star items (a)+(b)+(c).
Workload (BASELINE config #2 scaled to the size the metric is quoted on): one
tetrahedral FE cuboid, ~10 M DOF, 1 kW on the bottom face, 10 W/m2K film on the other
faces, operator generated in HBM by csrc/synth.cu.

  python bench.py --gpus 1 --steps 8 --warmup 3             (our arm)
  python bench.py --impl reference --steps 3 --warmup 1     (CPU reference arm)
  torchrun ... bench.py --gpus N ...                          (row-partitioned, N > 1)

Prints ONE JSON line (see the driver contract in the repo README / DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

THETA = 0.5
H_STEP = 1.0            # seconds of simulated time per step
PCG_RTOL = 1e-8
FILM, HEAT, T_ENV = 10.0, 1000.0, 20.0


def algorithmic_bytes_per_iteration(n_rows, nnz):
    """Jacobi-PCG iteration on CSR-equivalent storage:
    SpMV  nnz*(8+4) + rows*(4 indptr + 8 p gather + 8 mass + 8 q write)
    update x,r (+dots): read x,p,r,q,dinv + write x,r = 56 B/row
    direction        : read r,dinv,p + write p       = 32 B/row"""
    return nnz * 12 + n_rows * (28 + 56 + 32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the GPU is under load (one persistent
    `nvidia-smi -lms 100` process; started before the warm-up steps, stopped after the timed region)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            time.sleep(0.25)          # first sample before the load starts
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], None, set()
        for line in self.lines:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for nm, v in zip(self.NAMES, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_traffic(n, iters_per_launch):
    """DRAM bytes per PCG launch from the committed ncu capture (profiles/r01_pcg_traffic.json), scaled to this
    run's iteration count; None when the capture was taken on a different problem size."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_pcg_traffic.json")) as f:
            t = json.load(f)
        if abs(t["dof"] - n) > 0.02 * n:
            return None
        return t["dram_bytes_per_row_iter"] * n * iters_per_launch
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ---- CPU reference arm -------------------------------------------------------------------------
def cpu_theta_baseline(target_dof, steps, warmup, threads):
    """The like-for-like CPU port (oracle/theta_oracle.py) of the same step on a bounded
    sample of the workload: same generator, same physics, smaller mesh."""
    from thermca_b200.synth import cuboid_spec
    from thermca_b200.workloads import cuboid_dims
    from theta_oracle import theta_cg_steps
    import scipy.sparse as sp
    nx, ny, nz = cuboid_dims(target_dof)
    spec = cuboid_spec(nx, ny, nz, jitter=0.15, heat=HEAT, film=FILM, env_temp=T_ENV)
    p = spec["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    mass = 1.0 / p["M_inv"]
    b = np.zeros(n)
    flux = [HEAT / p["surf_areas"][0], FILM * T_ENV]
    for s in range(2):
        b[p["B_idx"][s]] += p["B_val"][s] * flux[s]
    y = np.full(n, 20.0)
    best = None
    runs = {}
    for backend, thr in (("scipy", 1), ("c_omp", threads)):
        try:
            yw, _, _ = theta_cg_steps(A, mass, b, y, warmup, H_STEP, THETA, PCG_RTOL, backend=backend, threads=thr)
            _, iters, sec = theta_cg_steps(A, mass, b, yw, steps, H_STEP, THETA, PCG_RTOL, backend=backend, threads=thr)
        except Exception:
            continue
        val = n * steps / sec
        runs[backend] = {"value": val, "cores": thr, "backend": backend, "iters_per_step": iters / steps,
                         "ms_per_step": 1e3 * sec / steps}
    # headline CPU figure: the reference's own path (scipy csr_matvec, one thread; MKL is not installable
    # here); the multi-threaded C/OpenMP port of the same algorithm is reported next to it
    best = dict(runs.get("scipy") or next(iter(runs.values())))
    best["all_backends"] = runs
    # the reference exactly as shipped: explicit adaptive RK45 (scipy) around the restated right-hand side
    try:
        from scipy.integrate import RK45
        from rhs_oracle import OracleRHS
        from thermca_b200.lowering import initial_state
        rhs = OracleRHS(spec)
        y0 = initial_state(spec)
        rhs(0.0, y0)
        solver = RK45(rhs, 0.0, y0, 1.0e9, rtol=1e-3, atol=1e-6)
        for _ in range(3):
            solver.step()
        nf0, t0 = solver.nfev, time.perf_counter()
        nst = 12
        for _ in range(nst):
            solver.step()
        sec = time.perf_counter() - t0
        best["rk45_as_shipped"] = {"dof_steps_per_s": n * nst / sec, "dof_rhs_per_s": n * (solver.nfev - nf0) / sec,
                                   "steps": nst, "cores": 1}
    except Exception as e:                                  # pragma: no cover
        best["rk45_as_shipped"] = {"error": str(e)[:100]}
    best.update({"dof": n, "nnz": int(A.nnz), "dims": [nx, ny, nz]})
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = cpu_theta_baseline(args.cpu_dof, args.steps, max(args.warmup, 1), threads)
    sample = (f"{r['dof']} DOF Kuhn cuboid ({r['dims'][0]}x{r['dims'][1]}x{r['dims'][2]} cells, same generator), "
              f"{args.steps} theta-PCG steps, {r['backend']} CSR SpMV")
    line = {
        "impl": "reference", "metric": "DOF*steps/s (implicit theta-PCG step, full FE cuboid)",
        "value": r["value"], "unit": "DOF*steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg#2 full FE cuboid, theta=0.5 Jacobi-PCG step (CPU port, bounded sample)",
                   "dof": r["dof"], "pcg_rtol": PCG_RTOL, "h_step_s": H_STEP,
                   "iters_per_step": r["iters_per_step"], "rk45_as_shipped": r.get("rk45_as_shipped")},
        "cpu_baseline": {"value": r["value"], "unit": "DOF*steps/s", "cores": r["cores"], "kind": "port",
                         "sample": sample, "all_backends": r["all_backends"]},
        "e2e": {"value": r["value"], "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- our arm -------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    force_dist = os.environ.get("TC_FORCE_DIST") == "1"      # developer switch: partitioned path at world 1
    if world > 1 or force_dist:
        if force_dist and world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import cuboid_dims, device_cuboid_spec

    if world > 1 or force_dist:
        from thermca_b200.dist import run_partitioned_bench
        line = run_partitioned_bench(args, world, rank, local_rank)
        if rank == 0:
            print(json.dumps(line))
        dist.destroy_process_group()
        return

    nx, ny, nz = cuboid_dims(args.dof)
    spec = device_cuboid_spec(nx, ny, nz, heat=HEAT, film=FILM, env_temp=T_ENV)
    dev = DeviceSolver(spec, use_col16=os.environ.get("TC_COL16", "1") != "0")
    n = dev.n_state
    op = spec["ffe"][0]["sell"]
    nnz = int(torch.count_nonzero(op["val"]).item())

    def sync():
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        dev.theta_steps(args.warmup, H_STEP, THETA, t0=0.0, pcg_rtol=PCG_RTOL, pcg_mode=args.pcg_mode)
        sync()
        dev.timing(True)
        it0 = dev.pcg_stats()["iterations"]
        sync()
        with torch.cuda.stream(dev.stream):
            e0.record(dev.stream)
            dev.theta_steps(args.steps, H_STEP, THETA, pcg_rtol=PCG_RTOL, pcg_mode=args.pcg_mode, sync=False)
            e1.record(dev.stream)
        sync()
    ms = e0.elapsed_time(e1)
    tm = dev.timing(False)
    phase = dev.pcg_profile()
    iters = dev.pcg_stats()["iterations"] - it0
    value = n * args.steps / (ms * 1e-3)
    peak, peak_kind = measured_peak()
    pcg_ms_per_launch = tm["pcg_ms"] / max(tm["pcg_launches"], 1)
    iters_per_solve = iters / max(tm["pcg_launches"], 1)
    bytes_per_launch = algorithmic_bytes_per_iteration(n, nnz) * iters_per_solve + n * 48   # + init phase
    achieved = bytes_per_launch / (pcg_ms_per_launch * 1e-3) / 1e9
    y_dev = dev.get_state()

    # ---- end to end through the host-buffer API: state up, step, state down ---------------------
    host_y = torch.from_numpy(y_dev).pin_memory()
    dev_y = torch.empty(n, dtype=torch.float64, device="cuda")
    e2e_steps = max(2, min(args.steps, 4))
    sync()
    t_e0 = time.perf_counter()
    for _ in range(e2e_steps):
        dev_y.copy_(host_y, non_blocking=True)
        dev.set_state(dev_y)
        dev.theta_steps(1, H_STEP, THETA, pcg_rtol=PCG_RTOL, pcg_mode=args.pcg_mode)
        host_y.copy_(dev.get_state_into(dev_y))
    sync()
    e2e_sec = time.perf_counter() - t_e0
    e2e_val = n * e2e_steps / e2e_sec

    # extra: the reference's own scheme (explicit adaptive RK45, step control on the device) on the
    # same operator, 40 accepted/rejected attempts from the initial state
    dev.set_state(dev.initial_state())
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev._alloc_history(10 ** 9)                    # no snapshots inside the timed attempts
    from thermca_b200 import _cabi as _cb
    _cb.check(dev.lib.tc_rk_begin(dev._handle, 0, 0.0, 1.0e9, 1e-3, 1e-6, 0.0, 0.0))
    _cb.check(dev.lib.tc_rk_advance(dev._handle, 8, 8))
    sync()
    r0.record(dev.stream)
    _cb.check(dev.lib.tc_rk_advance(dev._handle, 40, 40))
    r1.record(dev.stream)
    sync()
    rk_ms = r0.elapsed_time(r1) / 40
    rk_info = dev.info()
    rk45 = {"ms_per_attempt": rk_ms, "dof_attempts_per_s": n / (rk_ms * 1e-3), "rhs_per_attempt": 6,
            "accepted": rk_info["accepted"], "rejected": rk_info["rejected"], "t_reached_s": rk_info["t"],
            "algorithmic_gbs": (6 * (nnz * 12 + n * 20) + n * 8 * 44) / (rk_ms * 1e-3) / 1e9}

    cpu = cpu_theta_baseline(args.cpu_dof, 2, 1, os.cpu_count() or 1) if not args.no_cpu else None
    line = {
        "metric": "DOF*steps/s (implicit theta-PCG step, full FE cuboid)", "value": value, "unit": "DOF*steps/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg#2 full FE cuboid {nx}x{ny}x{nz} cells = {n} DOF (metric's 10M-DOF size), "
                               "theta=0.5 Jacobi-PCG step", "dof": n, "nnz": nnz, "pcg_rtol": PCG_RTOL,
                   "h_step_s": H_STEP, "iters_per_step": iters / args.steps,
                   "l2": "inputs (1.8 GB operator + 0.5 GB vectors) exceed the 126 MB L2",
                   "pcg": "cooperative persistent kernel" if args.pcg_mode == 0 else "one launch per phase"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_val, "unit": "DOF*steps/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
                "steps": e2e_steps},
        "gpu_launches": tm["launches"],
        "rk45_explicit": rk45,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(n, iters_per_solve), "kernel": "tc_pcg_coop_kernel",
                     "algorithmic_bytes_per_launch": bytes_per_launch, "peak_kind": peak_kind,
                     "ms_per_launch": pcg_ms_per_launch, "iters_per_launch": iters_per_solve,
                     "algorithmic_bytes_per_row_iter": algorithmic_bytes_per_iteration(n, nnz) / n,
                     "share_of_step": tm["pcg_ms"] / ms, "phase_ms_block0": phase},
    }
    if cpu is not None:
        line["cpu_baseline"] = {
            "value": cpu["value"], "unit": "DOF*steps/s", "cores": cpu["cores"], "kind": "port",
            "sample": f"{cpu['dof']} DOF cuboid (same generator), 2 theta-PCG steps, {cpu['backend']} CSR SpMV, "
                      f"{cpu['iters_per_step']:.0f} iters/step", "all_backends": cpu["all_backends"]}
    dev.close()
    print(json.dumps(line))


# ---- secondary metric: batched MOR ensemble (BASELINE config #3) ---------------------------------
def run_mor(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from thermca_b200.mor_batch import MorEnsemble
    S, r, F, B = 3, 200, 2, 4096                      # per GPU; scenario slices shard without communication
    rng = np.random.default_rng(0)                    # SURVEY.md cfg#3: h0 ~ U[5,50], tau ~ U[60,600] s

    def spd(scale):
        Q = rng.standard_normal((r, r)) / np.sqrt(r)
        return (Q @ Q.T + 0.1 * np.eye(r)) * scale
    A_body = np.array([spd(1e-2) for _ in range(S)])
    A_films = np.array([[spd(1e-5) for _ in range(S)] for _ in range(F)])
    B_T = rng.standard_normal((S, r)) * 1e-3
    rng_b = np.random.default_rng(1000 + rank)
    h0, tau = rng_b.uniform(5, 50, (F, B)), rng_b.uniform(60, 600, (F, B))
    ens = MorEnsemble(A_body, A_films, B_T, np.array([0, 1], dtype=np.int32), h0, tau, np.array([-5.0, 3.0]),
                      np.array([100.0, 0.0, 0.0]), np.zeros((S, r)))
    h = 0.5
    # FP64 tensor-core ceiling measured here with a plain library DGEMM (not in MEASURED_PEAKS.json)
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(a, a)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); g0.record()
    for _ in range(5):
        torch.matmul(a, a)
    g1.record(); torch.cuda.synchronize()
    dgemm_tflops = 5 * 2 * 4096 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
    del a
    res = {}
    for dmma in (2, 1, 0):
        ens.steps(args.warmup * 4, h, use_dmma=dmma)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clocks:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0.record()
            ens.steps(args.steps * 4, h, use_dmma=dmma)
            e1.record()
            torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[dmma] = (float(ms.item()), clocks.summary())
    nsteps = args.steps * 4
    ms, clk = res[2]
    tflops = ens.flops_per_step() * nsteps / (ms * 1e-3) / 1e12
    line = {
        "metric": "scenarios*steps/s (batched MOR ensemble, RK4 step)", "value": B * world * nsteps / (ms * 1e-3),
        "unit": "scenarios*steps/s", "n_gpus": world, "steps": nsteps, "warmup": args.warmup * 4,
        "ms_per_step": ms / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"cfg#3 MOR ensemble S={S} r={r} F={F} B={B}/GPU, time-varying films, RK4 h={h}s",
                   "l2": "operators (8.6 MB) are L2-resident by design; state 3x19.7 MB/GPU streams"},
        "clocks": clk, "gpu_launches": nsteps * 5,     # film table + 4 DMMA stage kernels
        "roofline": {"bound": "tensor", "achieved": tflops, "peak": dgemm_tflops, "unit": "TFLOP/s",
                     "frac": tflops / dgemm_tflops, "traffic": None, "kernel": "tc_mor_stage_pipe_kernel<3>",
                     "peak_kind": "torch.matmul fp64 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
                     "simple_dmma_ms_per_step": res[1][0] / nsteps, "fma_variant_ms_per_step": res[0][0] / nsteps},
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = mor_cpu_baseline(A_body, A_films, B_T, h0, tau, h)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def mor_cpu_baseline(A_body, A_films, B_T, h0, tau, h, n_scen=48, n_steps=4):
    """What the reference does for this workload, one scenario at a time (thermca/solver.py:282-299 with
    mor_io.py:90-92): A = A_body + sum_f film_f A_films per right-hand side, then S small GEMVs; RK4, numpy, 1 core."""
    S, r = A_body.shape[:2]
    F = A_films.shape[0]
    film_surf = np.array([0, 1])
    t_link = np.array([-5.0, 3.0])
    q_surf = np.array([100.0, 0.0, 0.0])

    def rhs(t, x, b):
        films = h0[:, b] * (1.0 + 0.5 * np.sin(2.0 * np.pi * t / tau[:, b]))
        A = A_body + sum(f * Af for f, Af in zip(films, A_films))            # (S, r, r), rebuilt per call
        flux = q_surf.copy()
        np.add.at(flux, film_surf, films * t_link)
        return np.array([-a @ xs for a, xs in zip(A, x)]) + B_T * flux[:, None]
    t0 = time.perf_counter()
    for b in range(n_scen):
        x = np.zeros((S, r))
        t = 0.0
        for _ in range(n_steps):
            k1 = rhs(t, x, b); k2 = rhs(t + 0.5 * h, x + 0.5 * h * k1, b)
            k3 = rhs(t + 0.5 * h, x + 0.5 * h * k2, b); k4 = rhs(t + h, x + h * k3, b)
            x = x + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            t += h
    dt = time.perf_counter() - t0
    return {"value": n_scen * n_steps / dt, "unit": "scenarios*steps/s", "cores": 1, "kind": "port",
            "sample": f"{n_scen} scenarios x {n_steps} RK4 steps, numpy, one scenario at a time as the reference runs them"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dof", type=float, default=10.0e6, help="target DOF per GPU")
    ap.add_argument("--cpu-dof", type=float, default=0.5e6, help="DOF of the bounded CPU sample")
    ap.add_argument("--pcg-mode", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--workload", default="fe", choices=["fe", "mor"])
    args = ap.parse_args()
    args.dof, args.cpu_dof = int(args.dof), int(args.cpu_dof)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "mor":
        run_mor(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
