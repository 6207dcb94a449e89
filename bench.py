#!/usr/bin/env python
"""Headline benchmark: DOF*steps/s of the implicit transient step on a synthetic FE cuboid.

A "step" is one implicit theta step (theta = 1/2) of the stacked thermal system: one
right-hand-side evaluation (surface means -> fused network kernel -> SELL SpMV -> surface
loads) and one Jacobi-PCG solve of (M + theta*h*K) — north_star items (a)+(b)+(c).
Workload: the mesh BASELINE.json's metric is quoted on — one tetrahedral FE cuboid, ~10 M DOF,
1 kW on the bottom face, 10 W/m2K film on the other faces, operator generated in HBM by
csrc/synth.cu — on 1 GPU, and the SAME mesh row-partitioned over N GPUs for N > 1 (strong scaling).

  python bench.py --gpus 1 --steps 8 --warmup 3             (our arm)
  python bench.py --impl reference --steps 3 --warmup 1     (CPU reference arm)
  torchrun ... bench.py --gpus N ...                          (row-partitioned, N > 1)

ONE JSON line.  Besides the contract keys it carries
  parity            the timed kernel checked against the CPU oracle (N = 1: oracle/c/theta_cg_omp.c on the bounded
                    sample mesh at the benchmarked pcg_rtol; N > 1: the single-GPU kernel on the same mesh + global sums)
  extra.cfg5_20m    BASELINE config #5: the fixed 20 M-DOF mesh at this N (strong scaling series)
  extra.cfg2_1m     BASELINE config #2 literally (~1 M DOF) at N = 1
  extra.weak        N > 1: 10 M DOF per GPU (round-1 series)
  extra.mor_ensemble  the second half of the metric: scenarios*steps/s of the batched MOR ensemble (config #3),
                    4096 scenarios per GPU, sharded over N without communication
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
METRIC = "DOF*steps/s (implicit theta-PCG step, full FE cuboid)"


def algorithmic_bytes_per_iteration(n_rows, nnz):
    """Jacobi-PCG iteration on CSR-equivalent storage:
    SpMV  nnz*(8+4) + rows*(4 indptr + 8 p gather + 8 mass + 8 q write)
    update x,r (+dots): read x,p,r,q,dinv + write x,r = 56 B/row
    direction        : read r,dinv,p + write p       = 32 B/row"""
    return nnz * 12 + n_rows * (28 + 56 + 32)


def streamed_bytes_per_iteration(n_rows, nnz_stored, col16):
    """What the kernel actually streams: SELL storage incl. padding, 2-byte relative columns when available; the
    x update rides on the direction phase (p is read twice per iteration, not three times): 48 + 48 - 8 B/row."""
    return nnz_stored * (8 + (2 if col16 else 4)) + n_rows * (24 + 48 + 32)


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
        # "under load": the samples of the busier half (idle samples before/after the region do not count)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "sm_mhz_upper_half_median": float(np.median(load)) if load else None}


def measured_traffic(n, iters_per_launch):
    """DRAM bytes per PCG launch from the committed ncu capture of the timed kernel (profiles/r02_pcg_traffic.json,
    falling back on round 1's), scaled to this run's iteration count; None when the capture was taken on a different
    problem size."""
    for name in ("r02_pcg_traffic.json", "r01_pcg_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)
            if abs(t["dof"] - n) > 0.02 * n:
                continue
            return t["dram_bytes_per_row_iter"] * n * iters_per_launch, name
        except Exception:
            continue
    return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def guarded(fn, *a, **k):
    """Extras never take the headline down with them."""
    try:
        return fn(*a, **k)
    except Exception as e:                                   # pragma: no cover
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


# ---- CPU legs (oracle) ---------------------------------------------------------------------------
_CPU_PROBLEM = {}


def _cpu_problem(target_dof):
    if target_dof not in _CPU_PROBLEM:
        _CPU_PROBLEM.clear()
        _CPU_PROBLEM[target_dof] = _build_cpu_problem(target_dof)
    return _CPU_PROBLEM[target_dof]


def _build_cpu_problem(target_dof):
    from thermca_b200.synth import cuboid_spec
    from thermca_b200.workloads import cuboid_dims
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
    return spec, A, mass, b, (nx, ny, nz)


def cpu_theta_baseline(target_dof, steps, warmup, threads):
    """The like-for-like CPU port (oracle/theta_oracle.py + oracle/c/theta_cg_omp.c) of the same step on a bounded
    sample of the workload: same generator, same physics, smaller mesh."""
    from theta_oracle import theta_cg_steps
    spec, A, mass, b, dims = _cpu_problem(target_dof)
    n = A.shape[0]
    y = np.full(n, 20.0)
    runs = {}
    for backend, thr in (("c_omp", threads), ("scipy", 1)):
        try:
            yw, _, _ = theta_cg_steps(A, mass, b, y, warmup, H_STEP, THETA, PCG_RTOL, backend=backend, threads=thr)
            _, iters, sec = theta_cg_steps(A, mass, b, yw, steps, H_STEP, THETA, PCG_RTOL, backend=backend, threads=thr)
        except Exception:
            continue
        runs[backend] = {"value": n * steps / sec, "cores": thr, "backend": backend, "iters_per_step": iters / steps,
                         "ms_per_step": 1e3 * sec / steps}
    # headline CPU figure: all host threads (C/OpenMP port; stands in for the reference's optional MKL backend,
    # thermca/_utils/sparse.py:6-9); the reference's actual fallback path (scipy csr_matvec, one thread) beside it
    best = dict(max(runs.values(), key=lambda r: r["value"]))
    best["all_backends"] = runs
    # the reference's own scheme: explicit adaptive RK45 (scipy) around the restated right-hand side
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
    best.update({"dof": n, "nnz": int(A.nnz), "dims": list(dims)})
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = cpu_theta_baseline(args.cpu_dof, args.steps, max(args.warmup, 1), threads)
    sample = (f"{r['dof']} DOF Kuhn cuboid ({r['dims'][0]}x{r['dims'][1]}x{r['dims'][2]} cells, same generator), "
              f"{args.steps} theta-PCG steps, {r['backend']} CSR SpMV, {r['cores']} threads")
    line = {
        "impl": "reference", "metric": METRIC,
        "value": r["value"], "unit": "DOF*steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "full FE cuboid, theta=0.5 Jacobi-PCG step (CPU port of the same algorithm, bounded "
                               "sample of the 10M-DOF mesh; the reference itself has no implicit scheme)",
                   "dof": r["dof"], "pcg_rtol": PCG_RTOL, "h_step_s": H_STEP,
                   "iters_per_step": r["iters_per_step"], "rk45_as_shipped": r.get("rk45_as_shipped")},
        "cpu_baseline": {"value": r["value"], "unit": "DOF*steps/s", "cores": r["cores"], "kind": "port",
                         "sample": sample, "all_backends": r["all_backends"]},
        "e2e": {"value": r["value"], "unit": "DOF*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def parity_vs_omp(target_dof, steps, threads):
    """Device theta steps against the CPU oracle (oracle/c/theta_cg_omp.c) on the bounded sample mesh (default: the
    ~1 M-DOF cuboid of BASELINE config #2), both at the BENCHMARKED pcg_rtol, from the same initial state; bar 1e-9
    relative (north_star, fixed step)."""
    import torch
    from theta_oracle import theta_cg_steps
    from thermca_b200.device import DeviceSolver
    spec, A, mass, b, dims = _cpu_problem(target_dof)
    n = A.shape[0]
    y_cpu, it_cpu, _ = theta_cg_steps(A, mass, b, np.full(n, 20.0), steps, H_STEP, THETA, PCG_RTOL, backend="c_omp",
                                      threads=threads)
    dev = DeviceSolver(spec)
    dev.theta_steps(steps, H_STEP, THETA, t0=0.0, pcg_rtol=PCG_RTOL)
    y_dev = dev.get_state()[int(spec["u"]):]
    it_dev = dev.pcg_stats()["iterations"]
    dev.close()
    torch.cuda.empty_cache()
    rel = float(np.max(np.abs(y_dev - y_cpu) / np.abs(y_cpu)))
    return {"vs": "theta_cg_omp (oracle/c/theta_cg_omp.c)", "dof": n, "steps": steps, "pcg_rtol": PCG_RTOL,
            "max_rel": rel, "bar": 1e-9, "ok": bool(rel <= 1e-9), "iters_device": it_dev, "iters_cpu": it_cpu,
            "temp_rise_max_K": float(np.max(y_cpu) - 20.0)}


# ---- single GPU ----------------------------------------------------------------------------------
def fe_single(dof, steps, warmup, pcg_mode, local_rank, want_e2e=False, want_rk=False):
    """Time `steps` theta steps of the cuboid with ~dof vertices on this GPU (state resident in HBM)."""
    import torch
    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import cuboid_dims, device_cuboid_spec
    nx, ny, nz = cuboid_dims(dof)
    spec = device_cuboid_spec(nx, ny, nz, heat=HEAT, film=FILM, env_temp=T_ENV)
    dev = DeviceSolver(spec, use_col16=os.environ.get("TC_COL16", "1") != "0")
    n = dev.n_state
    op = spec["ffe"][0]["sell"]
    nnz = int(torch.count_nonzero(op["val"]).item())
    nnz_stored = int(op["val"].numel())
    col16 = op.get("col16") is not None and os.environ.get("TC_COL16", "1") != "0"
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        dev.theta_steps(warmup, H_STEP, THETA, t0=0.0, pcg_rtol=PCG_RTOL, pcg_mode=pcg_mode)
        torch.cuda.synchronize()
        dev.timing(True)
        it0 = dev.pcg_stats()["iterations"]
        torch.cuda.synchronize()
        with torch.cuda.stream(dev.stream):
            e0.record(dev.stream)
            dev.theta_steps(steps, H_STEP, THETA, pcg_rtol=PCG_RTOL, pcg_mode=pcg_mode, sync=False)
            e1.record(dev.stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tm = dev.timing(False)
    phase = dev.pcg_profile()
    bad = dev.pcg_status()
    iters = dev.pcg_stats()["iterations"] - it0
    peak, peak_kind = measured_peak()
    pcg_ms_per_launch = tm["pcg_ms"] / max(tm["pcg_launches"], 1)
    iters_per_solve = iters / max(tm["pcg_launches"], 1)
    bytes_per_launch = algorithmic_bytes_per_iteration(n, nnz) * iters_per_solve + n * 48   # + init phase
    streamed = streamed_bytes_per_iteration(n, nnz_stored, col16) * iters_per_solve + n * 48
    achieved = bytes_per_launch / (pcg_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = measured_traffic(n, iters_per_solve)
    out = {
        "value": n * steps / (ms * 1e-3), "ms_per_step": ms / steps, "dof": n, "nnz": nnz, "dims": [nx, ny, nz],
        "iters_per_step": iters / steps, "gpu_launches": tm["launches"], "clocks": clocks.summary(),
        "pcg_failed_solves": bad["failed_solves"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "frac_dram": (traffic / (pcg_ms_per_launch * 1e-3) / 1e9 / peak) if traffic else None,
                     "frac_streamed": streamed / (pcg_ms_per_launch * 1e-3) / 1e9 / peak,
                     "kernel": "tc_pcg_coop_kernel", "algorithmic_bytes_per_launch": bytes_per_launch,
                     "streamed_bytes_per_launch": streamed, "peak_kind": peak_kind,
                     "ms_per_launch": pcg_ms_per_launch, "iters_per_launch": iters_per_solve,
                     "algorithmic_bytes_per_row_iter": algorithmic_bytes_per_iteration(n, nnz) / n,
                     "share_of_step": tm["pcg_ms"] / ms, "phase_ms_block0": phase},
    }
    if want_e2e:
        # end to end through the host-buffer API: every job = pinned host state up, one step, state down; uploads and
        # downloads double buffered against the steps (DeviceSolver.theta_stream)
        y_dev = torch.from_numpy(dev.get_state())
        jobs = max(6, min(steps, 12))
        host_in = [y_dev.clone().pin_memory() for _ in range(2)]
        host_out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
        dev.theta_stream(host_in, host_out, H_STEP, THETA, pcg_rtol=PCG_RTOL)          # warm-up (allocates staging)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.theta_stream([host_in[j & 1] for j in range(jobs)], [host_out[j & 1] for j in range(jobs)], H_STEP, THETA,
                         pcg_rtol=PCG_RTOL)
        sec = time.perf_counter() - t0
        out["e2e"] = {"value": n * jobs / sec, "unit": "DOF*steps/s", "h2d_bytes_per_step": n * 8,
                      "d2h_bytes_per_step": n * 8, "steps": jobs,
                      "how": "pinned host state -> device, one theta step, state -> pinned host, per job; copies of "
                             "neighbouring jobs overlap the step (double buffered, 3 streams)"}
    if want_rk:
        # the reference's own scheme (explicit adaptive RK45, step control on the device) on the same operator
        from thermca_b200 import _cabi as _cb
        dev.set_state(dev.initial_state())
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dev._alloc_history(10 ** 9)                    # no snapshots inside the timed attempts
        _cb.check(dev.lib.tc_rk_begin(dev._handle, 0, 0.0, 1.0e9, 1e-3, 1e-6, 0.0, 0.0))
        _cb.check(dev.lib.tc_rk_advance(dev._handle, 8, 8))
        torch.cuda.synchronize()
        r0.record(dev.stream)
        _cb.check(dev.lib.tc_rk_advance(dev._handle, 40, 40))
        r1.record(dev.stream)
        torch.cuda.synchronize()
        rk_ms = r0.elapsed_time(r1) / 40
        rk_info = dev.info()
        out["rk45_explicit"] = {"ms_per_attempt": rk_ms, "dof_attempts_per_s": n / (rk_ms * 1e-3), "rhs_per_attempt": 6,
                                "accepted": rk_info["accepted"], "rejected": rk_info["rejected"],
                                "t_reached_s": rk_info["t"],
                                "algorithmic_gbs": (6 * (nnz * 12 + n * 20) + n * 8 * 44) / (rk_ms * 1e-3) / 1e9}
    dev.close()
    del dev, spec, op
    torch.cuda.empty_cache()
    return out


def run_ours_single(args, local_rank):
    import torch
    main = fe_single(args.dof, args.steps, args.warmup, args.pcg_mode, local_rank, want_e2e=True, want_rk=True)
    nx, ny, nz = main["dims"]
    n = main["dof"]
    line = {
        "metric": METRIC, "value": main["value"], "unit": "DOF*steps/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"full FE cuboid {nx}x{ny}x{nz} cells = {n} DOF (the metric's 10M-DOF mesh; the same mesh is "
                               "row-partitioned for N > 1), theta=0.5 Jacobi-PCG step", "dof": n, "nnz": main["nnz"],
                   "pcg_rtol": PCG_RTOL, "h_step_s": H_STEP, "iters_per_step": main["iters_per_step"],
                   "l2": "inputs (1.8 GB operator + 0.5 GB vectors) exceed the 126 MB L2",
                   "pcg": "cooperative persistent kernel, reductions folded into the grid barriers"
                          if args.pcg_mode == 0 else "one launch per phase"},
        "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": main["gpu_launches"],
        "rk45_explicit": main.get("rk45_explicit"), "roofline": main["roofline"],
        "pcg_failed_solves": main["pcg_failed_solves"],
    }
    threads = os.cpu_count() or 1
    extra = {}
    if not args.no_extra:
        line["parity"] = guarded(parity_vs_omp, args.cpu_dof, 3, threads)
        for key, dof in (("cfg5_20m", args.dof20), ("cfg2_1m", 1.0e6)):
            r = guarded(fe_single, int(dof), max(3, args.steps // 2), 3, args.pcg_mode, local_rank)
            if "error" not in r:
                r = {"value": r["value"], "unit": "DOF*steps/s", "ms_per_step": r["ms_per_step"], "dof": r["dof"],
                     "iters_per_step": r["iters_per_step"], "dims": r["dims"], "n_gpus": 1,
                     "roofline_frac": r["roofline"]["frac"], "roofline_frac_streamed": r["roofline"]["frac_streamed"],
                     "achieved_gbs": r["roofline"]["achieved"], "clocks": r["clocks"]}
            extra[key] = r
        extra["mor_ensemble"] = guarded(mor_ensemble, args, 1, 0, local_rank)
    if not args.no_cpu:
        cpu = guarded(cpu_theta_baseline, args.cpu_dof, 2, 1, threads)
        if "error" not in cpu:
            line["cpu_baseline"] = {
                "value": cpu["value"], "unit": "DOF*steps/s", "cores": cpu["cores"], "kind": "port",
                "sample": f"{cpu['dof']} DOF cuboid (same generator), 2 theta-PCG steps, {cpu['backend']} CSR SpMV on "
                          f"{cpu['cores']} threads, {cpu['iters_per_step']:.0f} iters/step",
                "all_backends": cpu["all_backends"], "rk45_as_shipped": cpu.get("rk45_as_shipped")}
        else:
            line["cpu_baseline"] = cpu
    line["extra"] = extra
    print(json.dumps(line))


# ---- N > 1: row-partitioned ------------------------------------------------------------------------
def fe_partitioned(dims, ly, heat, steps, warmup, local_rank, parity_steps=0, want_e2e=False):
    """Time theta steps of the cuboid `dims` row-partitioned over all ranks (csrc/dist.cu)."""
    import torch
    import torch.distributed as dist
    from thermca_b200.dist import DistCuboid
    world, rank = dist.get_world_size(), dist.get_rank()
    nx, ny, nz = dims
    dc = DistCuboid(nx, ny, nz, ly=ly, heat=heat, film=FILM, env_temp=T_ENV)
    out = {}
    if parity_steps:
        # the partitioned kernel against the single-GPU kernel on the SAME global mesh (it fits one GPU), plus two
        # global sums (capacity-weighted temperature sum = thermal energy / (rho c), and the 2-norm)
        from thermca_b200.device import DeviceSolver
        from thermca_b200.workloads import device_cuboid_spec
        dc.theta_steps(parity_steps, H_STEP, THETA, PCG_RTOL, sync=True)
        y_loc = dc.get_state()
        e_d, n_d = dc.global_sums(y_loc)
        spec = device_cuboid_spec(nx, ny, nz, ly=ly, heat=heat, film=FILM, env_temp=T_ENV)
        ref = DeviceSolver(spec)
        ref.theta_steps(parity_steps, H_STEP, THETA, t0=0.0, pcg_rtol=PCG_RTOL)
        y_ref = torch.empty(ref.n_state, dtype=torch.float64, device="cuda")
        ref.get_state_into(y_ref)
        mass = spec["ffe"][0]["sell"]["mass"]
        e_s, n_s = float((mass * y_ref).sum().item()), float((y_ref * y_ref).sum().item())
        sl = y_ref[dc.row_begin: dc.row_begin + dc.n]
        rel = torch.max(torch.abs(y_loc - sl) / torch.abs(sl)).reshape(1)
        dist.all_reduce(rel, op=dist.ReduceOp.MAX)
        its_single = ref.pcg_stats()["iterations"]
        ref.close()
        del ref, spec, y_ref, mass, sl
        torch.cuda.empty_cache()
        out["parity"] = {"vs": "single-GPU kernel on the same mesh (tc_pcg_coop_kernel)", "dof": dc.n_global,
                         "steps": parity_steps, "pcg_rtol": PCG_RTOL, "max_rel": float(rel.item()), "bar": 1e-9,
                         "ok": bool(float(rel.item()) <= 1e-9), "iters_partitioned": dc.stats()["iters_total"],
                         "iters_single_gpu": its_single,
                         "energy_sum_M_T": {"partitioned": e_d, "single_gpu": e_s, "rel_diff": abs(e_d - e_s) / abs(e_s)},
                         "norm2_sq": {"partitioned": n_d, "single_gpu": n_s, "rel_diff": abs(n_d - n_s) / abs(n_s)}}
        dc.set_state(torch.full((dc.n,), 20.0, dtype=torch.float64, device="cuda"))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        dc.theta_steps(warmup, H_STEP, THETA, PCG_RTOL)
        it0 = dc.stats()["iters_total"]
        dc.profile()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(dc.stream)
        dc.theta_steps(steps, H_STEP, THETA, PCG_RTOL)
        e1.record(dc.stream)
        torch.cuda.synchronize()
        dist.barrier()
    dc.check()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    st = dc.stats()
    iters = st["iters_total"] - it0
    prof = dc.profile()
    nnz = torch.tensor([dc.nnz], dtype=torch.int64, device="cuda")
    dist.all_reduce(nnz)
    n_tot = dc.n_global
    pl = prof["last"]
    busy = pl["spmv"] + pl["update x,r"] + pl["direction + halo push"] + pl["dot"]
    wait = pl["reduce-barrier p.q (+allreduce)"] + pl["reduce-barrier r.z,r.r (+allreduce)"] + pl["barrier"]
    mine = torch.tensor([busy, wait], dtype=torch.float64, device="cuda")
    allp = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(allp, mine)
    peak, peak_kind = measured_peak()
    bytes_step = algorithmic_bytes_per_iteration(dc.n, dc.nnz) * (iters / steps) + dc.n * (12 * 15 + 80)
    achieved = bytes_step / (ms / steps * 1e-3) / 1e9
    out.update({
        "value": n_tot * steps / (ms * 1e-3), "ms_per_step": ms / steps, "dof": n_tot, "dof_per_gpu": dc.n,
        "nnz": int(nnz.item()), "iters_per_step": iters / steps, "ghost_rows": dc.n_ghost,
        "boundary_slices": dc.n_bnd, "grid": st["grid"], "late_ghosts": st.get("late_ghosts"), "clocks": clocks.summary(),
        "pcg": dc.pcg,
        "phase_ms_rank0": prof, "busy_wait_ms_per_rank": [[round(float(v[0]), 1), round(float(v[1]), 1)] for v in allp],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "tc_dist_theta_kernel (per GPU)", "peak_kind": peak_kind},
    })
    if want_e2e:
        jobs = max(4, min(steps, 8))
        y0 = dc.get_state().cpu()
        host_in = [y0.clone().pin_memory() for _ in range(2)]
        host_out = [torch.empty(dc.n, dtype=torch.float64).pin_memory() for _ in range(2)]
        dc.theta_stream(host_in, host_out, H_STEP, THETA, PCG_RTOL)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        dc.theta_stream([host_in[j & 1] for j in range(jobs)], [host_out[j & 1] for j in range(jobs)], H_STEP, THETA,
                        PCG_RTOL)
        dist.barrier()
        sec = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": n_tot * jobs / float(sec.item()), "unit": "DOF*steps/s",
                      "h2d_bytes_per_step": n_tot * 8, "d2h_bytes_per_step": n_tot * 8, "steps": jobs,
                      "how": "per rank: pinned host rows -> device, one step, rows -> pinned host, per job; copies of "
                             "neighbouring jobs overlap the step (double buffered)"}
    dc.close()
    del dc
    torch.cuda.empty_cache()
    return out


def run_ours_partitioned(args, world, rank, local_rank):
    from thermca_b200.workloads import cuboid_dims
    dims10 = cuboid_dims(args.dof)
    main = fe_partitioned(dims10, 1.0, HEAT, args.steps, args.warmup, local_rank, parity_steps=3, want_e2e=True)
    nx, ny, nz = dims10
    line = {
        "metric": METRIC, "value": main["value"], "unit": "DOF*steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"full FE cuboid {nx}x{ny}x{nz} cells = {main['dof']} DOF (the metric's 10M-DOF mesh, "
                               f"FIXED size) row-partitioned into {world} y-slabs, theta=0.5 Jacobi-PCG step, ghosts + "
                               "reductions by peer stores inside one cooperative kernel per step",
                   "dof": main["dof"], "dof_per_gpu": main["dof_per_gpu"], "nnz": main["nnz"], "pcg_rtol": PCG_RTOL,
                   "h_step_s": H_STEP, "iters_per_step": main["iters_per_step"], "ghost_rows": main["ghost_rows"],
                   "l2": "per-GPU inputs exceed the 126 MB L2", "grid": main["grid"], "late_ghosts": main.get("late_ghosts"),
                   "pcg": main["pcg"] + " (csrc/dist.cu: classic = two all-reduces + a barrier per iteration, x-update hidden "
                          "behind the second one)",
                   "phase_ms_rank0": main["phase_ms_rank0"], "busy_wait_ms_per_rank": main["busy_wait_ms_per_rank"]},
        "clocks": main["clocks"], "e2e": main["e2e"], "gpu_launches": args.steps * world,
        "roofline": main["roofline"], "parity": main["parity"],
    }
    extra = {}
    if not args.no_extra:
        def slim(r):
            if "error" in r:
                return r
            keep = ("value", "ms_per_step", "dof", "dof_per_gpu", "iters_per_step", "ghost_rows", "clocks", "parity",
                    "busy_wait_ms_per_rank")
            d = {k: r[k] for k in keep if k in r}
            d.update({"unit": "DOF*steps/s", "n_gpus": world, "roofline_frac": r["roofline"]["frac"]})
            return d
        st = max(3, args.steps // 2)
        extra["cfg5_20m"] = slim(guarded(fe_partitioned, cuboid_dims(args.dof20), 1.0, HEAT, st, 3, local_rank,
                                         parity_steps=2))
        extra["cfg5_20m"]["scaling"] = "strong"
        nyw = (ny + 1) * world - 1                     # weak: the same number of y-layers per rank as the 1-GPU mesh
        extra["weak"] = slim(guarded(fe_partitioned, (nx, nyw, nz), 1.0 * world, HEAT * world, st, 3, local_rank))
        extra["weak"]["scaling"] = "weak"
        extra["mor_ensemble"] = guarded(mor_ensemble, args, world, rank, local_rank)
    line["extra"] = extra
    return line


# ---- second half of the metric: batched MOR ensemble (BASELINE config #3) ----------------------------
def reduced_cuboid(r, dof=24000):
    """MOR operators of a REAL body: a ~24k-DOF FE cuboid (3 coupling surfaces: top, sides, bottom), reduced with the
    device Arnoldi build (thermca_b200.mor_basis.transform_device = thermca/fem/mor_io.py:139-220) to r per surface."""
    import scipy.sparse as sp
    from thermca_b200.mor_basis import transform_device
    from thermca_b200.synth import kuhn_cuboid_mesh, p1_operators
    from thermca_b200.workloads import cuboid_dims
    nx, ny, nz = cuboid_dims(dof)
    lx, ly, lz = 0.5, 1.0, 0.05
    pts, tets, tri_bottom, tri_upper = kuhn_cuboid_mesh(nx, ny, nz, lx, ly, lz, jitter=0.15)
    top = np.all(np.abs(pts[tri_upper][:, :, 2] - lz) < 1e-12, axis=1)
    surfs = [tri_upper[top], tri_upper[~top], tri_bottom]
    Lx, Mx, Qs_idx, Qs_val, areas = p1_operators(pts, tets, surfs)
    n = len(Mx)
    M = 8000.0 * 500.0 * Mx
    L = 50.0 * sp.csr_matrix(Lx)
    Qs = np.zeros((n, 3))
    for s in range(3):
        Qs[Qs_idx[s], s] = Qs_val[s]
    C_surf = Qs / Qs.sum(axis=0)
    # thermca/fem/mor_io.py:95-136 (to_mor_state_space), films on surfaces 0 and 1
    M_is = sp.diags(1.0 / np.sqrt(M), format="csr")
    A_body = (M_is @ L @ M_is).tocsr()
    film_idx = np.array([0, 1])
    A_film = [(M_is @ sp.diags(Qs[:, f], format="csr") @ M_is).tocsr() for f in film_idx]
    A = (A_body + sum(A_film)).tocsr()
    B = M_is @ Qs
    (Ab, Af, BT, Cs, VT0, VT1), info = transform_device(M_is, A_body, A_film, A, B, M_is @ C_surf, r, pcg_rtol=1e-12,
                                                        return_info=True)
    return {"A_body": Ab, "A_films": Af, "B_T": BT, "areas": np.asarray(areas), "dof": n, "dims": [nx, ny, nz],
            "film_surf": film_idx.astype(np.int32), "build": info}


def mor_ensemble(args, world, rank, local_rank):
    import torch
    import torch.distributed as dist
    from thermca_b200.mor_batch import MorEnsemble
    S, r, F, B = 3, 200, 2, 4096                      # per GPU; scenario slices shard without communication
    t_build = time.perf_counter()
    red = reduced_cuboid(r)
    t_build = time.perf_counter() - t_build
    A_body, A_films, B_T = red["A_body"], red["A_films"], red["B_T"]
    rng_b = np.random.default_rng(1000 + rank)        # SURVEY.md cfg#3: h0 ~ U[5,50], tau ~ U[60,600] s
    h0, tau = rng_b.uniform(5, 50, (F, B)), rng_b.uniform(60, 600, (F, B))
    t_link = np.array([-5.0, 3.0])
    q_surf = np.array([0.0, 0.0, 100.0]) / red["areas"]
    ens = MorEnsemble(A_body, A_films, B_T, red["film_surf"], h0, tau, t_link, q_surf, np.zeros((S, r)))
    lam = max(float(np.linalg.eigvalsh(A_body[s] + 75.0 * sum(A_films[f][s] for f in range(F)))[-1]) for s in range(S))
    h = min(0.5, 2.0 / lam)                           # inside the RK4 stability interval (2.78 / lambda_max)
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
    nsteps = args.steps * 4
    ens.steps(args.warmup * 4, h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        ens.steps(nsteps, h)
        e1.record()
        torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    X = ens.state()
    finite = bool(np.isfinite(X).all())
    tflops = ens.flops_per_step() * nsteps / (ms * 1e-3) / 1e12
    out = {
        "metric": "scenarios*steps/s (batched MOR ensemble, RK4 step)", "value": B * world * nsteps / (ms * 1e-3),
        "unit": "scenarios*steps/s", "n_gpus": world, "steps": nsteps, "warmup": args.warmup * 4,
        "ms_per_step": ms / nsteps, "scaling": "weak", "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg#3 MOR ensemble S={S} r={r} F={F} B={B}/GPU, time-varying films, RK4 h={h:.3g}s",
                   "operators": f"reduced from a {red['dof']}-DOF FE cuboid by the device Arnoldi build "
                                f"({red['build']['pcg_solves']} PCG solves, {t_build:.1f} s)",
                   "l2": "operators (8.6 MB) are L2-resident by design; state 3x19.7 MB/GPU streams",
                   "state_finite": finite, "lambda_max_h": lam * h},
        "clocks": clocks.summary(), "gpu_launches": ens.launches(nsteps),
        "roofline": {"bound": "tensor", "achieved": tflops, "peak": dgemm_tflops, "unit": "TFLOP/s",
                     "frac": tflops / dgemm_tflops, "traffic": None, "kernel": ens.kernel_name(),
                     "peak_kind": "torch.matmul fp64 4096^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)"},
    }
    if rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = guarded(mor_cpu_baseline, A_body, A_films, B_T, h0, tau, h, t_link, q_surf,
                                      red["film_surf"])
    del ens
    torch.cuda.empty_cache()
    return out


def mor_cpu_baseline(A_body, A_films, B_T, h0, tau, h, t_link, q_surf, film_surf, n_scen=48, n_steps=4):
    """What the reference does for this workload, one scenario at a time (thermca/solver.py:282-299 with
    mor_io.py:90-92): A = A_body + sum_f film_f A_films per right-hand side, then S small GEMVs; RK4, numpy, 1 core."""
    S, r = A_body.shape[:2]

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


def run_mor_only(args, world, rank, local_rank):
    out = mor_ensemble(args, world, rank, local_rank)
    out.update({"higher_is_better": True, "vs_baseline": None})
    if rank == 0:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dof", type=float, default=10.0e6, help="target DOF of the headline mesh")
    ap.add_argument("--dof20", type=float, default=20.0e6, help="target DOF of the cfg#5 mesh (extra)")
    ap.add_argument("--cpu-dof", type=float, default=1.0e6,
                    help="DOF of the bounded CPU sample (default: BASELINE config #2's ~1 M-DOF cuboid)")
    ap.add_argument("--pcg-mode", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline only (profiling runs)")
    ap.add_argument("--workload", default="fe", choices=["fe", "mor"])
    args = ap.parse_args()
    args.dof, args.cpu_dof, args.dof20 = int(args.dof), int(args.cpu_dof), int(args.dof20)
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    force_dist = os.environ.get("TC_FORCE_DIST") == "1"      # developer switch: partitioned path at world 1
    if world > 1 or force_dist:
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "mor":
            run_mor_only(args, world, rank, local_rank)
        elif world > 1 or force_dist:
            line = run_ours_partitioned(args, world, rank, local_rank)
            if rank == 0:
                print(json.dumps(line))
        else:
            run_ours_single(args, local_rank)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
