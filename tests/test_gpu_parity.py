"""Parity of the CUDA path against the oracle and the reference-generated golden vectors.
Everything here goes through the C ABI (thermca_b200/_cabi.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN_MODELS

pytestmark = pytest.mark.gpu
DEVICE_MODELS = list(GOLDEN_MODELS)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def _spmv(torch, lib, m, x, mode=0, b=None, mass=None, shift=0.0, col16=False):
    from thermca_b200 import _cabi
    from thermca_b200.device import csr_to_sell
    s = csr_to_sell(m.indptr, m.indices, m.data)
    dev = torch.device("cuda")
    t = {k: torch.from_numpy(np.ascontiguousarray(s[k])).to(dev) for k in ("slice_ptr", "col", "val")}
    xd = torch.from_numpy(x).to(dev)
    yd = torch.zeros(m.shape[0], dtype=torch.float64, device=dev)
    bd = torch.from_numpy(b).to(dev) if b is not None else None
    md = torch.from_numpy(mass).to(dev) if mass is not None else None
    p = lambda a: C.c_void_p(a.data_ptr()) if a is not None else C.c_void_p(0)
    c16 = None
    if col16 and s["col16"] is not None:
        c16 = torch.from_numpy(s["col16"]).to(dev)
    _cabi.check(lib.tc_sell_spmv(C.c_void_p(torch.cuda.current_stream().cuda_stream), mode, m.shape[0],
                                 s["nslices"], p(t["slice_ptr"]), p(t["col"]), p(c16), p(t["val"]), p(xd), p(yd),
                                 p(bd), p(md), float(shift), 0))
    torch.cuda.synchronize()
    return yd.cpu().numpy()


def test_spmv_bit_exact_vs_scipy(torch_cuda, golden):
    """Row sums in the reference's entry order, no FMA: identical bits to scipy csr_matvec."""
    from thermca_b200 import _cabi
    lib = _cabi.load_library()
    rng = np.random.default_rng(0)
    spec, _ = golden("plate_ffe")
    ab = spec["ffe"][0]["A_body"]
    mats = [sp.csr_matrix((spec["ffe"][0]["A_data"], ab["indices"], ab["indptr"]), shape=tuple(ab["shape"]))]
    mats.append(sp.csr_matrix(sp.random(1000, 1000, density=0.02, random_state=1, format="csr")))
    mats.append(sp.csr_matrix((5, 5)))                       # empty rows
    ragged = sp.lil_matrix((97, 97))
    for i in range(97):
        ragged[i, rng.choice(97, size=1 + (i * 7) % 40, replace=False)] = rng.standard_normal(1 + (i * 7) % 40)
    mats.append(sp.csr_matrix(ragged))
    for m in mats:
        x = rng.standard_normal(m.shape[1])
        y = _spmv(torch_cuda, lib, m, x, mode=0)
        assert np.array_equal(y, -(m @ x))
        y = _spmv(torch_cuda, lib, m, x, mode=0, col16=True)          # 16-bit relative columns: same bits
        assert np.array_equal(y, -(m @ x))
        b = rng.standard_normal(m.shape[0])
        y = _spmv(torch_cuda, lib, m, x, mode=1, b=b)
        assert np.array_equal(y, b - m @ x)
        mass = rng.uniform(1, 2, m.shape[0])
        y = _spmv(torch_cuda, lib, m, x, mode=2, mass=mass, shift=0.37)
        assert np.allclose(y, mass * (x + 0.37 * (m @ x)), rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("name", DEVICE_MODELS)
def test_rhs_matches_reference_and_oracle(torch_cuda, golden, name):
    """f(t, y) on the device == update_time_derivative of the reference (recorded) == oracle."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import OracleRHS
    spec, ex = golden(name)
    dev = DeviceSolver(spec)
    orc = OracleRHS(spec)
    t0, y0 = float(ex["sim"]["t0"]), ex["y0"]
    dev.rhs(t0, y0)
    orc(t0, y0)
    R = ex["rhs"]
    for t, y, fref, href in zip(R["t"], R["y"], R["f"], R["heat"]):
        f = dev.rhs(t, y)
        fo = orc(t, y)
        scale = np.max(np.abs(fref)) + 1e-300
        assert np.max(np.abs(f - fref)) <= 1e-12 * scale, name
        assert np.max(np.abs(f - fo)) <= 1e-12 * scale, name
        st = dev.network_state()
        assert np.allclose(st["heat_src"], href, rtol=1e-12, atol=1e-12)
        assert np.allclose(st["T"], orc.net_temp, rtol=1e-12, atol=1e-12)
        assert np.allclose(st["cond"], orc.cond_data, rtol=1e-9, atol=1e-4)
    dev.close()


@pytest.mark.parametrize("name", DEVICE_MODELS)
def test_adaptive_rk_matches_reference(torch_cuda, golden, name):
    """Same accepted-step sequence and temperatures as Network.sim of the reference
    (1e-9 relative; the step controller runs on the device)."""
    from thermca_b200.device import DeviceSolver
    spec, ex = golden(name)
    sim, ref = ex["sim"], ex["result"]
    dev = DeviceSolver(spec)
    out = dev.run_rk(float(sim["t0"]), float(sim["t1"]), float(sim["rel_tol"]), float(sim["abs_tol"]),
                     str(sim["method"]))
    dev.close()
    assert len(out["times"]) == len(ref["times"]), (len(out["times"]), len(ref["times"]))
    assert np.allclose(out["times"], ref["times"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["net_temps"], ref["net_temps"], rtol=1e-9, atol=1e-9)
    assert np.allclose(out["net_capys"], ref["net_capys"], rtol=1e-9, atol=0, equal_nan=True)
    if ref["io_temps"].size:
        # part states (reduced MOR coordinates span many magnitudes): 1e-9 of the state norm
        io = out["io_temps"][ref["io_snap_idx"]]
        assert np.max(np.abs(io - ref["io_temps"])) <= 1e-9 * np.max(np.abs(ref["io_temps"]))
    assert out["nfev"] == int(ref["nfev"]) + 1


def test_rk_result_skipping_and_zero_span(torch_cuda, golden):
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden("coffeecup")
    dev = DeviceSolver(spec)
    out = dev.run_rk(0.0, 300.0, 1e-5, 1e-6, "RK45", num_skip=3)
    ref = integrate(spec, [0.0, 300.0], 1e-5, 1e-6, "RK45", num_skip=3)
    assert np.allclose(out["times"], ref["times"], rtol=1e-9)
    assert np.allclose(out["net_temps"], ref["net_temps"], rtol=1e-9)
    dev.close()
    dev = DeviceSolver(spec)
    out = dev.run_rk(5.0, 5.0)                 # t1 <= t0: only the initial state (network.py:972-976)
    assert len(out["times"]) == 1 and out["times"][0] == 5.0
    dev.close()


def _theta_reference(spec, nsteps, h, theta):
    """CPU restatement of the theta step with a sparse direct solver (oracle side)."""
    from rhs_oracle import OracleRHS
    from thermca_b200.lowering import initial_state
    import scipy.sparse.linalg as spla
    rhs = OracleRHS(spec)
    y = initial_state(spec)
    t = 0.0
    p = spec["ffe"][0]
    n = int(p["dof"])
    for _ in range(nsteps):
        f = rhs(t + theta * h, y)
        A = rhs.ffe[0]["A"]
        z = spla.spsolve((sp.identity(n, format="csc") + theta * h * A.tocsc()), f)
        y = y + h * z
        t += h
    return y


@pytest.mark.parametrize("mode", [0, 1])
def test_theta_pcg_fixed_step(torch_cuda, golden, mode):
    """Fixed-step theta scheme with Jacobi-PCG == CPU restatement with a direct solver (1e-9)."""
    from thermca_b200.device import DeviceSolver
    spec, _ = golden("kuhn_cuboid")
    dev = DeviceSolver(spec)
    dev.theta_steps(5, 2.0, theta=0.5, t0=0.0, pcg_rtol=1e-13, pcg_maxit=500, pcg_mode=mode)
    y = dev.get_state()
    st = dev.pcg_stats()
    dev.close()
    yref = _theta_reference(spec, 5, 2.0, 0.5)
    assert st["solves"] == 5 and 0 < st["iterations"] < 5 * 500
    assert np.max(np.abs(y - yref) / np.abs(yref)) < 1e-9


def test_ros2_adaptive_within_tolerance(torch_cuda, golden):
    """Adaptive implicit path (ROS2 + PCG, step control on the device) against the oracle's RK45
    run at tight tolerance: the error at the output time stays within a few rtol."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden("plate_ffe")
    t1 = float(ex["sim"]["t1"])
    tight = integrate(spec, [0.0, t1], 1e-10, 1e-12, "RK45")
    errs = []
    for rtol in (1e-3, 1e-4, 1e-5):
        dev = DeviceSolver(spec)
        out = dev.run_ros2(0.0, t1, rtol=rtol, atol=1e-6, pcg_rtol=1e-10)
        dev.close()
        assert out["times"][-1] == t1 and out["status"] == 0
        err = np.max(np.abs(out["io_temps"][-1] - tight["io_temps"][-1]) / np.abs(tight["io_temps"][-1]))
        assert err < 5 * rtol, (rtol, err)
        errs.append(err)
    assert errs[2] < errs[0]


def test_device_generated_cuboid_matches_host_assembly(torch_cuda):
    """csrc/synth.cu (row-wise gather assembly in HBM) == numpy P1 assembly of thermca_b200/synth.py."""
    from thermca_b200.synth import cuboid_spec
    from thermca_b200.workloads import device_cuboid_spec, sell_to_csr_host
    nx, ny, nz = 6, 9, 4
    host = cuboid_spec(nx, ny, nz, jitter=0.15)
    devs = device_cuboid_spec(nx, ny, nz, jitter=0.15)
    hp, dp = host["ffe"][0], devs["ffe"][0]
    n = int(hp["dof"])
    A_host = sp.csr_matrix((hp["A_data"], hp["A_body"]["indices"], hp["A_body"]["indptr"]), shape=(n, n))
    A_dev = sell_to_csr_host(dp["sell"])
    diff = abs(A_host - A_dev)
    assert diff.max() <= 1e-11 * abs(A_host).max()
    assert np.allclose(dp["sell"]["mass"].cpu().numpy(), 1.0 / hp["M_inv"], rtol=1e-12)
    for s in range(2):
        assert np.array_equal(dp["B_idx"][s], hp["B_idx"][s])
        assert np.allclose(dp["B_val"][s], hp["B_val"][s], rtol=1e-11)
        assert np.allclose(dp["C_val"][s], hp["C_val"][s], rtol=1e-11)
    assert np.allclose(dp["surf_areas"], hp["surf_areas"], rtol=1e-12)
    assert np.allclose(dp["surf_areas"], [0.5, 0.5 + 2 * 0.05 * 1.5], rtol=1e-12)


def test_device_generated_cuboid_runs_like_host_spec(torch_cuda):
    from thermca_b200.device import DeviceSolver
    from thermca_b200.synth import cuboid_spec
    from thermca_b200.workloads import device_cuboid_spec
    a = DeviceSolver(cuboid_spec(6, 9, 4, jitter=0.15))
    b = DeviceSolver(device_cuboid_spec(6, 9, 4, jitter=0.15))
    ra = a.run_rk(0.0, 2.0, 1e-6, 1e-9, "RK45")
    rb = b.run_rk(0.0, 2.0, 1e-6, 1e-9, "RK45")
    assert len(ra["times"]) == len(rb["times"])
    assert np.allclose(ra["io_temps"][-1], rb["io_temps"][-1], rtol=1e-10)
    a.close(); b.close()


def test_history_ring_drains_and_options(torch_cuda, golden):
    """Many stored steps through a tiny device history ring (forced drains), plus the scipy
    options the reference forwards (max_step, first_step; thermca/solver.py:362)."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden("coffeecup")
    ref = ex["result"]
    dev = DeviceSolver(spec, history_bytes=12 * 8 * (dev_row := 1 + 3 * int(spec["N"]) + 2 * len(spec["rc_data"]) + 7))
    out = dev.run_rk(0.0, 1200.0, 1e-5, 1e-6, "RK45", attempts_per_poll=4)
    dev.close()
    assert dev.capacity < 20 < len(out["times"])
    assert np.allclose(out["times"], ref["times"], rtol=1e-9)
    assert np.allclose(out["net_temps"], ref["net_temps"], rtol=1e-9)
    dev = DeviceSolver(spec)
    out = dev.run_rk(0.0, 200.0, 1e-4, 1e-6, "RK23", max_step=7.5, first_step=0.25)
    dev.close()
    orc = integrate(spec, [0.0, 200.0], 1e-4, 1e-6, "RK23", max_step=7.5, first_step=0.25)
    assert len(out["times"]) == len(orc["times"]) and np.max(np.diff(out["times"])) <= 7.5 + 1e-12
    assert np.allclose(out["times"], orc["times"], rtol=1e-9)
    assert np.allclose(out["net_temps"], orc["net_temps"], rtol=1e-9)


def test_large_cuboid_properties(torch_cuda):
    """Full-size properties on a 1 M-DOF device-generated operator (no oracle run at this size):
    (1) the theta-PCG solution satisfies the linear system to the requested tolerance (residual
    recomputed with the bit-exact SpMV kernel), (2) energy balance: d/dt sum(M T) equals the net
    heat input, (3) linearity of the SpMV."""
    import torch
    from thermca_b200 import _cabi
    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import cuboid_dims, device_cuboid_spec
    lib = _cabi.load_library()
    nx, ny, nz = cuboid_dims(1.0e6)
    spec = device_cuboid_spec(nx, ny, nz)
    op = spec["ffe"][0]["sell"]
    dev = DeviceSolver(spec)
    n = dev.n_state
    y0 = torch.full((n,), 20.0, dtype=torch.float64, device="cuda")
    f0 = torch.from_numpy(dev.rhs(0.0, y0.cpu().numpy())).cuda()
    # (2) at uniform T = T_env the film exchanges nothing: sum(M f) == heat input (1 kW)
    power = float((op["mass"] * f0).sum())
    assert abs(power - 1000.0) < 1e-6 * 1000.0
    # (1) one theta step: (I + theta h A) z = f0, y1 = y0 + h z
    h, theta = 1.0, 0.5
    dev.theta_steps(1, h, theta, t0=0.0, pcg_rtol=1e-10)
    y1 = torch.from_numpy(dev.get_state()).cuda()
    z = (y1 - y0) / h
    Az = torch.zeros_like(z)
    p = lambda a: C.c_void_p(a.data_ptr()) if a is not None else C.c_void_p(0)
    _cabi.check(lib.tc_sell_spmv(C.c_void_p(torch.cuda.current_stream().cuda_stream), 0, n, op["nslices"],
                                 p(op["slice_ptr"]), p(op["col"]), p(None), p(op["val"]), p(z), p(Az), p(None), p(None),
                                 0.0, 0))
    torch.cuda.synchronize()
    resid = op["mass"] * (z - theta * h * Az - f0)           # Az holds -A z
    rel = float(resid.norm() / (op["mass"] * f0).norm())
    assert rel < 5e-10, rel
    # (3) linearity: A(2x + y) == 2 A x + A y up to rounding
    x1 = torch.rand(n, dtype=torch.float64, device="cuda")
    x2 = torch.rand(n, dtype=torch.float64, device="cuda")
    outs = []
    for v in (x1, x2, 2 * x1 + x2):
        o = torch.zeros_like(v)
        _cabi.check(lib.tc_sell_spmv(C.c_void_p(torch.cuda.current_stream().cuda_stream), 0, n, op["nslices"],
                                     p(op["slice_ptr"]), p(op["col"]), p(op["col16"]), p(op["val"]), p(v), p(o), p(None),
                                     p(None), 0.0, 0))
        torch.cuda.synchronize()
        outs.append(o)
    assert float((outs[2] - 2 * outs[0] - outs[1]).abs().max()) < 1e-9 * float(outs[2].abs().max())
    dev.close()


def test_ros2_with_implicit_lp_part(torch_cuda, golden):
    """ROS2 on the coffee cup (LP part solved implicitly by PCG, network node explicit) against the
    oracle's RK45 at tight tolerance."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden("coffeecup")
    tight = integrate(spec, [0.0, 300.0], 1e-7, 1e-8, "RK45")
    dev = DeviceSolver(spec)
    out = dev.run_ros2(0.0, 300.0, rtol=1e-5, atol=1e-7, pcg_rtol=1e-10)
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == 300.0
    assert np.allclose(out["net_temps"][-1], tight["net_temps"][-1], rtol=2e-4)
    assert np.allclose(out["io_temps"][-1], tight["io_temps"][-1], rtol=2e-4)
    # below the stability bound the W-method keeps the block explicit (TC_W_EXPLICIT_BOUND in csrc/solver.cu); with
    # the bound at 0 every attempt solves the LP part by PCG - same accuracy requirement
    os.environ["TC_W_EXPLICIT_BOUND"] = "0"
    try:
        dev = DeviceSolver(spec)
        imp = dev.run_ros2(0.0, 300.0, rtol=1e-5, atol=1e-7, pcg_rtol=1e-10)
        dev.close()
    finally:
        del os.environ["TC_W_EXPLICIT_BOUND"]
    assert imp["status"] == 0 and imp["pcg"]["solves"] > 0 and out["pcg"]["solves"] == 0
    assert np.allclose(imp["net_temps"][-1], tight["net_temps"][-1], rtol=2e-4)
    assert np.allclose(imp["io_temps"][-1], tight["io_temps"][-1], rtol=2e-4)
    print("coffee cup ROS2 steps: gated", out["accepted"], "always implicit", imp["accepted"])


def test_probe_history_matches_full_history(torch_cuda, golden):
    """Results at scale: the probe ring stores the selected part DOFs of every accepted step; values and step
    sequence equal the full-history run, the final field is fetched once."""
    from thermca_b200.device import DeviceSolver
    from thermca_b200.solver import simulate
    spec, ex = golden("plate_ffe_var")
    sim = ex["sim"]
    args = (float(sim["t0"]), float(sim["t1"]), float(sim["rel_tol"]), float(sim["abs_tol"]), str(sim["method"]))
    dev = DeviceSolver(spec)
    full = dev.run_rk(*args)
    dev.close()
    probes = np.array([0, 7, 511, 934, 3])
    res = simulate(spec, args[:2], args[2], args[3], args[4], probe_dofs=probes)
    assert res["probes"] and res["probe_temps"].shape == (len(full["times"]), len(probes))
    assert np.array_equal(res["times"], full["times"])
    assert np.array_equal(res["net_temps"], full["net_temps"])
    assert np.array_equal(res["probe_temps"], full["io_temps"][:, probes])
    assert np.array_equal(res["final_io_temps"], full["io_temps"][-1])
    # implicit path and thinning
    res2 = simulate(spec, (0.0, 60.0), 1e-4, 1e-6, "ROS2_PCG", num_res_skip=2, probe_dofs=probes[:2])
    assert res2["probe_temps"].shape[1] == 2 and len(res2["times"]) == res2["probe_temps"].shape[0]
    with pytest.raises(IndexError):
        simulate(spec, (0.0, 1.0), probe_dofs=[10 ** 6])
    # an EMPTY probe list is a zero-width probe mode (no part DOFs per step), not full rows
    res0 = simulate(spec, args[:2], args[2], args[3], args[4], probe_dofs=())
    assert res0["probes"] and res0["probe_temps"].shape == (len(full["times"]), 0)
    assert np.array_equal(res0["net_temps"], full["net_temps"])
    assert np.array_equal(res0["final_io_temps"], full["io_temps"][-1])


@pytest.mark.parametrize("name", ["tdep_rod", "assembly"])
def test_ros2_temperature_dependent_body(torch_cuda, golden, name):
    """Temperature dependent FE bodies inside the implicit scheme: the W-method solves them with the operator
    frozen at the initial temperatures (the right-hand side uses the exact k(T), c(T)); result against the
    oracle's RK45 at tight tolerance."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden(name)
    t0, t1 = float(ex["sim"]["t0"]), float(ex["sim"]["t1"])
    tight = integrate(spec, [t0, t1], 1e-9, 1e-11, "RK45")
    dev = DeviceSolver(spec)
    out = dev.run_ros2(t0, t1, rtol=1e-5, atol=1e-8, pcg_rtol=1e-10)
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == t1
    # LP and FE blocks (physical temperatures); reduced MOR coordinates are judged through the network temperatures
    from thermca_b200.spec import state_offsets
    _, _, mor_off, n_state = state_offsets(spec)
    end = (mor_off[0] if len(mor_off) else n_state) - int(spec["u"])
    got, want = out["io_temps"][-1][:end], tight["io_temps"][-1][:end]
    assert np.abs(got - want).max() < 2e-5 * np.abs(want).max()
    assert np.allclose(out["net_temps"][-1], tight["net_temps"][-1], rtol=2e-4, atol=1e-6)
    print(name, "ROS2 accepted", out["accepted"], "rejected", out["rejected"], "pcg", out["pcg"], "RK45 tight steps",
          tight["nsteps"])


def test_ros2_stiff_network_node_is_implicit(torch_cuda, golden):
    """A lumped node with a tiny capacity (time constant ~2 ms against a 20 s simulation) must not restrict the
    implicit scheme: the unknown network nodes are part of the W-method's block solve."""
    import copy
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden("plate_ffe_var")
    spec = copy.deepcopy(spec)
    assert int(spec["u"]) == 1
    spec["capy"] = np.array(spec["capy"], dtype=np.float64)
    spec["capy"][0] = 0.01                                    # was 2e3 J/K; conductances of ~5 W/K stay
    t1 = 20.0
    tight = integrate(spec, [0.0, t1], 1e-8, 1e-10, "RK45")
    dev = DeviceSolver(spec)
    out = dev.run_ros2(0.0, t1, rtol=1e-4, atol=1e-7, pcg_rtol=1e-10)
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == t1
    assert np.allclose(out["net_temps"][-1], tight["net_temps"][-1], rtol=5e-4, atol=1e-4)
    assert out["accepted"] + out["rejected"] < tight["nsteps"] / 10, (out["accepted"], tight["nsteps"])
    print("stiff node: ROS2 steps", out["accepted"], "+", out["rejected"], "explicit RK45 steps", tight["nsteps"])


@pytest.mark.parametrize("name", ["plate_mor", "plate_mor_var"])
def test_ros2_with_implicit_mor_blocks(torch_cuda, golden, name):
    """MOR bodies inside the implicit scheme (dense r x r CG per surface, films of the latest right-hand side)."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden(name)
    t1 = float(ex["sim"]["t1"])
    tight = integrate(spec, [0.0, t1], 1e-9, 1e-11, "RK45")
    dev = DeviceSolver(spec)
    out = dev.run_ros2(0.0, t1, rtol=1e-5, atol=1e-8, pcg_rtol=1e-10)
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == t1
    assert np.allclose(out["net_temps"][-1], tight["net_temps"][-1], rtol=2e-4, atol=1e-6)
    scale = np.abs(tight["io_temps"][-1]).max()
    assert np.abs(out["io_temps"][-1] - tight["io_temps"][-1]).max() < 1e-3 * scale
    print(name, "ROS2 accepted", out["accepted"], "rejected", out["rejected"])


@pytest.mark.parametrize("name,t1", [("getting_started", 5.0), ("plate_ffe_var", 20.0), ("kuhn_cuboid", 5.0), ("lp_tdep", 3.0),
                                     ("steel_sheet", 2000.0), ("cutting_disc", 2.0)])
def test_ros2_agrees_with_tight_rk45_on_every_model_family(torch_cuda, golden, name, t1):
    """The LSODA substitute on the remaining goldens (network-only, LP, temperature dependent LP, FE with
    closures / Inputs / stationary nodes): node temperatures and part states against a tight explicit run."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden(name)
    t0 = float(ex["sim"]["t0"])
    tight = integrate(spec, [t0, t1], 1e-8, 1e-10, "RK45")
    rtol = 1e-5
    dev = DeviceSolver(spec)
    out = dev.run_ros2(t0, t1, rtol=rtol, atol=1e-3 * rtol, pcg_rtol=1e-10)
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == t1
    assert np.allclose(out["net_temps"][-1], tight["net_temps"][-1], rtol=5e-4, atol=1e-5)
    if tight["io_temps"].shape[1]:
        scale = np.abs(tight["io_temps"][-1]).max()
        assert np.abs(out["io_temps"][-1] - tight["io_temps"][-1]).max() < 5e-4 * scale


@pytest.mark.parametrize("name,t1,bound", [("coffeecup", 100.0, 2.0), ("coffeecup", 100.0, 0.0), ("plate_ffe", 30.0, 0.0),
                                           ("plate_mor", 100.0, 0.0), ("assembly", 30.0, 2.0), ("assembly", 30.0, 0.0),
                                           ("plate_ffe_var", 20.0, 0.0), ("tdep_rod", 0.5, 2.0), ("lp_tdep", 0.5, 0.0),
                                           ("lp_tdep", 3.0, 2.0), ("plate_ffe_stat", 30.0, 2.0), ("plate_ffe_statc", 30.0, 2.0),
                                           ("plate_mor_stat", 100.0, 2.0)])
def test_ros2_matches_cpu_restatement(torch_cuda, golden, name, t1, bound):
    """The adaptive ROS2 W-method on the device against its CPU restatement (oracle/ros2_oracle.py: same stages, block
    structure, explicit/implicit switch and controller, direct solves instead of PCG): identical accept/reject
    sequence, temperatures to solver tolerance.  bound = 0 forces every block onto its implicit branch."""
    from thermca_b200.device import DeviceSolver
    from ros2_oracle import ros2_integrate
    spec, ex = golden(name)
    t0 = float(ex["sim"]["t0"])
    ref = ros2_integrate(spec, t0, t1, 1e-4, 1e-7, bound=bound)
    os.environ["TC_W_EXPLICIT_BOUND"] = repr(bound)
    try:
        dev = DeviceSolver(spec)
        out = dev.run_ros2(t0, t1, rtol=1e-4, atol=1e-7, pcg_rtol=1e-13, pcg_maxit=5000)
        dev.close()
    finally:
        del os.environ["TC_W_EXPLICIT_BOUND"]
    assert (out["accepted"], out["rejected"]) == (ref["accepted"], ref["rejected"])
    assert np.allclose(out["times"], ref["times"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["net_temps"], ref["net_temps"], rtol=1e-7, atol=1e-9)
    if ref["io_temps"].shape[1]:
        assert np.abs(out["io_temps"] - ref["io_temps"]).max() <= 1e-7 * np.abs(ref["io_temps"]).max()
    if bound == 0.0 and (spec["lp"] or spec["ffe"]):
        assert out["pcg"]["solves"] > 0


@pytest.mark.parametrize("scheme,kind", [("ROS2", 2), ("ROW3", 3)])
def test_w_method_attempt_graph_is_used_and_changes_nothing(torch_cuda, golden, scheme, kind):
    """A W-method attempt is replayed as ONE CUDA graph (cooperative PCG launches included); the trajectory is
    bit-identical to launching the same kernels one by one (TC_W_GRAPH=0), and a second simulation on the same solver
    re-captures (tolerances and history pointers are baked into the graph)."""
    from thermca_b200.device import DeviceSolver
    spec, ex = golden("assembly")
    t0 = float(ex["sim"]["t0"])
    runs = {}
    for mode in ("1", "0"):
        os.environ["TC_W_GRAPH"] = mode
        try:
            dev = DeviceSolver(spec)
            out = dev.run_ros2(t0, t0 + 20.0, rtol=1e-4, atol=1e-7, scheme=scheme)
            info = dev.graph_info()
            if mode == "1":
                assert info["w"] == kind, info
                again = dev.run_ros2(t0, t0 + 20.0, rtol=1e-3, atol=1e-6, scheme=scheme)    # other tolerances: new graph
                assert dev.graph_info()["w"] == kind and again["accepted"] > 0 and np.isfinite(again["net_temps"]).all()
            else:
                assert info["w"] == 0, info
            dev.close()
        finally:
            del os.environ["TC_W_GRAPH"]
        runs[mode] = out
    a, b = runs["1"], runs["0"]
    assert (a["accepted"], a["rejected"]) == (b["accepted"], b["rejected"])
    assert np.array_equal(a["times"], b["times"]) and np.array_equal(a["net_temps"], b["net_temps"])
    assert np.array_equal(a["io_temps"], b["io_temps"])


@pytest.mark.parametrize("name,t1,bound", [("coffeecup", 100.0, 0.5), ("coffeecup", 100.0, 0.0), ("plate_ffe", 30.0, 0.0),
                                           ("plate_mor", 100.0, 0.0), ("assembly", 30.0, 0.5), ("assembly", 30.0, 0.0),
                                           ("plate_ffe_var", 20.0, 0.0), ("tdep_rod", 0.5, 0.5), ("lp_tdep", 0.5, 0.0),
                                           ("lp_tdep", 3.0, 0.5), ("plate_ffe_stat", 30.0, 0.5), ("plate_ffe_statc", 30.0, 0.5),
                                           ("plate_mor_stat", 300.0, 0.5)])
def test_row3_matches_cpu_restatement(torch_cuda, golden, name, t1, bound):
    """The adaptive third-order W-method on the device against its CPU restatement (oracle/row3_oracle.py: same four
    stages, block structure, explicit/implicit switch and controller, direct solves instead of PCG): identical
    accept/reject sequence, temperatures to solver tolerance.  bound = 0 forces every block onto its implicit branch."""
    from thermca_b200.device import DeviceSolver
    from row3_oracle import row3_integrate
    spec, ex = golden(name)
    t0 = float(ex["sim"]["t0"])
    ref = row3_integrate(spec, t0, t1, 1e-4, 1e-7, bound=bound)
    os.environ["TC_W_EXPLICIT_BOUND"] = repr(bound)
    try:
        dev = DeviceSolver(spec)
        out = dev.run_ros2(t0, t1, rtol=1e-4, atol=1e-7, pcg_rtol=1e-13, pcg_maxit=5000, scheme="ROW3")
        dev.close()
    finally:
        del os.environ["TC_W_EXPLICIT_BOUND"]
    assert (out["accepted"], out["rejected"]) == (ref["accepted"], ref["rejected"])
    assert np.allclose(out["times"], ref["times"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["net_temps"], ref["net_temps"], rtol=1e-7, atol=1e-9)
    if ref["io_temps"].shape[1]:
        assert np.abs(out["io_temps"] - ref["io_temps"]).max() <= 1e-7 * np.abs(ref["io_temps"]).max()
    if bound == 0.0 and (spec["lp"] or spec["ffe"]):
        assert out["pcg"]["solves"] > 0


@pytest.mark.parametrize("name,t1,max_attempts", [("lp_tdep", 3.0, 150), ("plate_ffe_stat", 30.0, 30),
                                                    ("plate_ffe_statc", 30.0, 30), ("plate_mor_stat", 300.0, 250)])
def test_w_method_with_films_to_stationary_nodes(torch_cuda, golden, name, t1, max_attempts):
    """Surfaces fed through a stationary node (StatNode + HeatSource + stiff film: the reference's idiom for a prescribed
    heat flow, thermca/tests/test_lp_part.py:655-660) and films in series with a stationary node.  The node follows the
    surface instantly, so the film is a rank-one term of the Jacobian (tc_set_jac_films); with the film on the block
    diagonal only, the third-order W-method answered these models 7e-4 .. 3e-3 off at rtol 1e-5 (CPU restatement:
    tests/test_oracle_golden.py).  Bar: the LSODA substitute at rtol 1e-5 within 5 rtol of a tight explicit run, in a
    step count of LSODA's order."""
    from thermca_b200.device import DeviceSolver
    from rhs_oracle import integrate
    spec, ex = golden(name)
    t0 = float(ex["sim"]["t0"])
    tight = integrate(spec, [t0, t1], 1e-10, 1e-12, "RK45")
    dev = DeviceSolver(spec)
    out = dev.run_ros2(t0, t1, rtol=1e-5, atol=1e-8, pcg_rtol=1e-12, pcg_maxit=5000, scheme="ROW3")
    dev.close()
    assert out["status"] == 0 and out["times"][-1] == t1
    assert out["accepted"] + out["rejected"] <= max_attempts
    Tt = tight["net_temps"][-1]
    assert np.abs(out["net_temps"][-1] - Tt).max() <= 5e-5 * np.abs(Tt).max()
    it = tight["io_temps"][-1]
    assert np.abs(out["io_temps"][-1] - it).max() <= 5e-5 * np.abs(it).max()


def test_row3_needs_fewer_attempts_than_ros2_at_equal_tolerance(torch_cuda, golden):
    """What the third order buys (the LSODA substitute): attempts and linear solves for the plate body at rtol 1e-5,
    both schemes within 10 * rtol of a tight RK45 run."""
    from thermca_b200.device import DeviceSolver
    spec, ex = golden("plate_ffe")
    t0, t1 = float(ex["sim"]["t0"]), float(ex["sim"]["t1"])
    dev = DeviceSolver(spec)
    tight = dev.run_rk(t0, t1, 1e-10, 1e-12, "RK45")
    dev.close()
    res = {}
    for scheme in ("ROS2", "ROW3"):
        dev = DeviceSolver(spec)
        out = dev.run_ros2(t0, t1, rtol=1e-5, atol=1e-9, pcg_rtol=1e-11, scheme=scheme)
        dev.close()
        err = np.abs(out["io_temps"][-1] - tight["io_temps"][-1]).max() / np.abs(tight["io_temps"][-1]).max()
        res[scheme] = (out["accepted"] + out["rejected"], err)
        assert err < 1e-4
    assert 2 * res["ROW3"][0] < res["ROS2"][0], res          # 4 vs 2 solves per attempt: fewer solves in total


def test_theta_pcg_matches_cpu_oracle_at_the_benchmarked_tolerance(torch_cuda):
    """Fixed theta steps of the device (cooperative PCG kernel, pcg_rtol = 1e-8 as in bench.py) against the CPU oracle
    (oracle/c/theta_cg_omp.c, same algorithm) AND the sparse direct solver on a 200 k-DOF cuboid of the bench generator:
    the north_star's 1e-9 relative at fixed step holds at the benchmarked tolerance, not only at 1e-13 on 364 DOF."""
    import scipy.sparse as sp
    from thermca_b200.device import DeviceSolver
    from thermca_b200.synth import cuboid_spec
    from thermca_b200.workloads import cuboid_dims
    from theta_oracle import theta_cg_steps, theta_step_direct
    nx, ny, nz = cuboid_dims(2.0e5)
    spec = cuboid_spec(nx, ny, nz, jitter=0.15, heat=1000.0, film=10.0, env_temp=20.0)
    p = spec["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    mass = 1.0 / p["M_inv"]
    b = np.zeros(n)
    flux = [1000.0 / p["surf_areas"][0], 10.0 * 20.0]
    for s_ in range(2):
        b[p["B_idx"][s_]] += p["B_val"][s_] * flux[s_]
    steps, h = 3, 1.0
    y_cpu, it_cpu, _ = theta_cg_steps(A, mass, b, np.full(n, 20.0), steps, h, 0.5, 1e-8, backend="c_omp", threads=4)
    y_dir = np.full(n, 20.0)
    for _ in range(steps):
        y_dir = theta_step_direct(A, b - A @ y_dir, y_dir, h, 0.5)
    dev = DeviceSolver(spec)
    dev.theta_steps(steps, h, 0.5, t0=0.0, pcg_rtol=1e-8)
    y_dev = dev.get_state()[int(spec["u"]):]
    it_dev = dev.pcg_stats()["iterations"]
    dev.close()
    assert np.max(np.abs(y_dev - y_cpu) / np.abs(y_cpu)) <= 1e-9
    assert np.max(np.abs(y_dev - y_dir) / np.abs(y_dir)) <= 1e-9
    assert abs(it_dev - it_cpu) <= 2


# ---- block-Jacobi preconditioner (north_star (b)) ------------------------------------------------------------------
def test_block_jacobi_pcg_matches_point_jacobi_and_needs_fewer_iterations():
    """theta steps with the block-Jacobi variant (32 x 32 slice blocks of M + theta*h*K, FP32-stored inverses) give the
    same temperatures as the point-Jacobi path (both solve to pcg_rtol; 1e-9 relative bar) in fewer iterations."""
    from thermca_b200.device import DeviceSolver
    from thermca_b200.workloads import device_cuboid_spec
    out = {}
    for mode in (0, 2):
        dev = DeviceSolver(device_cuboid_spec(40, 80, 6))
        if mode == 2:
            dev.set_block_jacobi(10.0, 0.5)
        dev.theta_steps(4, 10.0, 0.5, t0=0.0, pcg_rtol=1e-12, pcg_mode=mode)
        out[mode] = (dev.get_state(), dev.pcg_stats()["iterations"])
        dev.close()
    y0, it0 = out[0]
    y2, it2 = out[2]
    assert np.max(np.abs(y2 - y0)) <= 1e-9 * np.max(np.abs(y0))
    assert it2 < it0, (it2, it0)


def test_block_jacobi_steady_state_solve():
    import scipy.sparse as sp
    from thermca_b200.stationary import spd_solve
    from thermca_b200.synth import cuboid_spec
    p = cuboid_spec(10, 20, 3, jitter=0.15)["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    K = sp.diags(1.0 / p["M_inv"]) @ A                       # stiffness + film diagonal: symmetric positive definite
    b = np.random.default_rng(0).standard_normal(n)
    xj, ij = spd_solve(K, b, rtol=1e-12, return_info=True)
    xb, ib = spd_solve(K, b, rtol=1e-12, return_info=True, precond="block")
    assert np.max(np.abs(xj - xb)) <= 1e-9 * np.max(np.abs(xj))
    assert ib["iterations"] < ij["iterations"]
