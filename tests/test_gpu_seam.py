"""The seam on hardware: the UNMODIFIED reference ``Network(model).sim(...)`` (thermca/network.py:934-986) on
top of the CUDA kernels (``thermca_b200.install()`` swaps ``thermca.solver.transient_solution``,
thermca/solver.py:61-69) against the same model run through the reference's own scipy/numba solver in the
SAME process.  The reference package is the staged copy ``oracle/_ref`` (oracle/build_ref.py; git-ignored,
shipped with the snapshot) — on the GPU box ``/root/reference`` does not exist.

Bars (north_star): identical accepted-step sequences, times / temperatures within 1e-9 relative for the
explicit pairs the reference itself runs; within the solver's tolerances at the stored times for LSODA
requests (served by the implicit W-method)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refenv  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not refenv.available(), reason="reference copy oracle/_ref not staged")]

MODELS = ["getting_started", "coffeecup", "cutting_disc", "plate_ffe", "plate_mor", "plate_ffe_var", "plate_mor_var",
          "kuhn_cuboid", "tdep_rod", "assembly", "lp_tdep", "steel_sheet", "plate_ffe_stat", "plate_ffe_statc",
          "plate_mor_stat"]


@pytest.fixture()
def seam(tmp_path, monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    refenv.activate()
    monkeypatch.chdir(tmp_path)                    # the reference's disk_cache writes ./.thermca_cache
    import thermca_b200
    yield thermca_b200
    thermca_b200.uninstall()


def _sim(net, kw):
    args = {k: v for k, v in kw.items() if k != "time_span"}
    return net.sim(list(kw["time_span"]), progress_bar=False, **args)


def _both(seam, name, **override):
    import ref_models
    import thermca.solver as ref_solver
    original = ref_solver.transient_solution
    seam.install(stationary=False, assembly=False, mor_basis=False)    # this file pins the transient seam only
    assert ref_solver.transient_solution is not original
    net, kw = ref_models.MODELS[name]()
    kw = dict(kw, **override)
    res = _sim(net, kw)
    seam.uninstall()
    assert ref_solver.transient_solution is original
    net_ref, _ = ref_models.MODELS[name]()
    ref = _sim(net_ref, kw)
    return net, res, net_ref, ref, kw


@pytest.mark.parametrize("name", MODELS)
def test_network_sim_on_device_matches_reference_solver(seam, name):
    net, res, net_ref, ref, kw = _both(seam, name)
    d, r = res._data, ref._data
    assert len(d.times) == len(r.times), "accepted-step sequences differ"
    assert np.allclose(d.times, r.times, rtol=1e-9, atol=0.0)
    scale = np.max(np.abs(r.net_temps))
    assert np.max(np.abs(d.net_temps - r.net_temps)) <= 1e-9 * scale
    if r.io_temps.size:
        assert d.io_temps.shape == r.io_temps.shape and d.io_temp_dt == r.io_temp_dt
        assert np.max(np.abs(d.io_temps - r.io_temps)) <= 1e-9 * max(np.max(np.abs(r.io_temps)), scale)
    # stored per-step network rows (store_results, thermca/solver.py:393-424): capacities, sources, conductances
    assert np.allclose(d.net_capys, r.net_capys, rtol=1e-9, atol=0.0)
    hs = max(np.max(np.abs(r.net_heat_srcs)), 1e-300)
    assert np.max(np.abs(d.net_heat_srcs - r.net_heat_srcs)) <= 1e-8 * hs
    # free-convection films ~|dT|^(1/4) turn 1 ulp of a surface mean at dT~0 into a visible conductance
    assert np.allclose(np.asarray(d.net_clean_conds), np.asarray(r.net_clean_conds), rtol=1e-7, atol=1e-4)
    # (an entry that is exactly 0 — the DOK fill value — is dropped from the key set, so compare over the union)
    kd, kr = d.net_conds.data, r.net_conds.data
    keys = sorted(set(kd) | set(kr))
    a = np.array([kd.get(k, 0.0) for k in keys]); b = np.array([kr.get(k, 0.0) for k in keys])
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    assert np.allclose(a[fin], b[fin], rtol=1e-7, atol=1e-4)
    # per-step LP capacity / conductance matrices the post-processing reads (temperature dependent LP bodies: lp_tdep)
    assert len(d.lp_M_diagss) == len(r.lp_M_diagss)
    for md, mr, ld, lr in zip(d.lp_M_diagss, r.lp_M_diagss, d.lp_Lss, r.lp_Lss):
        for a_, b_ in zip(md, mr):
            assert np.allclose(a_, b_, rtol=1e-9, atol=0.0)
        for a_, b_ in zip(ld, lr):
            assert np.allclose(a_.data, b_.data, rtol=1e-8, atol=1e-12 * np.max(np.abs(b_.data)))
    # what the reference leaves on the Network after the solve
    assert net.time == pytest.approx(net_ref.time, rel=1e-12)
    assert np.allclose(net.capy, net_ref.capy, rtol=1e-9)
    assert np.allclose(net.run_cond.data, net_ref.run_cond.data, rtol=1e-7, atol=1e-4)


def test_coffeecup_published_table_on_device(seam):
    """docs/examples/coffeecup.ipynb:533-542 through the reference's own post-processing, solved on the B200."""
    import ref_models
    seam.install(stationary=False, assembly=False, mor_basis=False)
    net, kw = ref_models.coffeecup()
    res = _sim(net, kw)
    times = np.linspace(0.0, 1200.0, 10)
    model = net.model
    coffee = [e for e in model._get_all_elems_of_types(type(net._node_elems[list(net._node_elems)[0]][0]))
              if getattr(e, "name", "") == "coffee"]
    table = [100.00, 82.03, 78.43, 75.43, 72.76, 70.32, 68.10, 66.06, 64.17, 62.41]
    if coffee:
        got = np.asarray(res[coffee[0]].temp(times)).ravel()
        assert np.allclose(got, table, atol=5.1e-3)
    else:                                            # fall back on the node order of the lowered network
        idx = int(np.argmax(res._data.net_temps[0]))
        got = np.interp(times, res._data.times, res._data.net_temps[:, idx])
        assert np.allclose(got, table, atol=0.05)


def test_result_accessors_on_device_collector(seam):
    """``Result[...]`` views that read conductances, io_temps and the MOR reconstruction (resultdata.py:220-242)."""
    net, res, net_ref, ref, kw = _both(seam, "plate_mor_var")
    part, part_ref = net._fe_parts[0], net_ref._fe_parts[0]
    t_end = float(kw["time_span"][1])
    assert np.allclose(res[part].temp(t_end), ref[part_ref].temp(t_end), rtol=1e-9)
    assert np.allclose(res[part.surf.top].temp([10.0, t_end]), ref[part_ref.surf.top].temp([10.0, t_end]), rtol=1e-9)


def test_lsoda_request_served_by_implicit_path_within_tolerance(seam):
    """The reference's FE tests run LSODA (thermca/tests/test_fe_part.py:13-99,230-303); here the request is served by
    the implicit W-method: temperatures at the stored end time agree within the requested tolerance."""
    rtol = 1e-5
    net, res, net_ref, ref, kw = _both(seam, "plate_ffe", method="LSODA", rel_tol=rtol)
    a, b = res._data.net_temps[-1], ref._data.net_temps[-1]
    assert res._data.times[-1] == pytest.approx(ref._data.times[-1], rel=1e-12)
    assert np.max(np.abs(a - b)) <= 20 * rtol * np.max(np.abs(b))
    fa, fb = res._data.io_temps[-1], ref._data.io_temps[-1]
    assert np.max(np.abs(fa - fb)) <= 20 * rtol * np.max(np.abs(fb))


@pytest.mark.parametrize("name", ["lp_tdep", "plate_ffe_stat", "plate_ffe_statc", "plate_mor_stat"])
def test_lsoda_request_on_models_fed_through_stationary_nodes(seam, name):
    """Surfaces fed through StatNode + HeatSource + a stiff film (the reference's idiom for a prescribed heat flow,
    thermca/tests/test_lp_part.py:655-660) or cooled through a film in series with a stationary node: an LSODA request
    through the unmodified ``Network.sim`` on the device against the reference's own LSODA run.  These are the models
    the rank-one Jacobian terms of tc_set_jac_films exist for."""
    rtol = 1e-5
    net, res, net_ref, ref, kw = _both(seam, name, method="LSODA", rel_tol=rtol, abs_tol=1e-8)
    a, b = res._data.net_temps[-1], ref._data.net_temps[-1]
    assert res._data.times[-1] == pytest.approx(ref._data.times[-1], rel=1e-12)
    assert np.max(np.abs(a - b)) <= 20 * rtol * np.max(np.abs(b))
    fa, fb = res._data.io_temps[-1], ref._data.io_temps[-1]
    assert np.max(np.abs(fa - fb)) <= 20 * rtol * np.max(np.abs(fb))
    assert len(res._data.times) <= 2 * len(ref._data.times) + 20        # step count of LSODA's order


def test_install_everything_full_model_build_on_device(seam):
    """All four swaps at once (transient, stationary, assembly, MOR basis): a model BUILT and SOLVED with the
    device routines agrees with the reference built and solved on the CPU (MOR bases differ by solver rounding
    amplified through the Krylov recurrence, DESIGN §8: compare physical outputs)."""
    import ref_models
    seam.install()
    net, kw = ref_models.plate_mor()
    res = _sim(net, kw)
    seam.uninstall()
    net_ref, _ = ref_models.plate_mor()
    ref = _sim(net_ref, kw)
    t_end = float(kw["time_span"][1])
    part, part_ref = net._fe_parts[0], net_ref._fe_parts[0]
    assert np.allclose(res[part.surf.top].temp(t_end), ref[part_ref.surf.top].temp(t_end), rtol=1e-5)
    assert np.allclose(res[part].temp(t_end), ref[part_ref].temp(t_end), rtol=1e-4, atol=1e-6)


def test_config4_sized_assembly_on_device(seam):
    """BASELINE config #4 at its stated size (oracle/ref_models.py:assembly_large): a 24 k-DOF temperature dependent
    full-order FE body, an 8.6 k-DOF body reduced to MOR, two linked LP blocks, temperature dependent fluid nodes with
    flow links, a stationary node, nonlinear films and an Input driven heat loss — built by the reference, solved on
    the device, against the reference's own RK45."""
    import ref_models
    seam.install(stationary=False, assembly=False, mor_basis=False)
    net, kw = ref_models.assembly_large()
    res = _sim(net, kw)
    seam.uninstall()
    net_ref, _ = ref_models.assembly_large()
    ref = _sim(net_ref, kw)
    d, r = res._data, ref._data
    assert [io.fe_sys.dof for io in net._ffe_ios] == [24309] and [io.fe_sys.dof for io in net._mor_ios] == [8575]
    assert len(d.times) == len(r.times), "accepted-step sequences differ"
    assert np.allclose(d.times, r.times, rtol=1e-9, atol=0.0)
    scale = np.max(np.abs(r.net_temps))
    assert np.max(np.abs(d.net_temps - r.net_temps)) <= 1e-9 * scale
    assert np.max(np.abs(d.io_temps - r.io_temps)) <= 1e-9 * np.max(np.abs(r.io_temps))
    frame, frame_ref = net._fe_parts[0], net_ref._fe_parts[0]
    assert np.allclose(res[frame.surf.top].temp(20.0), ref[frame_ref.surf.top].temp(20.0), rtol=1e-9)


def test_theta_pcg_method_through_the_seam(seam):
    """``Network.sim(..., method='THETA_PCG', step=h)``: the fixed-step implicit scheme (new method name, passes
    through the reference's ``sim`` untouched) against a CPU restatement with a sparse direct solver on the
    reference's own operators (``FFEIO.A``), 1e-9 relative (north_star: fixed step)."""
    import scipy.sparse as sp
    import ref_models
    from theta_oracle import theta_step_direct
    seam.install(stationary=False, assembly=False, mor_basis=False)
    net, kw = ref_models.kuhn_cuboid()
    h, nsteps = 0.5, 8
    res = net.sim([0.0, h * nsteps], method="THETA_PCG", step=h, pcg_rtol=1e-13, progress_bar=False, num_res_skip=1)
    d = res._data
    assert np.allclose(d.times, [0.0, 1.0, 2.0, 3.0, 4.0])
    io = net._ffe_ios[0]
    A = sp.csr_matrix(io.A)
    n = A.shape[0]
    flux = np.array([1000.0 / io.fe_sys.surf_areas[0], 10.0 * 20.0])          # heat on `bottom`, film * T_env on `upper_faces`
    b = np.asarray(io.B @ flux).ravel()
    y = np.full(n, 20.0)
    for _ in range(nsteps):
        y = theta_step_direct(A, b - A @ y, y, h, 0.5)
    got = d.io_temps[-1]
    assert np.max(np.abs(got - y)) <= 1e-9 * np.max(np.abs(y))
