"""Drop-in seam: ``Network.sim`` of the REAL reference routed through
``thermca_b200.solver.transient_solution``.  Runs only where /root/reference exists (the build
container, no GPU there): the device is replaced by an oracle-backed stand-in, so this checks
the lowering of live ``Network`` objects, the ``_ResultCollector`` we hand back and the
reference's own post-processing (``Result[...]``) on top of it."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refenv  # noqa: E402

pytestmark = pytest.mark.skipif(not refenv.available(), reason="reference tree not present")


class _OracleDevice:
    """Same surface as thermca_b200.device.DeviceSolver, computed by the CPU oracle."""

    def __init__(self, spec):
        self.spec = spec

    def run_rk(self, t0, t1, rtol, atol, method, num_skip, max_step=np.inf, first_step=None, progress=None):
        from rhs_oracle import integrate
        if progress is not None:                       # the device driver reports after every burst of attempts
            progress({"t": 0.5 * (t0 + t1)})
        opts = {}
        if np.isfinite(max_step):
            opts["max_step"] = max_step
        if first_step:
            opts["first_step"] = first_step
        out = integrate(self.spec, [t0, t1], rtol, atol, method, num_skip, **opts)
        self._last = out
        out.update({"status": 0, "accepted": out["nsteps"], "rejected": 0})
        return out

    def network_state(self):
        o = self._last
        return {"T": o["net_temps"][-1], "heat_src": o["net_heat_srcs"][-1], "capy": o["net_capys"][-1],
                "cond": o["cond_data"][-1], "clean": o["clean_data"][-1]}

    def close(self):
        pass


@pytest.fixture()
def dropin(monkeypatch, tmp_path):
    refenv.activate()
    import thermca_b200.solver as b200
    monkeypatch.setattr(b200, "DeviceSolver", _OracleDevice)
    monkeypatch.chdir(tmp_path)                   # disk_cache writes ./.thermca_cache
    b200.install()
    yield b200
    b200.uninstall()


def test_sim_through_dropin_matches_reference(dropin):
    import ref_models
    from thermca import Network
    net, kw = ref_models.coffeecup()
    res_new = net.sim(kw["time_span"], rel_tol=kw["rel_tol"], progress_bar=False)
    assert net._solve_step_time == kw["time_span"][1] and net.time == kw["time_span"][1]   # progress attributes kept up
    dropin.uninstall()
    net2, _ = ref_models.coffeecup()
    res_ref = net2.sim(kw["time_span"], rel_tol=kw["rel_tol"], progress_bar=False)
    times = np.linspace(0.0, 1200.0, 10)
    for elem_new, elem_ref in zip(net.model._get_all_elems_of_types(type(net._node_elems[list(net._node_elems)[0]][0])),
                                  net2.model._get_all_elems_of_types(type(net2._node_elems[list(net2._node_elems)[0]][0]))):
        a, b = res_new[elem_new].temp(times), res_ref[elem_ref].temp(times)
        assert np.allclose(a, b, rtol=1e-12)
    assert np.allclose(res_new._data.times, res_ref._data.times, rtol=1e-13)
    assert np.allclose(res_new._data.net_temps, res_ref._data.net_temps, rtol=1e-12)
    assert np.allclose(np.asarray(res_new._data.net_clean_conds), np.asarray(res_ref._data.net_clean_conds), rtol=1e-9,
                       atol=1e-9)


def test_link_results_and_part_views_through_dropin(dropin):
    """Result accessors that read conductances, io_temps views and MOR reconstruction."""
    import ref_models
    net, kw = ref_models.plate_mor_var()
    res = net.sim(kw["time_span"], method=kw["method"], progress_bar=False)
    dropin.uninstall()
    net2, _ = ref_models.plate_mor_var()
    ref = net2.sim(kw["time_span"], method=kw["method"], progress_bar=False)
    assert np.allclose(res._data.times, ref._data.times, rtol=1e-13)
    assert np.allclose(res._data.io_temps, ref._data.io_temps, rtol=1e-10, atol=1e-12)
    part, part2 = net._fe_parts[0], net2._fe_parts[0]
    t_end = float(kw["time_span"][1])
    assert np.allclose(res[part].temp(t_end), ref[part2].temp(t_end), rtol=1e-10)
    assert np.allclose(res[part.surf.top].temp([10.0, t_end]), ref[part2.surf.top].temp([10.0, t_end]), rtol=1e-10)


def test_result_doctest_table_through_dropin(dropin):
    """The printed table of the ``Result`` doctest (thermca/resultdata.py:125-131): LPPart steel sheet,
    spatially interpolated temperatures from the reference's own post-processing on our collector."""
    from numpy import geomspace
    import ref_models
    net, kw = ref_models.steel_sheet()
    res = net.sim(tuple(kw["time_span"]), progress_bar=False)
    sheet = net._lp_parts[0] if hasattr(net, "_lp_parts") else [e for e in net.model.elems if e.name == "sheet"][0]
    frame = res[sheet].temp_frame(time=geomspace(0.1, 36000.0, num=5), lctn=[(0.0, 0.0, 0.0), (0.0, 1.0, 0.0)])
    table = np.array([[0.010217, 0.000268], [0.249103, 0.006545], [5.769231, 0.157884], [83.300245, 3.310066],
                      [253.019457, 45.451702]])
    assert np.allclose(frame.to_numpy(), table, rtol=0, atol=5.1e-7)       # the digits the doctest prints


def test_unknown_method_and_untraceable_callable(dropin):
    from thermca import Model, Node, BoundNode, CondLink, Network
    from thermca_b200.trace import UntraceableError
    with Model() as model:
        a = Node(capy=1.0, init_temp=0.0)
        b = BoundNode(temp=1.0)
        def halving(t0, t1, matl):                   # trip count depends on data: no finite set of paths
            x = abs(t0 - t1) + 2.0
            while x > 1.0:
                x = x / 2.0
            return x
        CondLink(a, b, cond=halving)
    with pytest.raises(UntraceableError):
        Network(model).sim([0, 1.0], progress_bar=False)
    with Model() as model2:
        a = Node(capy=1.0, init_temp=0.0)
        b = BoundNode(temp=1.0)
        CondLink(a, b, cond=1.0)
    with pytest.raises(ValueError):
        Network(model2).sim([0, 1.0], method="NOPE", progress_bar=False)


def test_linear_diag_search_is_result_identical(dropin):
    import scipy.sparse as sp
    from thermca._utils.sparse import coo_diag_data_idxs
    from thermca_b200.solver import linear_coo_diag_data_idxs
    rng = np.random.default_rng(0)
    m = sp.random(60, 60, density=0.1, random_state=2, format="csr") + sp.eye(60, format="csr")
    m = sp.csr_matrix(m)
    m[7, 7] = 0.0
    m.eliminate_zeros()                       # a row without a stored diagonal
    coo = m.tocoo()
    a = coo_diag_data_idxs(m.shape, m.data.shape, coo.row, coo.col)
    b = linear_coo_diag_data_idxs(m.shape, m.data.shape, coo.row, coo.col)
    assert np.array_equal(np.asarray(a), b)
    # and a model builds through the patched constructors
    import ref_models
    net, kw = ref_models.kuhn_cuboid()
    assert net._ffe_ios[0].fe_sys.dof == 364
