"""Physical validation the reference ships: the steel plate against ANSYS
(thermca/tests/test_fe_part.py:230-303, 'cuboid flat ansys.csv'; sample points in tests/golden/plate_ansys.npz,
oracle/make_golden_ansys.py), with the reference test's own tolerances.  The reference runs this case with
LSODA; the drop-in serves LSODA requests with the adaptive third-order W-method (ROW3), the second-order ROS2-PCG and
RK45 are checked too."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _ansys():
    return np.load(os.path.join(GOLDEN, "plate_ansys.npz"))


def _probe(spec, out, a):
    """Mean bottom-surface temperature and the two probed vertices at the ANSYS sample times (linear in time, as
    thermca's Result.temp interpolates between stored steps)."""
    part = spec["ffe"][0] if spec["ffe"] else None
    names = [str(s) for s in a["surf_names"]]
    t = out["times"]
    if part is not None:
        net_idx = int(np.asarray(part["surf_net_idxs"])[names.index("bottom")])
        io = out["io_temps"]
        bv = np.interp(a["time"], t, io[:, int(a["dof_bottom_vertex"])])
        tv = np.interp(a["time"], t, io[:, int(a["dof_top_vertex"])])
    else:
        net_idx = int(np.asarray(spec["mor"][0]["surf_net_idxs"])[names.index("bottom")])
        bv = tv = None
    bottom = np.interp(a["time"], t, out["net_temps"][:, net_idx])
    return bottom, bv, tv


@pytest.mark.parametrize("method", ["ROW3_PCG", "ROS2_PCG", "RK45"])
def test_full_order_plate_against_ansys(golden, method):
    from thermca_b200.device import DeviceSolver
    a = _ansys()
    spec, _ = golden("plate_ffe")
    dev = DeviceSolver(spec)
    t_end = float(a["time"][-1])
    if method == "RK45":
        # explicit RK45 runs at its stability limit on this stiff body; at 1e-3 the probed corner vertex is 0.8 %
        # off (the reference's RK45 gives the same trajectory, parity 1e-9), at 1e-5 it meets the 0.6 %
        out = dev.run_rk(0.0, t_end, 1e-5, 1e-6, "RK45", num_skip=20)
    else:
        # second-order ROS2 with linear interpolation between its (long) steps needs 1e-4 for the 0.3 % the
        # reference reaches with LSODA at 1e-3
        out = dev.run_ros2(0.0, t_end, 1e-4, 1e-6, scheme="ROW3" if method == "ROW3_PCG" else "ROS2")
        print(method, "on the full-order plate: accepted", out["accepted"], "rejected", out["rejected"])
    dev.close()
    bottom, bv, tv = _probe(spec, out, a)
    init = 20.0
    assert np.allclose(a["bottom"] - init, bottom - init, rtol=0.003)              # test_fe_part.py:297-299
    assert np.allclose(a["top_vertex"] - init, tv - init, rtol=0.02)
    assert np.allclose(a["bottom_vertex"] - init, bv - init, rtol=0.006)


@pytest.mark.parametrize("method", ["RK45", "ROS2_PCG", "ROW3_PCG"])
def test_reduced_plate_against_ansys(golden, method):
    """MOR body (golden plate_mor, mor_dof = 10; the reference test uses 30): mean bottom temperature within the
    reference test's 0.3 % (test_fe_part.py:301).  With ROS2 the long steps of this 9 h simulation put the reduced
    blocks into the implicit branch of the block solve (dense CG per surface)."""
    from thermca_b200.device import DeviceSolver
    a = _ansys()
    spec, _ = golden("plate_mor")
    dev = DeviceSolver(spec)
    if method == "RK45":
        out = dev.run_rk(0.0, float(a["time"][-1]), 1e-3, 1e-6, "RK45", num_skip=20)
    else:
        # the third-order scheme takes ~130 steps over these 9 h at 1e-4: the linear interpolation between stored
        # steps (what Result.temp does) then dominates the 0.3 %, so it runs at 1e-5 (~290 steps; ROS2: 1630 at 1e-4)
        out = dev.run_ros2(0.0, float(a["time"][-1]), 1e-5 if method == "ROW3_PCG" else 1e-4, 1e-6,
                           scheme="ROW3" if method == "ROW3_PCG" else "ROS2")
        print(method, "on the reduced plate: accepted", out["accepted"], "rejected", out["rejected"])
    dev.close()
    bottom, _, _ = _probe(spec, out, a)
    assert np.allclose(a["bottom"] - 20.0, bottom - 20.0, rtol=0.003)
