"""Two "ranks" of the row-partitioned kernels (csrc/dist.cu) as sibling handles of ONE process on one GPU, each on its
own stream (``tc_dist_open_peers_local``): the single-GPU test box reaches what world size 1 never touches — boundary
slices and their ghost side lists, the per-slice send lists, LL-encoded ghosts and the reduction mailboxes — for the
fused theta/PCG step in its three formulations, the residual of the explicit pairs (which aliases its output with the
load vector) and the solve-only instantiation of the W-methods.  References: the CPU oracle (oracle/theta_oracle.py) and
scipy on the same operator.  The real multi-process runs are scripts/dist_check.py and friends under torchrun.

If the two cooperative kernels cannot run side by side on this device the first one times out waiting for its peer
(TC_DIST_TIMEOUT_S) and the test is skipped — wrong numbers fail."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu

H, THETA = 2.0, 0.5


def _problem():
    import scipy.sparse as sp
    from thermca_b200.synth import cuboid_spec
    spec = cuboid_spec(8, 23, 4, jitter=0.15)                      # 9 * 24 * 5 = 1080 rows, 34 slices
    p = spec["ffe"][0]
    n = int(p["dof"])
    A = sp.csr_matrix((p["A_data"], p["A_body"]["indices"], p["A_body"]["indptr"]), shape=(n, n))
    A.sort_indices()
    mass = 1.0 / p["M_inv"]
    b = np.zeros(n)
    for s, fl in enumerate([1000.0 / p["surf_areas"][0], 200.0]):
        b[p["B_idx"][s]] += p["B_val"][s] * fl
    return A, mass, b


class _Ranks:
    """World-size-2 partition of one operator driven through the C ABI, both ranks in this process."""

    def __init__(self, A, mass, b, world=2):
        import torch
        from thermca_b200 import _cabi
        from thermca_b200.dist import csr_rows_to_sell, even_row_bounds, send_plan
        from thermca_b200.partition import split_local_rows
        self.torch, self.lib, self.cabi = torch, _cabi.load_library(), _cabi
        self.world, n = world, A.shape[0]
        self.bounds = even_row_bounds(n, world, 32)
        dev = torch.device("cuda", 0)
        adiag = A.diagonal()
        parts = []
        for w in range(world):
            r0, r1 = int(self.bounds[w]), int(self.bounds[w + 1])
            e0, e1 = int(A.indptr[r0]), int(A.indptr[r1])
            lp, li, keep, pre = split_local_rows(A.indptr, A.indices, A.data, r0, r1)
            sp_, col, val = csr_rows_to_sell(lp, li, A.data[e0:e1][keep], r1 - r0)
            parts.append((r0, r1, sp_, col, val, pre))
        all_ghosts = [p[5]["ghosts"] for p in parts]
        ext = [len(g) for g in all_ghosts]
        up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev).to(dt)
        self.h, self.keep, self.streams, self.rows = [], [], [], []
        for w, (r0, r1, sp_, col, val, pre) in enumerate(parts):
            st = torch.cuda.Stream(dev)
            h = _cabi.vp()
            _cabi.check(self.lib.tc_dist_create(C.byref(h), C.c_void_p(st.cuda_stream), w, world, r1 - r0, ext[w]))
            self.h.append(h); self.streams.append(st); self.rows.append((r0, r1))
        arr = (_cabi.vp * world)(*[h.value for h in self.h])
        for w, (r0, r1, sp_, col, val, pre) in enumerate(parts):
            _cabi.check(self.lib.tc_dist_open_peers_local(self.h[w], arr, (_cabi.i64 * world)(*ext)))
            sr, spr, sd, nb = [np.ascontiguousarray(a) for a in send_plan(all_ghosts, self.bounds, w)]
            ptr = lambda a: a.ctypes.data_as(C.c_void_p)
            _cabi.check(self.lib.tc_dist_set_halo(self.h[w], len(sr), ptr(sr), ptr(spr), ptr(sd), len(nb), ptr(nb)))
            t = {"sp": up(sp_, torch.int64), "col": up(col, torch.int32), "val": up(val, torch.float64),
                 "mass": up(mass[r0:r1], torch.float64), "adiag": up(adiag[r0:r1], torch.float64),
                 "b": up(b[r0:r1], torch.float64), "bidx": up(pre["bidx"], torch.int32), "gptr": up(pre["gptr"], torch.int64),
                 "gslot": up(pre["gslot"] if len(pre["gslot"]) else np.zeros(1), torch.int32),
                 "gval": up(pre["gval"] if len(pre["gval"]) else np.zeros(1), torch.float64)}
            self.keep.append(t)
            p = lambda x: C.c_void_p(x.data_ptr())
            n_bnd = int((pre["bidx"] >= 0).sum())
            assert n_bnd > 0 and len(sr) > 0, "the partition must have boundary slices and send rows"
            torch.cuda.synchronize()
            _cabi.check(self.lib.tc_dist_set_operator(self.h[w], len(sp_) - 1, p(t["sp"]), p(t["col"]), C.c_void_p(0),
                                                      p(t["val"]), p(t["mass"]), p(t["adiag"]), p(t["b"]), n_bnd,
                                                      p(t["bidx"]), p(t["gptr"]), p(t["gslot"]), p(t["gval"])))
        torch.cuda.synchronize()

    def vec(self, v_global):
        torch = self.torch
        return [torch.from_numpy(np.ascontiguousarray(v_global[r0:r1])).to("cuda") for r0, r1 in self.rows]

    def done(self):
        """Synchronise both ranks; skip when a rank timed out waiting for its sibling, fail on any other error."""
        rcs = [int(self.lib.tc_dist_check(h)) for h in self.h]
        if any(rc == -3 for rc in rcs):
            for h in self.h:
                self.lib.tc_dist_reset_error(h)
            pytest.skip("the sibling cooperative kernels did not run concurrently on this device")
        assert rcs == [0] * self.world, (rcs, self.lib.tc_last_error())

    def close(self):
        for h in self.h:
            self.lib.tc_dist_close_peers(h)
        for h in self.h:
            self.lib.tc_dist_destroy(h)


@pytest.fixture()
def ranks(monkeypatch):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    monkeypatch.setenv("TC_DIST_TIMEOUT_S", "3")
    A, mass, b = _problem()
    r = _Ranks(A, mass, b)
    yield r, A, mass, b
    r.close()


@pytest.mark.parametrize("mode", [0, 2, 1], ids=["classic", "single-reduction", "pipelined"])
def test_two_sibling_ranks_theta_steps_match_the_cpu_oracle(ranks, mode):
    from theta_oracle import theta_cg_steps
    r, A, mass, b = ranks
    torch, lib, cabi = r.torch, r.lib, r.cabi
    n, steps = A.shape[0], 3
    y0 = np.full(n, 20.0)
    for h, y in zip(r.h, r.vec(y0)):
        cabi.check(lib.tc_dist_set_pcg(h, mode))
        cabi.check(lib.tc_dist_set_state(h, C.c_void_p(y.data_ptr())))
    for h in r.h:                                              # enqueue only: the ranks run side by side
        cabi.check(lib.tc_dist_theta_steps(h, steps, THETA, H, 1e-12, 2000))
    r.done()
    out = r.vec(np.zeros(n))
    for h, o in zip(r.h, out):
        cabi.check(lib.tc_dist_get_state(h, C.c_void_p(o.data_ptr())))
    torch.cuda.synchronize()
    y_dev = np.concatenate([o.cpu().numpy() for o in out])
    y_cpu, it_cpu, _ = theta_cg_steps(A, mass, b, y0, steps, H, THETA, 1e-12)
    assert np.max(np.abs(y_dev - y_cpu)) <= 1e-9 * np.max(np.abs(y_cpu))
    st = (cabi.i64 * 6)()
    cabi.check(lib.tc_dist_stats(r.h[0], st))
    assert abs(int(st[1]) - it_cpu) <= steps and int(st[4]) == 0


def test_two_sibling_ranks_residual_and_stage_solve(ranks):
    """b - A v with the output aliasing the load vector (tc_dist_residual, explicit pairs) and the solve-only
    instantiation z = (I + c h A)^-1 rhs (tc_dist_solve, W-methods) across a partition boundary."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    r, A, mass, b = ranks
    torch, lib, cabi = r.torch, r.lib, r.cabi
    n = A.shape[0]
    rng = np.random.default_rng(5)
    v = 20.0 + rng.standard_normal(n)
    vs, outs = r.vec(v), r.vec(b)                              # out holds b on entry
    for _ in range(2):                                         # twice: both ghost buffers of the stage vectors
        for o, (r0, r1) in zip(outs, r.rows):
            o.copy_(torch.from_numpy(b[r0:r1]))
        torch.cuda.synchronize()
        for h, x, o in zip(r.h, vs, outs):
            cabi.check(lib.tc_dist_residual(h, C.c_void_p(x.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(0)))
        r.done()
        f_dev = np.concatenate([o.cpu().numpy() for o in outs])
        f_ref = b - A @ v
        assert np.max(np.abs(f_dev - f_ref)) <= 1e-12 * np.max(np.abs(f_ref))
    coef, hh = 0.4358665215, 5.0
    h_dev = torch.tensor([hh], dtype=torch.float64, device="cuda")
    rhs = rng.standard_normal(n)
    rs, zs = r.vec(rhs), r.vec(np.zeros(n))
    torch.cuda.synchronize()
    for h, x, z in zip(r.h, rs, zs):
        cabi.check(lib.tc_dist_solve(h, coef, C.c_void_p(h_dev.data_ptr()), C.c_void_p(x.data_ptr()),
                                     C.c_void_p(z.data_ptr()), 1e-13, 2000, C.c_void_p(0), C.c_void_p(0)))
    r.done()
    z_dev = np.concatenate([z.cpu().numpy() for z in zs])
    z_ref = spla.spsolve((sp.identity(n, format="csc") + coef * hh * A.tocsc()), rhs)
    assert np.max(np.abs(z_dev - z_ref)) <= 1e-9 * np.max(np.abs(z_ref))
