"""SELL-32 device copy of a scipy CSR matrix with the two operations the one-off computations around the
transient solve need: products (csrc/parts.cuh) and symmetric positive definite solves by the cooperative
Jacobi-PCG kernel (csrc/pcg.cuh).

The PCG kernel applies ``q = m (p + shift A p)`` (the implicit time step's operator).  With ``m = eps``,
``shift = 1/eps`` and ``eps = 2**-80`` (exact scalings) this is ``A p + eps p``: the plain operator up to a
relative perturbation of 1e-24, far below any solver tolerance.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .device import csr_to_sell

_EPS = 2.0 ** -80


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceOperator:
    """SELL-32 copy of a scipy CSR matrix with the two things the basis build needs: products and SPD solves."""

    def __init__(self, A, dev):
        import scipy.sparse as sp
        import torch
        self.torch, self.dev = torch, dev
        self.lib = _cabi.load_library()
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.n = A.shape[0]
        s = csr_to_sell(A.indptr, A.indices, A.data)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.nslices = s["nslices"]
        self.slice_ptr, self.col, self.val = up(s["slice_ptr"]), up(s["col"]), up(s["val"])
        self.diag = up(A.diagonal())
        self.mass = torch.full((self.n,), _EPS, dtype=torch.float64, device=dev)
        self.iterations = 0
        self.solves = 0

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def matvec_into(self, x, y):
        """y = A x"""
        _cabi.check(self.lib.tc_sell_spmv(self._stream(), 0, self.n, self.nslices, _p(self.slice_ptr), _p(self.col),
                                          _p(None), _p(self.val), _p(x), _p(y), _p(None), _p(None), 0.0, 0))
        y.neg_()
        return y

    def solve(self, b, rtol, maxit=None, precond="jacobi"):
        """x = A^-1 b by preconditioned CG (A symmetric positive definite); b, x device vectors.
        ``precond``: ``"jacobi"`` (default) or ``"block"`` (32 x 32 slice blocks, csrc/pcg.cuh)."""
        torch = self.torch
        x = torch.zeros(self.n, dtype=torch.float64, device=self.dev)
        rhs = b / _EPS
        iters = C.c_int32(0)
        torch.cuda.synchronize(self.dev)
        if precond == "block":
            if getattr(self, "_binv", None) is None:
                self._binv = slice_block_inverses(self.slice_ptr, self.col, self.val, self.mass, self.n, 1.0 / _EPS)
                torch.cuda.synchronize(self.dev)
            _cabi.check(self.lib.tc_pcg_solve_bj(self._stream(), self.n, self.nslices, _p(self.slice_ptr), _p(self.col),
                                                 _p(self.val), _p(self.mass), _p(self.diag), _p(rhs), _p(x), 1.0 / _EPS,
                                                 float(rtol), int(maxit) if maxit else min(20 * self.n + 1000, 100000),
                                                 C.byref(iters), _p(self._binv)))
            self.last_iterations = int(iters.value)
            self.iterations += int(iters.value)
            self.solves += 1
            return x
        elif precond != "jacobi":
            raise ValueError(f"unknown preconditioner {precond!r}")
        _cabi.check(self.lib.tc_pcg_solve(self._stream(), 0, self.n, self.nslices, _p(self.slice_ptr), _p(self.col),
                                          _p(self.val), _p(self.mass), _p(self.diag), _p(rhs), _p(x), 1.0 / _EPS,
                                          float(rtol), int(maxit) if maxit else min(20 * self.n + 1000, 100000), C.byref(iters)))
        self.last_iterations = int(iters.value)
        self.iterations += int(iters.value)
        self.solves += 1
        return x

    def solve_multi(self, B, rtol, maxit=None):
        """X = A^-1 B for up to 8 right-hand sides (rows of the (K, n) tensor B) with one matrix sweep per PCG
        iteration (csrc/pcg_multi.cu).  Returns the (K, n) solutions."""
        torch = self.torch
        K = int(B.shape[0])
        if K == 1:
            return self.solve(B[0], rtol, maxit)[None, :]
        Bi = B.T.contiguous()                                  # [n][K]: the K values of a row contiguous
        Xi = torch.empty_like(Bi)
        nwork = int(self.lib.tc_pcg_multi_work_doubles(self.n, K))
        if getattr(self, "_mwork", None) is None or self._mwork.numel() < nwork:
            self._mwork = torch.empty(nwork, dtype=torch.float64, device=self.dev)
        iters = torch.zeros(8, dtype=torch.int32, device=self.dev)
        _cabi.check(self.lib.tc_pcg_solve_multi(self._stream(), self.n, self.nslices, _p(self.slice_ptr), _p(self.col),
                                                _p(self.val), _p(self.diag), K, _p(Bi), _p(Xi), _p(self._mwork),
                                                float(rtol), int(maxit) if maxit else min(20 * self.n + 1000, 100000),
                                                _p(iters)))
        it = iters[:K].cpu().numpy()
        self.last_iterations = int(it.max())
        self.iterations += int(it.max())                       # matrix sweeps
        self.solves += K
        return Xi.T.contiguous()

    def residual_into(self, x, b, r):
        """r = b - A x (separate SpMV launch; used to re-check a solve independently of the PCG recurrence)"""
        _cabi.check(self.lib.tc_sell_spmv(self._stream(), 1, self.n, self.nslices, _p(self.slice_ptr), _p(self.col),
                                          _p(None), _p(self.val), _p(x), _p(r), _p(b), _p(None), 0.0, 0))
        return r


def slice_block_inverses(slice_ptr, col, val, mass, n, shift, chunk=1 << 15):
    """Block-Jacobi preconditioner of ``M + shift*K`` for a SELL-32 operator ``A = M^-1 K`` on the device: the inverses
    of the 32 x 32 diagonal blocks that belong to the slices, symmetric, FP32, layout ``[slice][j][i]`` (what
    ``csrc/pcg.cuh:tc_block_apply`` reads).  The blocks are gathered with index arithmetic on the device and inverted
    by the batched library routine (one-off per (h, films): a library call, like the projections of ``mor_basis``)."""
    import torch
    nsl = int(slice_ptr.numel()) - 1
    dev = val.device
    out = torch.empty((nsl, 32, 32), dtype=torch.float32, device=dev)
    eye = torch.eye(32, dtype=torch.float64, device=dev)
    for s0 in range(0, nsl, chunk):
        s1 = min(nsl, s0 + chunk)
        p0, p1 = int(slice_ptr[s0].item()), int(slice_ptr[s1].item())
        widths = (slice_ptr[s0 + 1: s1 + 1] - slice_ptr[s0: s1]) // 32
        sl = torch.repeat_interleave(torch.arange(s0, s1, device=dev), widths * 32)
        pos = torch.arange(p0, p1, device=dev)
        lane = (pos - slice_ptr[sl]) & 31
        row = sl * 32 + lane
        c = col[p0:p1].to(torch.int64)
        v = val[p0:p1]
        live = row < n
        inblk = live & ((c >> 5) == sl) & (v != 0.0)
        blocks = torch.zeros((s1 - s0, 32, 32), dtype=torch.float64, device=dev)
        m_row = mass[row.clamp(max=n - 1)]
        blocks.index_put_(((sl - s0)[inblk], lane[inblk], (c & 31)[inblk]), (m_row * shift * v)[inblk], accumulate=True)
        rows_all = torch.arange(s0 * 32, s1 * 32, device=dev)
        diag = torch.where(rows_all < n, mass[rows_all.clamp(max=n - 1)], torch.ones_like(rows_all, dtype=torch.float64))
        blocks += eye * diag.reshape(s1 - s0, 32, 1)
        inv = torch.linalg.inv(blocks)
        inv = 0.5 * (inv + inv.transpose(1, 2))
        out[s0:s1] = inv.transpose(1, 2).to(torch.float32)
    return out.contiguous()
