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

    def solve(self, b, rtol, maxit=None):
        """x = A^-1 b by Jacobi-PCG (A symmetric positive definite); b, x device vectors."""
        torch = self.torch
        x = torch.zeros(self.n, dtype=torch.float64, device=self.dev)
        rhs = b / _EPS
        iters = C.c_int32(0)
        torch.cuda.synchronize(self.dev)
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
