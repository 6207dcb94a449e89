"""CPU ORACLE for the implicit path — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference has no implicit scheme ("parity unpinned" by the
reference's own tests), so this file restates, on the reference's operators
(FFEIO.A / M_inv, thermca/fem/ffe_io.py:40-65):

* ``theta_step_direct``  – one theta step with scipy's sparse direct solver (parity target
  of the device PCG at 1e-9),
* ``theta_cg_steps``     – the like-for-like CPU port of the device algorithm: Jacobi-PCG
  on M + theta*h*K, SpMV by scipy ``csr_matvec`` (1 thread, what thermca/_utils/sparse.py:10
  falls back to) or by torch's CSR kernel on all host threads (stands in for the optional
  MKL path, thermca/_utils/sparse.py:6-9).
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def theta_step_direct(A, f, y, h, theta):
    n = A.shape[0]
    z = spla.spsolve((sp.identity(n, format="csc") + theta * h * A.tocsc()), f)
    return y + h * z


class _TorchCsr:
    def __init__(self, A, threads):
        import torch
        torch.set_num_threads(threads)
        self.torch = torch
        self.A = torch.sparse_csr_tensor(torch.from_numpy(A.indptr.astype(np.int64)),
                                         torch.from_numpy(A.indices.astype(np.int64)),
                                         torch.from_numpy(A.data), size=A.shape, dtype=torch.float64)

    def __call__(self, x):
        return (self.A @ self.torch.from_numpy(x)).numpy()


def _c_omp_steps(A, mass, b, y, nsteps, h, theta, rtol, maxit, threads):
    """oracle/c/theta_cg_omp.c through ctypes (all host threads)."""
    import ctypes as C
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from build_c import build
    lib = C.CDLL(build())
    lib.theta_cg_steps.restype = C.c_int64
    n = A.shape[0]
    ip, ix = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    a = np.ascontiguousarray(A.data)
    adiag = np.ascontiguousarray(A.diagonal())
    mass = np.ascontiguousarray(mass); b = np.ascontiguousarray(b)
    yo = np.array(y, dtype=np.float64)
    work = np.empty(6 * n)
    ptr = lambda v: v.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    iters = lib.theta_cg_steps(C.c_int64(n), ptr(ip), ptr(ix), ptr(a), ptr(mass), ptr(adiag), ptr(b), ptr(yo),
                               C.c_int32(nsteps), C.c_double(h), C.c_double(theta), C.c_double(rtol),
                               C.c_int32(maxit), ptr(work), C.c_int32(threads))
    return yo, int(iters), time.perf_counter() - t0


def theta_cg_steps(A, mass, b, y, nsteps, h, theta=0.5, rtol=1e-8, maxit=2000, backend="scipy", threads=1):
    """``nsteps`` theta steps of y' = b - A y with Jacobi-PCG on mass*(I + theta h A).
    Returns (y, total CG iterations, seconds)."""
    A = sp.csr_matrix(A)
    if backend == "c_omp":
        return _c_omp_steps(A, mass, b, y, nsteps, h, theta, rtol, maxit, threads)
    matvec = _TorchCsr(A, threads) if backend == "torch" else (lambda x: A @ x)
    shift = theta * h
    adiag = A.diagonal()
    dinv = 1.0 / (mass * (1.0 + shift * adiag))
    iters = 0
    t0 = time.perf_counter()
    for _ in range(nsteps):
        f = b - matvec(y)
        r = mass * f
        x = np.zeros_like(y)
        z = dinv * r
        p = z.copy()
        rho = r @ z
        tol2 = rtol * rtol * (r @ r)
        for _it in range(maxit):
            q = mass * (p + shift * matvec(p))
            alpha = rho / (p @ q)
            x += alpha * p
            r -= alpha * q
            iters += 1
            rr = r @ r
            if not rr > tol2:
                break
            z = dinv * r
            rho_new = r @ z
            p = z + (rho_new / rho) * p
            rho = rho_new
        y = y + h * x
    return y, iters, time.perf_counter() - t0


def theta_pipecg_steps(A, mass, b, y, nsteps, h, theta=0.5, rtol=1e-8, maxit=2000):
    """CPU restatement of the PIPELINED Jacobi-PCG of csrc/dist.cu (``tc_dist_theta_pipe_kernel``): the recurrences of
    Ghysels & Vanroose (Parallel Computing 40, 2014, Alg. 4) in Jacobi-scaled variables
    u = D r, wh = D w, zh = D z, q = D s with B = D (M + theta h K) = c o (I + theta h A):
    one operator sweep and ONE reduction (gamma, delta, r.r) per iteration.  Mathematically the same iterates as
    ``theta_cg_steps``.  Returns (y, total iterations, seconds)."""
    A = sp.csr_matrix(A)
    shift = theta * h
    t = 1.0 + shift * A.diagonal()
    g = mass * t                      # 1 / d
    c = 1.0 / t
    iters = 0
    t0 = time.perf_counter()
    for _ in range(nsteps):
        f = b - A @ y
        u = (mass * f) / g
        x = np.zeros_like(y)
        wh = c * (u + shift * (A @ u))
        zh = np.zeros_like(y); q = np.zeros_like(y); p = np.zeros_like(y)
        gu = g * u
        gamma, delta, bb = u @ gu, wh @ gu, gu @ gu
        tol2 = rtol * rtol * bb
        gamma_old = alpha_old = 0.0
        if bb > 0.0:
            for it in range(maxit):
                if it == 0:
                    beta, alpha = 0.0, gamma / delta
                else:
                    beta = gamma / gamma_old
                    alpha = gamma / (delta - beta * gamma / alpha_old)
                nh = c * (wh + shift * (A @ wh))
                zh = nh + beta * zh
                q = wh + beta * q
                p = u + beta * p
                x += alpha * p
                u = u - alpha * q
                wh = wh - alpha * zh
                gu = g * u
                iters += 1
                gamma_old, alpha_old = gamma, alpha
                gamma, delta, rr = u @ gu, wh @ gu, gu @ gu
                if not rr > tol2:
                    break
        y = y + h * x
    return y, iters, time.perf_counter() - t0


def theta_srcg_steps(A, mass, b, y, nsteps, h, theta=0.5, rtol=1e-8, maxit=2000):
    """CPU restatement of the SINGLE-REDUCTION Jacobi-PCG of csrc/dist.cu (``tc_dist_theta_sr_kernel``): the classic
    iterates, with all inner products of an iteration taken from the product sweep (p.q, r.Dr, r.Dq, q.Dq, r.r, r.q,
    q.q) and rho', r'.r' expanded from them (D'Azevedo, Eijkhout & Romine 1993), so the iteration has one reduction.
    Returns (y, total iterations, seconds)."""
    A = sp.csr_matrix(A)
    shift = theta * h
    d = 1.0 / (mass * (1.0 + shift * A.diagonal()))
    iters = 0
    t0 = time.perf_counter()
    for _ in range(nsteps):
        f = b - A @ y
        r = mass * f
        x = np.zeros_like(y)
        p = d * r
        tol2 = None
        for it in range(maxit):
            q = mass * (p + shift * (A @ p))
            dr, dq = d * r, d * q
            pq, rho, rdq, qdq, rr, rq, qq = p @ q, r @ dr, r @ dq, q @ dq, r @ r, r @ q, q @ q
            if tol2 is None:
                tol2 = rtol * rtol * rr
                if not rr > 0.0:
                    break
            alpha = rho / pq
            rho_new = rho - 2.0 * alpha * rdq + alpha * alpha * qdq
            rr_new = rr - 2.0 * alpha * rq + alpha * alpha * qq
            x += alpha * p
            iters += 1
            if not rr_new > tol2:
                break
            r -= alpha * q
            p = d * r + (rho_new / rho) * p
        y = y + h * x
    return y, iters, time.perf_counter() - t0
