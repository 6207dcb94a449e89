"""MOR basis generation on the device.

Device replacement of ``thermca.fem.mor_io.transform`` (thermca/fem/mor_io.py:139-166) and its inverse-Krylov
Arnoldi process (``transformation_matrices`` / ``arnoldi_iteration``, :169-220):

* the reference factorises the symmetrised state matrix with a sparse LU and does ``dim-1`` solves per
  coupling surface; here every solve is one launch of the cooperative Jacobi-PCG kernel (csrc/pcg.cuh) on
  the SELL-32 copy of the same matrix (it is symmetric positive definite);
* the modified Gram-Schmidt loop becomes two passes of classical Gram-Schmidt (csrc/krylov.cu);
* the projections ``V^T A V`` are r SELL products (csrc/parts.cuh) followed by a plain FP64 library GEMM.

Same inputs, same outputs (numpy), same ``disk_cache`` key (function name ``transform`` + joblib hash of the
arguments, thermca/_utils/func_tools.py:74-92), so caches written by either implementation are found by
the other.  Krylov vectors amplify solver rounding (measured ~3e3 x for dim = 10), so parity with the
reference's LU-based result is at the 1e-6 level of the reduced operators, not bit-wise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .linsolve import DeviceOperator, _p


def _arnoldi_device(op, Bs, dim, pcg_rtol, work, h2):
    """Orthonormal bases (one per row of ``Bs``; each ``dim`` columns, column-major: tensor (dim, n)) of the Krylov
    spaces of A^-1 started at those vectors (thermca/fem/mor_io.py:176-218).  The k-th solves of all bases run
    together (one matrix sweep per PCG iteration, csrc/pcg_multi.cu).  Columns after a breakdown (|v| <= 1e-12)
    stay zero, as in the reference."""
    torch = op.torch
    lib, n = op.lib, op.n
    stream = op._stream()
    S = int(Bs.shape[0])
    Qs = [torch.zeros((dim, n), dtype=torch.float64, device=op.dev) for _ in range(S)]
    alive = [True] * S
    for s in range(S):
        v = Bs[s].clone()
        _cabi.check(lib.tc_krylov_orthogonalize(stream, n, 0, _p(Qs[s]), n, _p(v), 1, _p(work), _p(None), _p(h2)))
        _cabi.check(lib.tc_krylov_normalize(stream, n, _p(v), _p(h2), _p(Qs[s][0])))
    for k in range(dim - 1):
        idx = [s for s in range(S) if alive[s]]
        # measured on B200 (scripts/pcg_multi_bandwidth.py): up to 6 right-hand sides per launch cost ~45 us per
        # right-hand side and sweep at 1 M DOF, 8 cost 64 us (register-limited occupancy) -> groups of <= 6, else <= 4
        step = len(idx) if len(idx) <= 6 else 4
        for lo in range(0, len(idx), max(step, 1)):
            grp = idx[lo: lo + step]
            V = op.solve_multi(torch.stack([Qs[s][k] for s in grp]), pcg_rtol)
            for j, s in enumerate(grp):
                v = V[j].contiguous()
                _cabi.check(lib.tc_krylov_orthogonalize(stream, n, k + 1, _p(Qs[s]), n, _p(v), 2, _p(work), _p(None),
                                                        _p(h2)))
                if float(h2.item()) ** 0.5 <= 1e-12:
                    alive[s] = False
                    continue
                _cabi.check(lib.tc_krylov_normalize(stream, n, _p(v), _p(h2), _p(Qs[s][k + 1])))
    return Qs


def transform_device(M_is, A_body, A_film, A, B, C, dim, pcg_rtol=1e-13, device=None, return_info=False):
    """Device version of ``transform(M_is, A_body, A_film, A, B, C, dim)`` (thermca/fem/mor_io.py:139-166)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("thermca_b200.mor_basis needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _cabi.load_library()
    B = np.asarray(B, dtype=np.float64)
    C_out = np.asarray(C, dtype=np.float64)
    n, S = B.shape
    dim = int(dim)
    opA = DeviceOperator(A, dev)
    opBody = DeviceOperator(A_body, dev)
    mis = torch.from_numpy(np.ascontiguousarray(M_is.diagonal())).to(dev)
    film_diag = [torch.from_numpy(np.ascontiguousarray(Af.diagonal())).to(dev) for Af in A_film] if A_film is not None else []
    Bd = torch.from_numpy(np.ascontiguousarray(B.T)).to(dev)            # (S, n)
    Cd = torch.from_numpy(np.ascontiguousarray(C_out)).to(dev)          # (n, S)
    work = torch.empty(int(lib.tc_krylov_work_doubles(dim)), dtype=torch.float64, device=dev)
    h2 = torch.zeros(1, dtype=torch.float64, device=dev)
    Y = torch.empty((dim, n), dtype=torch.float64, device=dev)
    Ar_body = np.empty((S, dim, dim))
    Ar_film = np.empty((len(film_diag), S, dim, dim)) if A_film is not None else None
    Br_T = np.empty((S, dim))
    Crs = np.empty((S, dim, S))
    VT1 = np.empty((S, n, dim))
    VT0 = np.empty((S, dim, n))
    Qs = _arnoldi_device(opA, Bd, dim, pcg_rtol, work, h2)             # per surface (dim, n): row j = basis vector j
    for s in range(S):
        Q = Qs[s]
        for j in range(dim):
            opBody.matvec_into(Q[j], Y[j])
        Ar_body[s] = torch.matmul(Q, Y.T).cpu().numpy()                 # V^T A_body V
        for f, d in enumerate(film_diag):
            Ar_film[f, s] = torch.matmul(Q, (Q * d).T).cpu().numpy()
        Br_T[s] = torch.mv(Q, Bd[s]).cpu().numpy()
        Crs[s] = torch.matmul(Q, Cd).cpu().numpy()
        V1 = (Q * mis).T.contiguous()                                   # (n, dim) = M_is V
        VT1[s] = V1.cpu().numpy()
        # pinv of a full-column-rank matrix: (V1^T V1)^-1 V1^T; rank-deficient bases (Arnoldi breakdown) go
        # through the SVD-based library routine like the reference
        rank_ok = bool((Q.abs().amax(dim=1) > 0).all())
        if rank_ok:
            G = torch.matmul(V1.T, V1)
            VT0[s] = torch.linalg.solve(G, V1.T).cpu().numpy()
        else:
            VT0[s] = torch.linalg.pinv(V1).cpu().numpy()
    out = (Ar_body, Ar_film, Br_T, Crs, VT0, VT1)
    if return_info:
        return out, {"pcg_iterations": opA.iterations, "pcg_solves": opA.solves}
    return out


_original = {}


def install():
    """Route ``MORIO.__init__``'s basis build (thermca/fem/mor_io.py:75-77) through the device, keeping the
    reference's on-disk cache (same function name in the cache key)."""
    import thermca.fem.mor_io as mio
    from thermca._utils.func_tools import disk_cache
    if "transform" not in _original:
        _original["transform"] = mio.transform

    def transform(M_is, A_body, A_film, A, B, C, dim):
        return transform_device(M_is, A_body, A_film, A, B, C, dim)
    mio.transform = disk_cache(transform)


def uninstall():
    if "transform" in _original:
        import thermca.fem.mor_io as mio
        mio.transform = _original.pop("transform")
