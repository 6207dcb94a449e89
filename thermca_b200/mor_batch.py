"""Batched MOR ensemble (BASELINE config #3): B load cases of one MOR body in lockstep.

Host wrapper of ``tc_mor_batch_steps`` (csrc/mor_batch.cu).  Per scenario b the kernel
computes exactly what the reference computes for one MOR part per right-hand side
(thermca/solver.py:282-299 with A = A_body + sum_f film_f A_films[f], mor_io.py:90-92),
with the film coefficients of scenario b following
``film_f,b(t) = h0_f,b * (1 + 0.5 sin(2 pi t / tau_f,b))`` (SURVEY.md §8d cfg#3).
Scenario slices shard over GPUs without communication.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


class MorEnsemble:
    def __init__(self, A_body, A_films, B_T, film_surf, film_h0, film_tau, t_link, q_surf, x0, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("thermca_b200 needs a CUDA device; there is no CPU fallback")
        self.torch = torch
        self.lib = _cabi.load_library()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        A_body = np.ascontiguousarray(A_body, dtype=np.float64)
        self.S, self.r = A_body.shape[0], A_body.shape[1]
        A_films = np.ascontiguousarray(A_films, dtype=np.float64).reshape(-1, self.S, self.r, self.r)
        self.F = A_films.shape[0]
        self.B = int(np.asarray(film_h0).reshape(self.F, -1).shape[1]) if self.F else int(np.asarray(x0).shape[-1])
        film_h0 = np.ascontiguousarray(film_h0, dtype=np.float64).reshape(self.F, self.B)
        up = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.dev)
        self.t = {
            "A_body": up(A_body), "A_films": up(A_films if self.F else np.zeros(1)), "B_T": up(B_T),
            "film_surf": up(film_surf if self.F else [0], np.int32), "film_h0": up(film_h0 if self.F else np.zeros(1)),
            "film_tau": up(np.reshape(film_tau, (self.F, -1)) if self.F else np.ones(1)),
            "t_link": up(t_link if self.F else np.zeros(1)), "q_surf": up(q_surf),
        }
        x0 = np.asarray(x0, dtype=np.float64)
        if x0.ndim == 2:                      # (S, r) -> broadcast over scenarios
            x0 = np.repeat(x0[:, :, None], self.B, axis=2)
        self.X = up(x0.reshape(self.S, self.r, self.B))
        nwork = int(self.lib.tc_mor_batch_work_doubles(self.S, self.r, self.F, self.B))
        self.work = torch.empty(nwork, dtype=torch.float64, device=self.dev)
        self.time = 0.0
        # fused persistent kernel (csrc/mor_fused.cu): operators packed once into DMMA fragment order
        self.fused = bool(self.lib.tc_mor_fused_supported(self.S, self.r, self.F, self.B))
        if self.fused:
            st = torch.cuda.current_stream(self.dev)
            self.packed = torch.empty(int(self.lib.tc_mor_fused_packed_doubles(self.S, self.r, self.F)),
                                      dtype=torch.float64, device=self.dev)
            self.fwork = torch.empty(int(self.lib.tc_mor_fused_work_bytes(self.S, self.r, self.F, self.B)),
                                     dtype=torch.uint8, device=self.dev)
            _cabi.check(self.lib.tc_mor_fused_pack(C.c_void_p(st.cuda_stream), self.S, self.r, self.F,
                                                   C.c_void_p(self.t["A_body"].data_ptr()),
                                                   C.c_void_p(self.t["A_films"].data_ptr()),
                                                   C.c_void_p(self.packed.data_ptr())))
        torch.cuda.synchronize(self.dev)

    @classmethod
    def from_part(cls, part, link_temps, surf_heat, film_h0, film_tau, device=None):
        """Ensemble over ONE MOR part of a lowered network (``spec["mor"][k]``, i.e. the reference's
        ``MORIO`` operators): ``link_temps[f]`` = temperature of the node behind film f, ``surf_heat[s]`` = heat
        flow into surface s (W), ``film_h0 / film_tau`` (F x B) = the per-scenario film laws.  The input term
        follows thermca/solver.py:282-292: flux_s = heat_s / area_s + sum_f film_f (T_link,f - T_ref)."""
        S, r = int(part["S"]), int(part["r"])
        q_surf = np.asarray(surf_heat, dtype=np.float64) / np.asarray(part["surf_areas"], dtype=np.float64)
        t_link = np.asarray(link_temps, dtype=np.float64) - float(part["ref_temp"])
        ens = cls(part["A_body"], part["A_films"], part["B_T"], np.asarray(part["film_surf_idxs"], dtype=np.int32),
                  film_h0, film_tau, t_link, q_surf, np.asarray(part["x0"], dtype=np.float64).reshape(S, r),
                  device=device)
        ens.C_surfs = np.asarray(part["C_surfs"], dtype=np.float64)
        ens.ref_temp = float(part["ref_temp"])
        return ens

    def surface_temps(self):
        """Mean surface temperatures (S x B) of every scenario: sum_s C_surfs[s]^T x_s + T_ref
        (thermca/solver.py:216-220)."""
        X = self.state()                                       # (S, r, B)
        return np.einsum("srq,srb->qb", self.C_surfs, X) + self.ref_temp

    def steps(self, nsteps, h, use_dmma=3, stream=None):
        """use_dmma: 3 = fused persistent DMMA kernel (default; shapes it does not cover — F > 2 or r > 200 — run
        mode 2), 2 = one pipelined DMMA kernel per RK4 stage, 1 = simple DMMA kernel, 0 = FMA cross-check."""
        torch = self.torch
        st = torch.cuda.current_stream(self.dev) if stream is None else stream
        p = lambda a: C.c_void_p(a.data_ptr())
        t = self.t
        self.last_mode = use_dmma
        if use_dmma == 3 and not self.fused:
            use_dmma = self.last_mode = 2
        if use_dmma == 3:
            _cabi.check(self.lib.tc_mor_fused_steps(
                C.c_void_p(st.cuda_stream), self.S, self.r, self.F, self.B, p(self.packed), p(t["B_T"]),
                p(t["film_surf"]), p(t["film_h0"]), p(t["film_tau"]), p(t["t_link"]), p(t["q_surf"]), p(self.X),
                p(self.fwork), float(self.time), float(h), int(nsteps)))
            self.time += nsteps * h
            return
        _cabi.check(self.lib.tc_mor_batch_steps(
            C.c_void_p(st.cuda_stream), self.S, self.r, self.F, self.B, p(t["A_body"]), p(t["A_films"]), p(t["B_T"]),
            p(t["film_surf"]), p(t["film_h0"]), p(t["film_tau"]), p(t["t_link"]), p(t["q_surf"]), p(self.X),
            p(self.work), float(self.time), float(h), int(nsteps), int(use_dmma) if use_dmma in (0, 1, 2) else 1))
        self.time += nsteps * h

    def state(self):
        self.torch.cuda.synchronize(self.dev)
        return self.X.cpu().numpy()

    def launches_per_step(self):
        """Kernels per RK4 step of the per-stage path (mode 2): film table + four stage kernels.  The fused path
        (mode 3) is ONE launch per ``steps`` call, whatever the number of steps."""
        return 5

    def launches(self, nsteps):
        """Kernel launches of one ``steps(nsteps, ...)`` call on the default path."""
        return 1 if self.fused else nsteps * self.launches_per_step()

    def kernel_name(self):
        return f"tc_mor_fused_kernel<{self.F + 1}>" if self.fused else f"tc_mor_stage_pipe_kernel<{self.F + 1}>"

    def flops_per_step(self):
        return 4 * 2.0 * (1 + self.F) * self.r * self.r * self.S * self.B
