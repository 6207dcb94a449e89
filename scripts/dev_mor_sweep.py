"""Dev probe: pipelined MOR stage kernel efficiency vs (r, B) — separates main-loop rate from per-CTA overheads."""
import sys
import numpy as np
import torch
from thermca_b200.mor_batch import MorEnsemble

def run(S, r, F, B, steps=8):
    rng = np.random.default_rng(0)
    A_body = rng.standard_normal((S, r, r)) * 1e-4
    A_films = rng.standard_normal((F, S, r, r)) * 1e-6
    B_T = rng.standard_normal((S, r)) * 1e-3
    h0, tau = rng.uniform(5, 50, (F, B)), rng.uniform(60, 600, (F, B))
    ens = MorEnsemble(A_body, A_films, B_T, np.arange(F, dtype=np.int32) % S, h0, tau, np.linspace(-5, 3, F),
                      np.linspace(100, 0, S), np.zeros((S, r)))
    ens.steps(4, 0.5, use_dmma=2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    ens.steps(steps, 0.5, use_dmma=2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    tf = ens.flops_per_step() / (ms * 1e-3) / 1e12
    print(f"S={S} r={r} F={F} B={B}: {ms:.4f} ms/step, {tf:.2f} TFLOP/s", flush=True)

if __name__ == "__main__":
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    torch.matmul(a, a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): torch.matmul(a, a)
    e1.record(); torch.cuda.synchronize()
    print("dgemm TF", 5 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    del a
    for (S, r, F, B) in [(3, 200, 2, 4096), (3, 192, 2, 4096), (3, 256, 2, 4096), (3, 1024, 2, 4096), (3, 1024, 2, 1024),
                         (3, 192, 2, 6336), (3, 200, 2, 16384), (3, 200, 0, 4096), (3, 200, 1, 4096)]:
        run(S, r, F, B)
