"""thermca_b200 — B200-native (sm_100a) transient solve for steffenschroe/Thermca.

Drop-in for ``thermca.solver.transient_solution`` (see ``thermca_b200.solver.install``); the hot
path runs in hand-written CUDA kernels behind the C ABI of ``include/thermca_b200.h``.
There is no CPU fallback.
"""
__version__ = "0.1.0"

__all__ = ["solver", "device", "lowering", "trace", "spec", "synth", "workloads", "dist", "mor_batch", "build"]
