"""thermca_b200 — B200-native (sm_100a) transient solve for steffenschroe/Thermca.

Drop-in for ``thermca.solver.transient_solution`` (see ``thermca_b200.solver.install``); the hot
path runs in hand-written CUDA kernels behind the C ABI of ``include/thermca_b200.h``.
There is no CPU fallback.

``thermca_b200.install()`` switches the transient solve and, optionally, the adjacent one-off
computations (steady-state solves, P1 operator assembly, MOR basis generation) of an unmodified
Thermca installation to the device; ``uninstall()`` restores the reference's own code.
"""
__version__ = "0.1.0"

__all__ = ["solver", "device", "lowering", "trace", "spec", "synth", "workloads", "dist", "mor_batch", "build",
           "stationary", "assembly", "mor_basis", "linsolve", "install", "uninstall"]


def install(transient=True, stationary=True, assembly=True, mor_basis=True):
    """Route the selected computations of the reference through the device (each swaps one name the reference
    reads at call time: thermca/network.py:978, fem/fe_system.py:17,124, lpm/lp_system.py:8, fem/mor_io.py:75)."""
    if transient:
        from . import solver
        solver.install()
    if stationary:
        from . import stationary as st
        st.install()
    if assembly:
        from . import assembly as asm
        asm.install()
    if mor_basis:
        from . import mor_basis as mb
        mb.install()


def uninstall():
    from . import assembly as asm
    from . import mor_basis as mb
    from . import solver
    from . import stationary as st
    solver.uninstall()
    st.uninstall()
    asm.uninstall()
    mb.uninstall()
