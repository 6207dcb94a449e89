"""P1 operator build on the device.

``p1_assemble_device`` restates what ``FESystem.fem_matrices_from_skfem``
(thermca/fem/fe_system.py:322-400) gets from scikit-fem for linear tetrahedra — stiffness ``Lx``,
row-sum-lumped mass ``Mx_diag``, per-surface load columns ``Qs`` — with the element pass and the
in-order reduction in csrc/assembly.cu.  ``Qs`` comes back as sparse columns (index, value) instead
of the reference's dense n x S array.  ``device_fe_operator`` goes on to the SELL-32 operator
``A = M^-1 L`` without leaving HBM, so 10^6..10^7-DOF bodies on arbitrary tetrahedral meshes can be
handed to the solver without host assembly; the host restatement it is checked against is
``thermca_b200.synth.p1_operators``.

``install()`` swaps ``FESystem.fem_matrices_from_skfem`` so that ``FEPart(mesh, ...)`` of the unmodified
reference builds its operators here (returned as scipy/numpy objects, which is what ``FESystem.__init__``
expects).

torch provides device memory and the radix sort / prefix sums that order the contributions (index
plumbing); all floating-point arithmetic is in the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


def _dev(device):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("thermca_b200.assembly needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _reduce_by_key(lib, stream, keys, vals):
    """Stable sort by key, then in-order segment sums -> (unique keys, sums)."""
    import torch
    skeys, order = torch.sort(keys, stable=True)
    svals = vals[order]
    del order
    head = torch.ones(skeys.numel(), dtype=torch.bool, device=keys.device)
    if skeys.numel() > 1:
        head[1:] = skeys[1:] != skeys[:-1]
    start = torch.nonzero(head).ravel()
    ukeys = skeys[start]
    nseg = int(start.numel())
    start = torch.cat([start, torch.tensor([skeys.numel()], dtype=torch.int64, device=keys.device)])
    out = torch.empty(nseg, dtype=torch.float64, device=keys.device)
    _cabi.check(lib.tc_segment_sum(stream, nseg, _p(start), _p(svals), _p(out)))
    return ukeys, out


def p1_assemble_device(points, tets, surf_tris, device=None):
    """Geometry-only P1 operators on the device.

    Returns a dict of device tensors: ``indptr`` (int64), ``indices`` (int32, sorted per row),
    ``data`` (Lx), ``Mx_diag``, ``Qs_idx`` / ``Qs_val`` (one pair per surface, ascending DOF order),
    ``surf_areas`` (host numpy) and ``n``.  DOF i is mesh vertex i (fe_system.py:384-387)."""
    import torch
    lib = _cabi.load_library()
    dev = _dev(device)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float64)).to(dev)
    tt = torch.as_tensor(np.ascontiguousarray(tets, dtype=np.int32)).to(dev)
    n, nt = int(pts.shape[0]), int(tt.shape[0])
    if tt.ndim != 2 or tt.shape[1] != 4:
        raise ValueError("tets must have shape (n_tets, 4)")
    if nt and (int(tt.min()) < 0 or int(tt.max()) >= n):
        raise ValueError("tetrahedron refers to a vertex outside the point array")
    keys = torch.empty(16 * nt, dtype=torch.int64, device=dev)
    vals = torch.empty(16 * nt, dtype=torch.float64, device=dev)
    mkeys = torch.empty(4 * nt, dtype=torch.int32, device=dev)
    mvals = torch.empty(4 * nt, dtype=torch.float64, device=dev)
    _cabi.check(lib.tc_p1_tet_contributions(stream, n, nt, _p(pts), _p(tt), _p(keys), _p(vals), _p(mkeys), _p(mvals)))
    ukeys, data = _reduce_by_key(lib, stream, keys, vals)
    del keys, vals
    rows = torch.div(ukeys, n, rounding_mode="floor")
    indices = (ukeys - rows * n).to(torch.int32)
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=indptr[1:])
    mk, mv = _reduce_by_key(lib, stream, mkeys.to(torch.int64), mvals)
    Mx = torch.zeros(n, dtype=torch.float64, device=dev)
    Mx[mk] = mv
    Qi, Qv, areas = [], [], []
    for tris in surf_tris:
        tr = torch.as_tensor(np.ascontiguousarray(tris, dtype=np.int32)).to(dev)
        ntri = int(tr.shape[0])
        k = torch.empty(3 * ntri, dtype=torch.int32, device=dev)
        v = torch.empty(3 * ntri, dtype=torch.float64, device=dev)
        _cabi.check(lib.tc_p1_tri_loads(stream, ntri, _p(pts), _p(tr), _p(k), _p(v)))
        uk, uv = _reduce_by_key(lib, stream, k.to(torch.int64), v)
        Qi.append(uk)
        Qv.append(uv)
        areas.append(float(uv.sum().item()))
    torch.cuda.synchronize(dev)
    return {"n": n, "indptr": indptr, "indices": indices, "data": data, "Mx_diag": Mx, "Qs_idx": Qi, "Qs_val": Qv,
            "surf_areas": np.array(areas)}


def device_fe_operator(points, tets, surf_tris, condy, vol_capy, device=None):
    """SELL-32 operator ``A = M^-1 L`` of a P1 body, built and kept on the device.  Same dict layout as
    ``workloads.device_cuboid_operator`` plus the sparse load columns."""
    import torch
    lib = _cabi.load_library()
    dev = _dev(device)
    asm = p1_assemble_device(points, tets, surf_tris, device=dev)
    n = asm["n"]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    indptr, indices = asm["indptr"], asm["indices"]
    lens = indptr[1:] - indptr[:-1]
    nsl = (n + 31) // 32
    lens_pad = torch.zeros(nsl * 32, dtype=torch.int64, device=dev)
    lens_pad[:n] = lens
    width = lens_pad.view(nsl, 32).max(dim=1).values
    slice_ptr = torch.zeros(nsl + 1, dtype=torch.int64, device=dev)
    torch.cumsum(width * 32, 0, out=slice_ptr[1:])
    total = int(slice_ptr[-1].item())
    rows = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), lens)
    band = int((indices.to(torch.int64) - rows).abs().max().item()) if indices.numel() else 0
    del rows
    col = torch.empty(total, dtype=torch.int32, device=dev)
    val = torch.empty(total, dtype=torch.float64, device=dev)
    col16 = torch.empty(total, dtype=torch.int16, device=dev) if band + 32 < 32768 else None
    mass = vol_capy * asm["Mx_diag"]
    row_scale = 1.0 / mass
    adiag = torch.empty(n, dtype=torch.float64, device=dev)
    _cabi.check(lib.tc_csr_to_sell(stream, n, nsl, _p(indptr), _p(indices), _p(asm["data"]), _p(row_scale), float(condy),
                                   _p(slice_ptr), _p(col), _p(col16), _p(val), _p(adiag)))
    torch.cuda.synchronize(dev)
    return {"n": n, "nslices": nsl, "n_global": n, "row_begin": 0, "slice_ptr": slice_ptr, "col": col, "col16": col16,
            "val": val, "mass": mass, "adiag": adiag, "Qs_idx": asm["Qs_idx"], "Qs_val": asm["Qs_val"],
            "surf_areas": asm["surf_areas"], "csr": asm}


def fem_matrices_device(body_mesh, surf_meshes):
    """Drop-in for ``FESystem.fem_matrices_from_skfem`` (thermca/fem/fe_system.py:322-400): same return
    tuple (numpy / scipy objects), operators assembled on the GPU."""
    import scipy.sparse as sp
    points = np.asarray(body_mesh.points, dtype=np.float64)
    tets = np.asarray(body_mesh.cell_blocks[0])
    asm = p1_assemble_device(points, tets, [np.asarray(c) for c in surf_meshes.cell_blocks])
    n = asm["n"]
    Lx = sp.csr_matrix((asm["data"].cpu().numpy(), asm["indices"].cpu().numpy(), asm["indptr"].cpu().numpy()),
                       shape=(n, n))
    Mx_diag = asm["Mx_diag"].cpu().numpy()
    S = len(asm["Qs_idx"])
    Qs = np.zeros((n, S))
    surf_dof_idxs = []
    for s in range(S):
        idx = asm["Qs_idx"][s].cpu().numpy()
        Qs[idx, s] = asm["Qs_val"][s].cpu().numpy()
        surf_dof_idxs.append(idx)                        # sorted unique vertex ids (fe_system.py:388-390)
    vert_to_dof_idx = np.arange(n)
    dof_to_vert_idx = np.arange(n)
    return Mx_diag, Lx, Qs, vert_to_dof_idx, dof_to_vert_idx, n, surf_dof_idxs, Qs.sum(axis=0)


_original = {}


def install():
    """Route ``FESystem(...)`` operator assembly (fe_system.py:114-125) through the device."""
    import thermca.fem.fe_system as fes
    if "skfem" not in _original:
        _original["skfem"] = fes.FESystem.__dict__["fem_matrices_from_skfem"]
    fes.FESystem.fem_matrices_from_skfem = staticmethod(fem_matrices_device)


def uninstall():
    if "skfem" in _original:
        import thermca.fem.fe_system as fes
        fes.FESystem.fem_matrices_from_skfem = _original.pop("skfem")
