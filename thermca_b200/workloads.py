"""Device-resident synthetic workloads of the BASELINE configs (bench + large-size tests).

``device_cuboid_spec`` builds BASELINE config #2/#5 (one full-order FE cuboid, 1 kW on the
bottom face, 10 W/m2K film on the other faces to a 20 C BoundNode) with the operator
generated directly in HBM by csrc/synth.cu, so 10^7 DOF need no host assembly.  The small
host restatement it is checked against is thermca_b200/synth.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .synth import c_surf_columns, single_part_network


def cuboid_dims(target_dof, aspect=(0.5, 1.0, 0.05)):
    """Cell counts (nx, ny, nz) with roughly cubic cells and about ``target_dof`` vertices."""
    lx, ly, lz = aspect
    h = (lx * ly * lz / target_dof) ** (1.0 / 3.0)
    nx, ny, nz = (max(2, int(round(l / h))) for l in (lx, ly, lz))
    return nx, ny, nz


def device_cuboid_operator(nx, ny, nz, lx=0.5, ly=1.0, lz=0.05, jitter=0.15, condy=50.0,
                           vol_capy=8000.0 * 500.0, row_begin=0, row_end=None, device=None):
    """SELL-32 operator rows [row_begin, row_end) of the Kuhn cuboid, generated on the device."""
    import torch
    lib = _cabi.load_library()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = (nx + 1) * (ny + 1) * (nz + 1)
    row_end = n if row_end is None else row_end
    nloc = row_end - row_begin
    nsl = (nloc + 31) // 32
    t = {
        "slice_ptr": torch.empty(nsl + 1, dtype=torch.int64, device=dev),
        "col": torch.empty(nsl * 32 * 15, dtype=torch.int32, device=dev),
        "col16": torch.empty(nsl * 32 * 15, dtype=torch.int16, device=dev),
        "val": torch.empty(nsl * 32 * 15, dtype=torch.float64, device=dev),
        "mass": torch.empty(nloc, dtype=torch.float64, device=dev),
        "adiag": torch.empty(nloc, dtype=torch.float64, device=dev),
        "q_bottom": torch.empty(nloc, dtype=torch.float64, device=dev),
        "q_upper": torch.empty(nloc, dtype=torch.float64, device=dev),
    }
    torch.cuda.synchronize(dev)
    p = lambda a: C.c_void_p(a.data_ptr())
    _cabi.check(lib.tc_synth_cuboid(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), nx, ny, nz,
                                    lx, ly, lz, jitter, condy, vol_capy, row_begin, row_end,
                                    p(t["slice_ptr"]), p(t["col"]), p(t["col16"]), p(t["val"]), p(t["mass"]),
                                    p(t["adiag"]), p(t["q_bottom"]), p(t["q_upper"])))
    if (nx + 1) * (nz + 1) + (nz + 1) + 1 >= 32768:
        t["col16"] = None                       # band too wide for 16-bit relative columns
    t.update({"n": nloc, "nslices": nsl, "n_global": n, "row_begin": row_begin})
    return t


def device_cuboid_spec(nx, ny, nz, lx=0.5, ly=1.0, lz=0.05, jitter=0.15, heat=1000.0, film=10.0,
                       env_temp=20.0, init_temp=20.0, condy=50.0, vol_capy=8000.0 * 500.0, device=None):
    """Lowered network of BASELINE config #2 whose FE operator already lives in HBM."""
    import torch
    op = device_cuboid_operator(nx, ny, nz, lx, ly, lz, jitter, condy, vol_capy, device=device)
    n = op["n"]
    mass = op["mass"]
    minv = 1.0 / mass
    Qi, Qv = [], []
    for key in ("q_bottom", "q_upper"):
        q = op[key]
        nz_idx = torch.nonzero(q).ravel()
        Qi.append(nz_idx.cpu().numpy().astype(np.int64))
        Qv.append(q[nz_idx].cpu().numpy())
    areas = np.array([v.sum() for v in Qv])
    minv_h = [minv[torch.from_numpy(ix).to(minv.device)].cpu().numpy() for ix in Qi]
    # film of surface 1 folded into the operator diagonal (thermca/fem/ffe_io.py:57-60)
    ix1 = torch.from_numpy(Qi[1]).to(mass.device)
    add = film * torch.from_numpy(minv_h[1] * Qv[1]).to(mass.device)
    rows = ix1
    # position of the diagonal slot (offset (0,0,0) is slot 7 of the 15 ordered offsets)
    pos = (rows >> 5) * (32 * 15) + 7 * 32 + (rows & 31)
    op["val"][pos] += add
    op["adiag"][rows] += add
    part = {
        "dof": np.int64(n), "init_temp": np.float64(init_temp),
        "A_body": None, "A_data": None, "M_inv": None, "sell": op,
        "B_idx": Qi, "B_val": [m * q for m, q in zip(minv_h, Qv)],
        "C_idx": Qi, "C_val": c_surf_columns(Qv),
        "surf_areas": areas, "S": np.int64(2), "films": np.array([film]),
        "film_surf_idxs": np.array([1], dtype=np.int64), "surf_net_idxs": np.array([1, 2], dtype=np.int64),
        "link_elem_net_idxs": np.array([0], dtype=np.int64), "film_didx": [], "film_adata": [],
        "dock_areas": areas[[1]], "var_film_rc_idx": None, "tdep": None,
    }
    spec = single_part_network(part, heat, film, env_temp)
    spec["_dims"] = (nx, ny, nz)
    return spec


def sell_to_csr_host(op):
    """Host CSR (scipy) of a device SELL operator with fixed width 15 (padding dropped)."""
    import scipy.sparse as sp
    n = op["n"]
    nsl = op["nslices"]
    col = op["col"].cpu().numpy().reshape(nsl, 15, 32).transpose(0, 2, 1).reshape(nsl * 32, 15)[:n]
    val = op["val"].cpu().numpy().reshape(nsl, 15, 32).transpose(0, 2, 1).reshape(nsl * 32, 15)[:n]
    rows = np.repeat(np.arange(n), 15).reshape(n, 15)
    keep = ~((val == 0.0) & (col == rows))
    m = sp.csr_matrix((val[keep], (rows[keep], col[keep])), shape=(n, op["n_global"]))
    return m


def device_mesh_spec(points, tets, tris_heat, tris_film, heat=1000.0, film=10.0, env_temp=20.0, init_temp=20.0,
                     condy=50.0, vol_capy=8000.0 * 500.0, device=None):
    """The BASELINE config #2 network (heat on surface 0, constant film on surface 1 to a BoundNode) on an
    ARBITRARY tetrahedral mesh, with the operator assembled on the device (thermca_b200/assembly.py)."""
    import torch
    from .assembly import device_fe_operator
    op = device_fe_operator(points, tets, [tris_heat, tris_film], condy, vol_capy, device=device)
    n = op["n"]
    csr = op.pop("csr")
    mass = op["mass"]
    Qi = [ix.cpu().numpy().astype(np.int64) for ix in op["Qs_idx"]]
    Qv = [v.cpu().numpy() for v in op["Qs_val"]]
    minv_h = [(1.0 / mass[ix]).cpu().numpy() for ix in op["Qs_idx"]]
    # film of surface 1 folded into the stored diagonal (thermca/fem/ffe_io.py:57-60)
    rows = op["Qs_idx"][1]
    add = film * torch.from_numpy(minv_h[1] * Qv[1]).to(mass.device)
    indptr, indices = csr["indptr"], csr["indices"]
    lens = indptr[1:] - indptr[:-1]
    row_of = torch.repeat_interleave(torch.arange(n, device=mass.device, dtype=torch.int64), lens)
    dpos = torch.nonzero(indices.to(torch.int64) == row_of).ravel()          # one stored diagonal per row
    if int(dpos.numel()) != n:
        raise ValueError("mesh has a vertex that belongs to no tetrahedron")
    k_diag = dpos[rows] - indptr[rows]
    pos = op["slice_ptr"][rows >> 5] + 32 * k_diag + (rows & 31)
    op["val"][pos] += add
    op["adiag"][rows] += add
    areas = op["surf_areas"]
    part = {
        "dof": np.int64(n), "init_temp": np.float64(init_temp),
        "A_body": None, "A_data": None, "M_inv": None, "sell": op,
        "B_idx": Qi, "B_val": [m * q for m, q in zip(minv_h, Qv)],
        "C_idx": Qi, "C_val": c_surf_columns(Qv),
        "surf_areas": areas, "S": np.int64(2), "films": np.array([film]),
        "film_surf_idxs": np.array([1], dtype=np.int64), "surf_net_idxs": np.array([1, 2], dtype=np.int64),
        "link_elem_net_idxs": np.array([0], dtype=np.int64), "film_didx": [], "film_adata": [],
        "dock_areas": areas[[1]], "var_film_rc_idx": None, "tdep": None,
    }
    return single_part_network(part, heat, film, env_temp)
