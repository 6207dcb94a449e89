"""Synthetic workloads of the BASELINE configs, built without the reference.

* :func:`kuhn_cuboid_mesh` – structured cuboid, every cell split into 6 tetrahedra
  along its main diagonal (Kuhn / Freudenthal), interior vertices touch 14 neighbours
  -> 15 non-zeros per operator row (SURVEY.md §8d cfg#2/#5).
* :func:`p1_operators` – P1 tetrahedral operators as ``FESystem`` derives them from
  scikit-fem (thermca/fem/fe_system.py:322-400): stiffness ``Lx``, row-sum lumped mass
  ``Mx_diag``, surface load columns ``Qs`` and the area-weighted mean extractor
  ``C_surf`` (thermca/fem/fe_system.py:152-160).
* :func:`ffe_part_spec` / :func:`cuboid_spec` – the lowered network a
  ``Network(model)`` with one full-order ``FEPart``, ``HeatSource(bottom, q)`` and
  ``FilmLink(upper_faces, BoundNode(T), h)`` produces (thermca/network.py:116-685,
  thermca/fem/ffe_io.py:24-117), bypassing the quadratic diagonal search of
  ``FESystem.__init__`` (thermca/_utils/sparse.py:13-22) so 10^6+ DOF can be built.

The device-side generator for 10^7 DOF lives in csrc/synth.cu and is checked against
this module at small sizes.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# the six axis permutations of the Kuhn split: path v000 -> +a -> +b -> +c
_KUHN = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]


def vertex_index(i, j, k, nx, ny, nz):
    """z fastest, then x, y slowest (keeps the operator band narrow when ny is largest)."""
    return (j * (nx + 1) + i) * (nz + 1) + k


def jitter_points(x, y, z, lx, ly, lz, hx, hy, hz, jitter):
    """Smooth deterministic warp of the interior vertices (same formula as csrc/synth.cu): the
    normal displacement vanishes on every face, so the cuboid keeps its shape while all 15
    P1 couplings of the Kuhn split become non-zero (on an orthogonal grid 8 of them cancel
    exactly and scipy's sparse product drops them)."""
    pi = 3.14159265358979323846
    xi, et, ze = x / lx, y / ly, z / lz
    xo = x + jitter * hx * np.sin(2 * pi * xi) * np.cos(3 * pi * et) * np.cos(2 * pi * ze)
    yo = y + jitter * hy * np.cos(2 * pi * xi + 0.3) * np.sin(2 * pi * et) * np.cos(pi * ze)
    zo = z + jitter * hz * np.cos(pi * xi) * np.cos(2 * pi * et + 0.7) * np.sin(2 * pi * ze)
    return xo, yo, zo


def kuhn_cuboid_mesh(nx, ny, nz, lx, ly, lz, jitter=0.0):
    """Returns (points (V,3), tets (6*nx*ny*nz,4), tri_bottom, tri_upper).

    ``tri_bottom`` is the z=0 face, ``tri_upper`` the five other faces (the two named
    surfaces of thermca/tests/test_fe_part.py:102-166)."""
    J, I, K = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), np.arange(nz + 1), indexing="ij")
    pts = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
    vid = vertex_index(I, J, K, nx, ny, nz).ravel()
    Ir, Jr, Kr = I.ravel(), J.ravel(), K.ravel()
    hx, hy, hz = lx / nx, ly / ny, lz / nz
    px, py, pz = hx * Ir, hy * Jr, hz * Kr
    if jitter:
        px, py, pz = jitter_points(px, py, pz, lx, ly, lz, hx, hy, hz, jitter)
    # keep the six faces exactly planar
    px = np.where(Ir == 0, 0.0, np.where(Ir == nx, lx, px))
    py = np.where(Jr == 0, 0.0, np.where(Jr == ny, ly, py))
    pz = np.where(Kr == 0, 0.0, np.where(Kr == nz, lz, pz))
    pts[vid, 0], pts[vid, 1], pts[vid, 2] = px, py, pz

    jc, ic, kc = np.meshgrid(np.arange(ny), np.arange(nx), np.arange(nz), indexing="ij")
    ic, jc, kc = ic.ravel(), jc.ravel(), kc.ravel()
    tets = []
    for perm in _KUHN:
        d = np.zeros((4, 3), dtype=np.int64)
        for s, ax in enumerate(perm):
            d[s + 1] = d[s]
            d[s + 1, ax] += 1
        tets.append(np.stack([vertex_index(ic + d[s, 0], jc + d[s, 1], kc + d[s, 2], nx, ny, nz)
                              for s in range(4)], axis=1))
    tets = np.vstack(tets)

    def face(fixed_axis, fixed_val, n_a, n_b):
        a, b = np.meshgrid(np.arange(n_a), np.arange(n_b), indexing="ij")
        a, b = a.ravel(), b.ravel()

        def vid2(da, db):
            c = [None, None, None]
            free = [ax for ax in range(3) if ax != fixed_axis]
            c[fixed_axis] = np.full_like(a, fixed_val)
            c[free[0]], c[free[1]] = a + da, b + db
            return vertex_index(c[0], c[1], c[2], nx, ny, nz)
        v00, v10, v01, v11 = vid2(0, 0), vid2(1, 0), vid2(0, 1), vid2(1, 1)
        return np.vstack([np.stack([v00, v10, v11], 1), np.stack([v00, v01, v11], 1)])

    n = (nx, ny, nz)
    tri_bottom = face(2, 0, nx, ny)
    tri_upper = np.vstack([face(2, nz, nx, ny), face(0, 0, ny, nz), face(0, nx, ny, nz),
                           face(1, 0, nx, nz), face(1, ny, nx, nz)])
    assert n == (nx, ny, nz)
    return pts, tets, tri_bottom, tri_upper


def p1_operators(points, tets, surf_tris, chunk=1 << 20):
    """P1 operators (geometry only): ``Lx`` (csr, sorted indices), ``Mx_diag``,
    ``Qs_idx/Qs_val`` (sparse columns of the load matrix) and ``surf_areas``.

    laplace_ij = vol * grad(lam_i).grad(lam_j); lumped mass = vol/4 per vertex
    (row sum of vol/20*(1+delta_ij), thermca/fem/fe_system.py:379-381);
    unit_load = area/3 per facet vertex (thermca/fem/fe_system.py:373-382)."""
    points = np.asarray(points, dtype=np.float64)
    tets = np.asarray(tets, dtype=np.int64)
    n = len(points)
    Lx = sp.csr_matrix((n, n))
    Mx = np.zeros(n)
    for lo in range(0, len(tets), chunk):
        tt = np.sort(tets[lo: lo + chunk], axis=1)
        P = points[tt]
        Jm = P[:, 1:, :] - P[:, :1, :]
        vol = np.abs(np.linalg.det(Jm)) / 6.0
        Jinv = np.linalg.inv(Jm)
        g = np.empty((len(tt), 4, 3))
        g[:, 1:, :] = np.transpose(Jinv, (0, 2, 1))
        g[:, 0, :] = -g[:, 1:, :].sum(axis=1)
        Ke = vol[:, None, None] * np.einsum("eik,ejk->eij", g, g)
        rows = np.repeat(tt, 4, axis=1).ravel()
        cols = np.tile(tt, (1, 4)).ravel()
        Lx = Lx + sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(n, n)).tocsr()
        np.add.at(Mx, tt.ravel(), np.repeat(vol / 4.0, 4))
    Lx.sum_duplicates()
    Lx.sort_indices()
    Qs_idx, Qs_val, areas = [], [], []
    for tris in surf_tris:
        tris = np.asarray(tris, dtype=np.int64)
        P = points[tris]
        a = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
        q = np.zeros(n)
        np.add.at(q, tris.ravel(), np.repeat(a / 3.0, 3))
        nz = np.nonzero(q)[0]
        Qs_idx.append(nz.astype(np.int64))
        Qs_val.append(q[nz])
        areas.append(q.sum())
    return Lx, Mx, Qs_idx, Qs_val, np.array(areas)


def c_surf_columns(Qs_val):
    """Area-weighted mean extractor per surface (thermca/fem/fe_system.py:152-160)."""
    out = []
    for q in Qs_val:
        num = q.size
        c_mean = q.sum() / num
        out.append((q / c_mean) * 1.0 / num)
    return out


def ffe_part_spec(Lx, Mx_diag, Qs_idx, Qs_val, surf_areas, condy, vol_capy, init_temp,
                  films, film_surf_idxs, surf_net_idxs, link_elem_net_idxs):
    """Lowered full-order FE part, restating ``FFEIO.__init__`` / ``to_state_space``
    (thermca/fem/ffe_io.py:24-65,85-117) on sparse surface columns."""
    n = Lx.shape[0]
    M_diag = vol_capy * Mx_diag
    M_inv = 1.0 / M_diag
    L = sp.csr_matrix(Lx, copy=True)
    L.data = condy * L.data
    A_body = sp.diags(M_inv, format="csr") @ L
    A_body.sort_indices()
    indptr, indices = A_body.indptr.astype(np.int32), A_body.indices.astype(np.int32)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    diag_pos = np.nonzero(rows == indices)[0]          # thermca/fem/ffe_io.py:52
    assert diag_pos.size == n
    A_data = A_body.data.copy()
    film_didx, film_adata = [], []
    for film, s in zip(films, film_surf_idxs):
        idx = Qs_idx[s]
        adata = M_inv[idx] * Qs_val[s]
        film_didx.append(diag_pos[idx].astype(np.int64))
        film_adata.append(adata)
        A_data[diag_pos[idx]] += film * adata
    C_val = c_surf_columns(Qs_val)
    return {
        "dof": np.int64(n), "init_temp": np.float64(init_temp),
        "A_body": {"indptr": indptr, "indices": indices, "data": A_body.data.copy(),
                   "shape": np.asarray(A_body.shape, dtype=np.int64)},
        "A_data": A_data, "M_inv": M_inv,
        "B_idx": [ix.copy() for ix in Qs_idx], "B_val": [M_inv[ix] * qv for ix, qv in zip(Qs_idx, Qs_val)],
        "C_idx": [ix.copy() for ix in Qs_idx], "C_val": C_val,
        "surf_areas": np.asarray(surf_areas, dtype=np.float64), "S": np.int64(len(surf_areas)),
        "films": np.asarray(films, dtype=np.float64),
        "film_surf_idxs": np.asarray(film_surf_idxs, dtype=np.int64),
        "surf_net_idxs": np.asarray(surf_net_idxs, dtype=np.int64),
        "link_elem_net_idxs": np.asarray(link_elem_net_idxs, dtype=np.int64),
        "film_didx": film_didx, "film_adata": film_adata,
        "dock_areas": np.asarray(surf_areas, dtype=np.float64)[np.asarray(film_surf_idxs, dtype=np.int64)],
        "var_film_rc_idx": None, "tdep": None,
    }


def single_part_network(part, heat_on_surf, film_value, env_temp=20.0):
    """Network arrays of: BoundNode(env) | FEPartSurf 0 | FEPartSurf 1, constant film on
    surface 1 to env, constant heat on surface 0 (node order thermca/network.py:56-100;
    pattern thermca/network.py:706-745)."""
    N, u = 3, 0
    area1 = float(part["surf_areas"][1])
    rows = np.array([0, 0, 1, 2, 2], dtype=np.int64)
    cols = np.array([0, 2, 1, 0, 2], dtype=np.int64)
    cond = film_value * area1
    data = np.array([0.0, cond, 0.0, cond, 0.0])
    m = sp.csr_matrix((data, (rows, cols)), shape=(N, N))
    z = lambda: np.zeros(N)
    empty_i = np.zeros(0, dtype=np.int64)
    spec = {
        "N": np.int64(N), "u": np.int64(u),
        "init_temp": np.array([env_temp, 0.0, 0.0]), "heat_src": np.array([0.0, heat_on_surf, 0.0]),
        "capy": np.array([np.inf, 0.0, 0.0]), "vol": z(), "vol_capy": z(), "condy": z(),
        "posn": np.zeros((N, 3)),
        "rc_indptr": m.indptr.astype(np.int32), "rc_indices": m.indices.astype(np.int32),
        "rc_data": m.data.copy(), "rc_clean_data": m.data.copy(),
        "stat_idx": empty_i, "inputs": [],
        "diag_didx": np.zeros(0, dtype=np.int32), "uu_didx": np.zeros(0, dtype=np.int32),
        "uk_didx": np.zeros(0, dtype=np.int32),
        "matl_groups": [], "capy_idxs": empty_i,
        "flow_const": {"didx": empty_i, "i0": empty_i, "i1": empty_i, "vol_flow": np.zeros(0)},
        "flow_funcs": [], "film_funcs": [], "heat_funcs": [], "flux_funcs": [], "bound_funcs": [],
        "lp": [], "ffe": [part], "mor": [], "lp_to_lp": [], "programs": [], "tables": [],
    }
    return spec


def cuboid_spec(nx, ny, nz, lx=0.5, ly=1.0, lz=0.05, heat=1000.0, film=10.0, env_temp=20.0,
                init_temp=20.0, dens=8000.0, spec_heat=500.0, condy=50.0, jitter=0.0):
    """BASELINE config #2 at any size, assembled on the host (use csrc/synth.cu for >2M DOF)."""
    pts, tets, tb, tu = kuhn_cuboid_mesh(nx, ny, nz, lx, ly, lz, jitter)
    Lx, Mx, Qi, Qv, areas = p1_operators(pts, tets, [tb, tu])
    part = ffe_part_spec(Lx, Mx, Qi, Qv, areas, condy, dens * spec_heat, init_temp,
                         films=[film], film_surf_idxs=[1], surf_net_idxs=[1, 2],
                         link_elem_net_idxs=[0])
    return single_part_network(part, heat, film, env_temp)
