"""Stand-in for scikit-fem 10.0.2 restricted to what thermca calls.

TEST INFRASTRUCTURE ONLY.  Restates P1 tetrahedral assembly as used by
``FESystem.fem_matrices_from_skfem`` (thermca/fem/fe_system.py:322-400):
``MeshTet(p, t[, boundaries])`` (.p, .facets, .boundaries), ``ElementTetP1``,
``Basis(mesh, elem)`` (.nodal_dofs, .get_dofs(name).flatten()),
``FacetBasis(mesh, elem, facets=)`` and ``asm(form, basis)`` for
form in {laplace, mass, unit_load}.  Published P1 formulas:
laplace_ij = vol * grad(lam_i).grad(lam_j);  mass_ij = vol/20 * (1 + delta_ij);
unit_load_i (facet) = area/3.
"""
import numpy as np
import scipy.sparse as sp


class ElementTetP1:
    pass


class MeshTet:
    def __init__(self, p, t, boundaries=None):
        self.p = np.ascontiguousarray(p, dtype=np.float64)            # (3, nverts)
        self.t = np.sort(np.asarray(t, dtype=np.int64), axis=0)       # (4, ntets), sorted like skfem
        f = np.hstack([self.t[[0, 1, 2]], self.t[[0, 1, 3]], self.t[[0, 2, 3]], self.t[[1, 2, 3]]])
        self.facets = np.unique(np.sort(f, axis=0), axis=1)           # (3, nfacets)
        self.boundaries = {k: np.asarray(v) for k, v in (boundaries or {}).items()}


class _Dofs:
    def __init__(self, idx):
        self._idx = idx

    def flatten(self):
        return self._idx


class Basis:
    def __init__(self, mesh, elem):
        self.mesh = mesh
        self.nodal_dofs = np.arange(mesh.p.shape[1], dtype=np.int64)[None, :]

    def get_dofs(self, name):
        fac = self.mesh.facets[:, self.mesh.boundaries[name]]
        return _Dofs(np.unique(fac.ravel()))


class FacetBasis:
    def __init__(self, mesh, elem, facets=None):
        self.mesh = mesh
        self.facets = np.asarray(facets)


def _tet_geometry(mesh):
    P = mesh.p.T[mesh.t.T]                       # (nt, 4, 3)
    J = P[:, 1:, :] - P[:, :1, :]                # rows = edge vectors
    det = np.linalg.det(J)
    vol = np.abs(det) / 6.0
    Jinv = np.linalg.inv(J)                      # columns = grad lam_1..3
    g = np.empty((len(P), 4, 3))
    g[:, 1:, :] = np.transpose(Jinv, (0, 2, 1))
    g[:, 0, :] = -g[:, 1:, :].sum(axis=1)
    return vol, g


def asm(form, basis):
    mesh = basis.mesh
    n = mesh.p.shape[1]
    if form is laplace or form is mass:
        vol, g = _tet_geometry(mesh)
        if form is laplace:
            Ke = vol[:, None, None] * np.einsum("eik,ejk->eij", g, g)
        else:
            Ke = vol[:, None, None] / 20.0 * (np.ones((4, 4)) + np.eye(4))[None]
        tt = mesh.t.T
        rows = np.repeat(tt, 4, axis=1).ravel()
        cols = np.tile(tt, (1, 4)).ravel()
        return sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    if form is unit_load:
        fac = mesh.facets[:, basis.facets].T      # (nf, 3)
        P = mesh.p.T[fac]
        area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
        out = np.zeros(n)
        np.add.at(out, fac.ravel(), np.repeat(area / 3.0, 3))
        return out
    raise NotImplementedError(form)


from .models.poisson import laplace, mass, unit_load  # noqa: E402
