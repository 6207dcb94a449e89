"""Stand-in for the third-party ``sparse`` (pydata/sparse 0.15) package.

TEST INFRASTRUCTURE ONLY. Lets the unmodified reference at /root/reference be
imported in this container, where pydata/sparse is not installed.  Only the
subset of the API that thermca touches is provided:
thermca/network.py:11,146,547  thermca/solver.py:15,408
thermca/resultdata.py:19,745-779  thermca/lpm/lp_cube.py:359-386
thermca/lpm/lp_cyl.py:272-298  thermca/lpm/lp_construction.py:338-396
thermca/lpm/lp_system.py:328-337  thermca/_utils/sparse_dok_indexer.py:21-22

``DOK`` keeps an insertion-ordered dict; ``to_coo`` returns coordinates in
lexicographic order (the LP assembly turns ``geom_ratio.to_coo().tocsr()`` into
the part's CSR operator).
"""
import numpy as np
import scipy.sparse as _sp

from . import _slicing  # noqa: F401


class SparseArray:
    pass


class COO(SparseArray):
    def __init__(self, coords, data=None, shape=None, fill_value=0.0):
        coords = np.asarray(coords, dtype=np.int64)
        if coords.ndim == 1:
            coords = coords[None, :]
        data = np.asarray(data)
        if data.ndim == 0:
            data = np.full(coords.shape[1], data)
        if shape is None:
            shape = tuple(int(c.max()) + 1 if c.size else 0 for c in coords)
        self.shape = tuple(int(s) for s in shape)
        self.fill_value = fill_value
        # canonical: lexicographically sorted coordinates
        if coords.shape[1] > 0:
            order = np.lexsort(coords[::-1])
            coords, data = coords[:, order], data[order]
        self.coords = coords
        self.data = data
        self.dtype = data.dtype

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nnz(self):
        return self.data.size

    def nonzero(self):
        mask = self.data != 0
        return tuple(c[mask] for c in self.coords)

    def reshape(self, shape):
        lin = np.ravel_multi_index(tuple(self.coords), self.shape) if self.nnz else np.zeros(0, np.int64)
        new = np.array(np.unravel_index(lin, shape)) if self.nnz else np.zeros((len(shape), 0), np.int64)
        return COO(new, self.data.copy(), shape, self.fill_value)

    def tocsr(self):
        assert self.ndim == 2
        return _sp.coo_matrix((self.data, (self.coords[0], self.coords[1])), shape=self.shape).tocsr()

    def todense(self):
        out = np.full(self.shape, self.fill_value, dtype=self.dtype if self.nnz else np.float64)
        if self.nnz:
            out[tuple(self.coords)] = self.data
        return out

    @classmethod
    def from_scipy_sparse(cls, m):
        m = m.tocoo()
        return cls(np.array([m.row, m.col]), m.data.copy(), m.shape)

    @classmethod
    def from_numpy(cls, a):
        a = np.asarray(a)
        nz = np.nonzero(a)
        return cls(np.array(nz), a[nz], a.shape)

    def _as_dict(self):
        return {tuple(int(i) for i in k): v for k, v in zip(self.coords.T, self.data)}

    def __mul__(self, other):
        if isinstance(other, COO):
            a, b = self._as_dict(), other._as_dict()
            keys = [k for k in a if k in b]
            coords = np.array(keys, dtype=np.int64).T.reshape(self.ndim, -1)
            data = np.array([a[k] * b[k] for k in keys], dtype=np.float64)
            return COO(coords, data, self.shape)
        return COO(self.coords.copy(), self.data * other, self.shape)

    __rmul__ = __mul__

    def sum(self, axis=None):
        if axis is None:
            return self.data.sum()
        keep = [i for i in range(self.ndim) if i != axis]
        acc = {}
        for k, v in zip(self.coords.T, self.data):
            kk = tuple(int(k[i]) for i in keep)
            acc[kk] = acc.get(kk, 0.0) + v
        coords = np.array(list(acc.keys()), dtype=np.int64).T.reshape(len(keep), -1)
        return COO(coords, np.array(list(acc.values()), dtype=np.float64),
                   tuple(self.shape[i] for i in keep))


class DOK(SparseArray):
    def __init__(self, shape, data=None, dtype=None, fill_value=None):
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype if dtype is not None else np.float64)
        self.fill_value = self.dtype.type(0 if fill_value is None else fill_value)
        self.data = dict(data) if data else {}

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nnz(self):
        return len(self.data)

    def to_coo(self):
        if self.data:
            coords = np.array(list(self.data.keys()), dtype=np.int64).T.reshape(self.ndim, -1)
            data = np.array(list(self.data.values()), dtype=self.dtype)
        else:
            coords = np.zeros((self.ndim, 0), dtype=np.int64)
            data = np.zeros(0, dtype=self.dtype)
        return COO(coords, data, self.shape, self.fill_value)

    @classmethod
    def from_coo(cls, coo):
        d = cls(coo.shape, dtype=coo.dtype if coo.nnz else np.float64, fill_value=coo.fill_value)
        d.data = coo._as_dict()
        return d

    def todense(self):
        return self.to_coo().todense()

    def _norm_key(self, key):
        if not isinstance(key, tuple):
            key = (key,)
        return tuple(int(k) for k in key)

    def __getitem__(self, key):
        return self.data.get(self._norm_key(key), self.fill_value)

    def __setitem__(self, key, value):
        key = self._norm_key(key)
        if value == self.fill_value:
            self.data.pop(key, None)
        else:
            self.data[key] = value


def tril(x, k=0):
    coo = x.to_coo() if isinstance(x, DOK) else x
    mask = coo.coords[-2] + k >= coo.coords[-1]
    return COO(coo.coords[:, mask], coo.data[mask], coo.shape)
