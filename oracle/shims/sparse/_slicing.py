"""normalize_index stand-in (used by thermca/_utils/sparse_dok_indexer.py:22,94)."""
from numbers import Integral
import numpy as np


def normalize_index(idx, shape):
    """Expand ``idx`` to one key per dimension; wrap negatives; lists -> 1-d arrays."""
    if not isinstance(idx, tuple):
        idx = (idx,)
    if any(k is Ellipsis for k in idx):
        pos = [i for i, k in enumerate(idx) if k is Ellipsis][0]
        fill = (slice(None),) * (len(shape) - (len(idx) - 1))
        idx = idx[:pos] + fill + idx[pos + 1:]
    idx = idx + (slice(None),) * (len(shape) - len(idx))
    out = []
    for k, n in zip(idx, shape):
        if isinstance(k, slice):
            out.append(k)
        elif isinstance(k, Integral) or (isinstance(k, np.ndarray) and k.ndim == 0):
            k = int(k)
            if k < 0:
                k += n
            if not 0 <= k < n:
                raise IndexError("index out of range")
            out.append(k)
        else:
            a = np.asarray(k)
            if a.dtype == bool:
                a = np.nonzero(a)[0]
            a = a.astype(np.int64)
            a = np.where(a < 0, a + n, a)
            out.append(a)
    return tuple(out)
