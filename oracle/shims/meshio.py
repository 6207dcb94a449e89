"""Stand-in for meshio: Gmsh 4.1 ASCII reader only (thermca/mesh.py:533).

TEST INFRASTRUCTURE ONLY.  Enough of meshio's object model for
``Mesh._read_gmsh_msh`` (thermca/mesh.py:526-558): ``points``, ``cells`` (blocks with
``.type`` / ``.data``), ``cell_data['gmsh:physical']`` and ``field_data``.
"""
from dataclasses import dataclass
import numpy as np


@dataclass
class CellBlock:
    type: str
    data: np.ndarray


class Mesh:
    def __init__(self, points, cells, cell_data=None, field_data=None, **kw):
        self.points = points
        self.cells = cells
        self.cell_data = cell_data or {}
        self.field_data = field_data or {}


_GMSH_TYPES = {1: ("line", 2), 2: ("triangle", 3), 4: ("tetra", 4), 15: ("vertex", 1)}


def _sections(lines):
    secs, name, buf = {}, None, []
    for ln in lines:
        ln = ln.strip()
        if ln.startswith("$End"):
            secs[name] = buf
            name, buf = None, []
        elif ln.startswith("$"):
            name, buf = ln[1:], []
        elif name is not None:
            buf.append(ln)
    return secs


def read(file_name, file_format=None):
    with open(file_name) as f:
        secs = _sections(f.readlines())
    ver = secs["MeshFormat"][0].split()
    if not ver[0].startswith("4.1") or ver[1] != "0":
        raise NotImplementedError("meshio stand-in reads Gmsh 4.1 ASCII only")
    field_data = {}
    for ln in secs.get("PhysicalNames", [])[1:]:
        dim, tag, name = ln.split(maxsplit=2)
        field_data[name.strip('"')] = np.array([int(tag), int(dim)])
    # entity -> first physical tag
    ent = secs["Entities"]
    npts, ncur, nsur, nvol = (int(v) for v in ent[0].split())
    phys = {}
    row = 1
    for dim, cnt in enumerate((npts, ncur, nsur, nvol)):
        for _ in range(cnt):
            tok = ent[row].split()
            row += 1
            tag = int(tok[0])
            off = 4 if dim == 0 else 7
            nphys = int(tok[off])
            phys[(dim, tag)] = int(tok[off + 1]) if nphys > 0 else 0
    # nodes
    nd = secs["Nodes"]
    nblocks, nnodes, _, _ = (int(v) for v in nd[0].split())
    tags, coords = [], []
    row = 1
    for _ in range(nblocks):
        _, _, parametric, n = (int(v) for v in nd[row].split())
        row += 1
        blk_tags = [int(nd[row + i]) for i in range(n)]
        row += n
        for i in range(n):
            coords.append([float(v) for v in nd[row + i].split()[:3]])
        row += n
        tags.extend(blk_tags)
    tags = np.array(tags)
    points = np.zeros((tags.max(), 3))
    points[tags - 1] = np.array(coords)
    # elements
    el = secs["Elements"]
    nblocks = int(el[0].split()[0])
    row = 1
    cells, phys_data = [], []
    for _ in range(nblocks):
        edim, etag, etype, n = (int(v) for v in el[row].split())
        row += 1
        name, nn = _GMSH_TYPES[etype]
        conn = np.array([[int(v) for v in el[row + i].split()[1:1 + nn]] for i in range(n)],
                        dtype=np.int64).reshape(n, nn) - 1
        row += n
        cells.append(CellBlock(name, conn))
        phys_data.append(np.full(n, phys.get((edim, etag), 0), dtype=np.int64))
    return Mesh(points, cells, {"gmsh:physical": phys_data}, field_data)


def write_points_cells(*a, **k):
    raise NotImplementedError("meshio stand-in: writing is out of scope")
