"""Stage the UNMODIFIED reference package under oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The reference is pure Python, so "building" it means placing its package where the GPU box can
import it: ``oracle/_ref/`` is git-ignored (the sources never enter this repository's history) but
not gpurun-ignored, so the staged tree travels to the B200 box with the snapshot — the same route
the in-tree ``.so`` files take.  There the ``-m gpu`` seam tests run the reference's own
``Network(model).sim(...)`` on top of the CUDA kernels (``thermca_b200.install()``) and, in the same
process, through the reference's own scipy/numba solver, and compare ``Result[...]``.

What is staged: ``/root/reference/thermca`` without the HDF5 ``.med`` meshes (unreadable without
``meshio``'s h5py backend) plus a small manifest (file count, sha256 of the concatenated sources) so a
stale copy is detected.  Third-party modules missing in this image are served by ``oracle/shims``
(tracked, written for this repo).  Nothing under ``thermca_b200/`` may import from here.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/thermca"
OUT = os.path.join(HERE, "_ref")
DST = os.path.join(OUT, "thermca")
MANIFEST = os.path.join(OUT, "MANIFEST.json")
SKIP_EXT = (".med", ".pyc")


def _tree_digest(root):
    h = hashlib.sha256()
    count = 0
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames[:] = sorted(d for d in dirnames if d != "__pycache__" and not d.startswith("."))
        for fn in sorted(filenames):
            if fn.endswith(SKIP_EXT):
                continue
            p = os.path.join(dirpath, fn)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as f:
                h.update(f.read())
            count += 1
    return h.hexdigest(), count


def staged():
    return os.path.isfile(os.path.join(DST, "solver.py"))


def build(force=False):
    """Copy the reference package into oracle/_ref/thermca.  No-op (returns the staged path, or None) where
    /root/reference does not exist — the GPU box only uses the prebuilt copy."""
    if not os.path.isdir(SRC):
        return DST if staged() else None
    digest, count = _tree_digest(SRC)
    if not force and staged() and os.path.isfile(MANIFEST):
        try:
            with open(MANIFEST) as f:
                if json.load(f).get("sha256") == digest:
                    return DST
        except Exception:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(OUT, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("*.med", "*.pyc", "__pycache__", ".thermca_cache"))
    for dirpath, _, filenames in os.walk(DST):          # the source tree is read-only; the copy must be removable
        os.chmod(dirpath, 0o755)
        for fn in filenames:
            os.chmod(os.path.join(dirpath, fn), 0o644)
    with open(MANIFEST, "w") as f:
        json.dump({"source": SRC, "files": count, "sha256": digest,
                   "note": "unmodified copy of the reference package; test infrastructure, git-ignored"}, f, indent=1)
    return DST


if __name__ == "__main__":
    print(build(force=True))
