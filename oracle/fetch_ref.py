"""Recipe for ``oracle/_ref``: a byte-for-byte copy of the UNMODIFIED reference package  --  test infrastructure.

    python oracle/fetch_ref.py            # build container only (needs /root/reference)

The reference (awkward0 0.15.5) is pure Python: there is nothing to compile, so "building" the reference arm means
copying ``/root/reference/awkward0`` to ``oracle/_ref/awkward0``.  ``oracle/_ref/`` is git-ignored (the sources never
enter this repository's history) but it is NOT gpurun-ignored, so the copy travels to the GPU box, where
``bench.py``'s CPU legs (``cpu_baseline``, the per-workflow ``cpu_ms`` and ``--impl reference``) time the reference
itself on the box's host cores (``kind: "reference"``).  ``MANIFEST.json`` records the sha256 of every copied file;
``verify()`` re-hashes the copy (and, in the build container, the originals) so a modified copy is detected.

The copy is imported through ``oracle/_refloader.py``, which applies the numpy-2 shim of SURVEY.md section 8(c)
BEFORE the import; no reference file is edited.  Only tests/, __graft_entry__ and bench.py's CPU legs may use it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("AWKWARD0_REFERENCE", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(DST_ROOT, "MANIFEST.json")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for block in iter(lambda: f.read(1 << 20), b""):
            h.update(block)
    return h.hexdigest()


def _files(root):
    out = []
    for d, _, names in os.walk(os.path.join(root, "awkward0")):
        if "__pycache__" in d:
            continue
        for name in names:
            if name.endswith(".py"):
                out.append(os.path.relpath(os.path.join(d, name), root))
    return sorted(out)


def fetch(verbose=True):
    """Copy the reference package; returns True when oracle/_ref is in place."""
    if not os.path.isdir(os.path.join(SRC_ROOT, "awkward0")):
        return os.path.exists(MANIFEST)
    manifest = {}
    for rel in _files(SRC_ROOT):
        dst = os.path.join(DST_ROOT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        src = os.path.join(SRC_ROOT, rel)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(MANIFEST, "w") as f:
        json.dump({"source": "scikit-hep/awkward-0.x (awkward0), copied unmodified from " + SRC_ROOT, "sha256": manifest},
                  f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d files of the unmodified reference" % len(manifest))
    return True


def verify():
    """True when every file of oracle/_ref matches its manifest (and the original, where that exists)."""
    if not os.path.exists(MANIFEST):
        return False
    with open(MANIFEST) as f:
        manifest = json.load(f)["sha256"]
    for rel, digest in manifest.items():
        p = os.path.join(DST_ROOT, rel)
        if not os.path.exists(p) or _sha(p) != digest:
            return False
        src = os.path.join(SRC_ROOT, rel)
        if os.path.exists(src) and _sha(src) != digest:
            return False
    return True


if __name__ == "__main__":
    ok = fetch()
    print("verified" if ok and verify() else "oracle/_ref NOT available")
    sys.exit(0 if ok else 1)
