"""TEST INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference package under oracle/_ref/ so that it travels to the
GPU box (oracle/_ref/ is git-ignored but not gpurun-ignored, exactly like the built .so files).

The reference is pure Python: there is nothing to compile, "building" it is a verbatim copy of the
``quant_transformer`` package directory from /root/reference, plus a manifest with the SHA-256 of every staged
file so a test can prove the staged tree IS the reference (tests/test_reference_dropin.py).  Nothing under
oracle/_ref/ is ever committed, imported by the product package, or edited.

    python -m oracle.make_ref          # (re)stage; no-op on a machine without /root/reference
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("OSQ_REFERENCE_ROOT", "/root/reference")
DST_ROOT = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(DST_ROOT, "MANIFEST.json")
PACKAGE = "quant_transformer"


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def _tree(root: str):
    out = {}
    base = os.path.join(root, PACKAGE)
    for dirpath, dirnames, files in os.walk(base):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        for f in files:
            if f.endswith(".py"):
                p = os.path.join(dirpath, f)
                out[os.path.relpath(p, root)] = _sha(p)
    return out


def source_available() -> bool:
    return os.path.isdir(os.path.join(SRC_ROOT, PACKAGE, "quantization"))


def staged() -> bool:
    return os.path.exists(MANIFEST) and os.path.isdir(os.path.join(DST_ROOT, PACKAGE, "quantization"))


def build(force: bool = False) -> str | None:
    """Copies /root/reference/quant_transformer -> oracle/_ref/quant_transformer.  Returns the staged root, or None
    when neither the source nor a previously staged copy exists."""
    if not source_available():
        return DST_ROOT if staged() else None
    want = _tree(SRC_ROOT)
    if not force and staged():
        try:
            have = json.load(open(MANIFEST))["files"]
        except Exception:
            have = None
        if have == want and _tree(DST_ROOT) == want:
            return DST_ROOT
    shutil.rmtree(os.path.join(DST_ROOT, PACKAGE), ignore_errors=True)
    os.makedirs(DST_ROOT, exist_ok=True)
    shutil.copytree(os.path.join(SRC_ROOT, PACKAGE), os.path.join(DST_ROOT, PACKAGE),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    json.dump({"source": SRC_ROOT, "package": PACKAGE, "files": want}, open(MANIFEST, "w"), indent=1, sort_keys=True)
    return DST_ROOT


def root() -> str | None:
    """Where the reference can be imported from on THIS machine: the read-only original if present, else the staged copy."""
    if source_available():
        return SRC_ROOT
    return DST_ROOT if staged() else None


def verify() -> bool:
    """True when the staged tree is byte-identical to what the manifest recorded at staging time."""
    if not staged():
        return False
    return _tree(DST_ROOT) == json.load(open(MANIFEST))["files"]


if __name__ == "__main__":
    print(build(force=True))
