#!/usr/bin/env python
"""Vendor the UNMODIFIED reference package into baseline/_ref/ (git-ignored, shipped to the GPU box by gpurun).

The contract's recipe -- `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target
baseline/_ref /root/reference` -- fails in this image: the reference builds with hatchling (pyproject.toml:1-3) and
hatchling is neither installed nor in /opt/wheelhouse ("ModuleNotFoundError: No module named 'hatchling'").  The wheel
that build would produce contains exactly one package, `UMNN` = the directory models/UMNN (pyproject.toml:30-31
`packages = ["models/UMNN"]`), pure Python with relative imports only.  This script copies that directory verbatim to
baseline/_ref/UMNN -- the same files at the same place `pip --target` would have put them -- and records a SHA-256
manifest so that tests can verify nothing was edited.

    python scripts/vendor_reference.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(REPO, "baseline", "_ref")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def vendor(src_root="/root/reference"):
    src = os.path.join(src_root, "models", "UMNN")
    if not os.path.isdir(src):
        raise SystemExit(f"{src} not found: the reference tree is only mounted in the build container")
    pkg = os.path.join(DST, "UMNN")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {"source": "AWehenkel/UMNN v2.0.5 (models/UMNN), copied verbatim", "files": {}}
    for name in sorted(os.listdir(pkg)):
        if name.endswith(".py"):
            manifest["files"][name] = sha256(os.path.join(pkg, name))
            assert manifest["files"][name] == sha256(os.path.join(src, name))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return pkg


if __name__ == "__main__":
    print(vendor(*(sys.argv[1:2])))
