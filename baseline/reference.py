"""Loader of the UNMODIFIED reference (lettucecfd/lettuce) for the reference arm of bench.py and for the drop-in
tests that drive the reference's own objects.

The reference is installed once, in the build container, with

    cp -r /root/reference /tmp/ref_copy
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
        --target baseline/_ref /tmp/ref_copy

(`--no-deps`: mmh3, pyevtk, h5py, vtk and matplotlib are not in the offline wheelhouse; from a copy because the
build writes an egg-info directory into the source tree and /root/reference is read-only).  `baseline/_ref` is
git-ignored and travels to the GPU box with the gpurun snapshot.  Three third-party imports that are NOT on the
arithmetic path are satisfied with stub modules: `mmh3` (hashes generated-module names,
lettuce/cuda_native/_util.py:3), `pyevtk.hl` (VTK output, lettuce/ext/_reporter/vtk_reporter.py:2), `h5py`
(lettuce/util/datautils.py:5).  Nothing of the reference is modified or re-implemented here.

Product code never imports this module: only bench.py's reference / cpu_baseline legs and tests do.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED = os.path.join(HERE, "_ref")
SOURCE_TREE = "/root/reference"          # build container only

_module = None


def available() -> bool:
    return os.path.isdir(os.path.join(INSTALLED, "lettuce")) or os.path.isdir(os.path.join(SOURCE_TREE, "lettuce"))


def location() -> str:
    return INSTALLED if os.path.isdir(os.path.join(INSTALLED, "lettuce")) else SOURCE_TREE


def _stubs():
    mmh3 = types.ModuleType("mmh3")
    mmh3.hash_bytes = lambda v: hashlib.md5(v.encode() if isinstance(v, str) else v).digest()
    hl = types.ModuleType("pyevtk.hl")
    hl.gridToVTK = lambda *a, **k: None
    pyevtk = types.ModuleType("pyevtk")
    pyevtk.hl = hl
    h5py = types.ModuleType("h5py")
    h5py.File = None
    return {"mmh3": mmh3, "pyevtk": pyevtk, "pyevtk.hl": hl, "h5py": h5py}


def load():
    """`import lettuce` from baseline/_ref (or /root/reference in the build container); returns the module."""
    global _module
    if _module is not None:
        return _module
    if not available():
        raise RuntimeError("the reference is not installed under baseline/_ref (see baseline/reference.py)")
    for name, stub in _stubs().items():
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = stub
    sys.path.insert(0, location())
    try:
        _module = importlib.import_module("lettuce")
    finally:
        sys.path.remove(location())
    return _module
