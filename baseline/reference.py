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


def install() -> bool:
    """The recipe of the module docstring, for __graft_entry__.build(): installs the unmodified reference under
    baseline/_ref when it is missing and the source tree is mounted (build container only).  True when installed."""
    import shutil
    import subprocess
    import tempfile
    if os.path.isdir(os.path.join(INSTALLED, "lettuce")):
        return True
    if not os.path.isdir(os.path.join(SOURCE_TREE, "lettuce")):
        return False
    tmp = tempfile.mkdtemp(prefix="ref_copy_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(SOURCE_TREE, src, symlinks=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                        "/opt/wheelhouse", "--no-deps", "--target", INSTALLED, src], check=True,
                       stdout=subprocess.DEVNULL)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return os.path.isdir(os.path.join(INSTALLED, "lettuce"))


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



MANIFEST = os.path.join(INSTALLED, "native_manifest.json")


def native_key(generator) -> str:
    """What a generated kernel depends on (lettuce/cuda_native/_generator.py:36-44) without the version string: the
    reference's versioneer asks `git describe` of whatever repository its files lie in, so `Generator.name` differs
    between the build container (this repo's HEAD, "dirty" or not) and the GPU box (no .git)."""
    return " ".join([generator.stencil.__class__.__name__, generator.streaming_strategy.name]
                    + [t.__class__.__name__ for t in generator.transformer])


def native_modules() -> dict:
    """{native_key: package name} of the generated CUDA packages built ahead of time by baseline/build_native.py."""
    import json
    if not os.path.isfile(MANIFEST):
        return {}
    with open(MANIFEST) as fh:
        manifest = json.load(fh)
    return {k: v for k, v in manifest.items() if os.path.isdir(os.path.join(INSTALLED, v))}


def native_simulation(flow, collision, boundaries, strategy):
    """`lettuce.Simulation` of the unmodified reference on its generated CUDA kernel (`Context(use_native=True)`,
    lettuce/_simulation.py:172-229).  The prebuilt package is imported under the name it was built with and made
    known under the name `Generator.resolve()` will ask for in this process (`sys.modules` alias; see native_key);
    returns None when there is none, instead of letting the reference run `setup.py install` (no network, minutes
    of nvcc)."""
    lt = load()
    from lettuce.cuda_native import Generator
    equilibrium = flow.equilibrium.native_generator() if flow.equilibrium is not None else None
    generator = Generator(flow.stencil, collision=collision.native_generator(0), pre_boundaries=[],
                          post_boundaries=[], equilibrium=equilibrium, streaming_strategy=strategy)
    package = native_modules().get(native_key(generator))
    if boundaries or package is None:
        return None
    sys.path.insert(0, INSTALLED)
    try:
        sys.modules[f"lettuce_{generator.name}"] = importlib.import_module(package)
        return lt.Simulation(flow, collision, boundaries, strategy)
    finally:
        sys.path.remove(INSTALLED)
