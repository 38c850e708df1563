"""Builds the reference's OWN generated CUDA kernel ("cuda_native", lettuce/cuda_native/_generator.py:86-127) for the
bench's `native_gpu_reference` leg -- in the build container, ahead of time, because the GPU box has no network and
the reference would otherwise run `setup.py install` into site-packages at run time (_generator.py:108-127).

Nothing of the reference is modified: its `Generator` writes the sources (`format`), its own generated `setup.py`
compiles them (same nvcc flags: --use_fast_math, -O3, --maxrregcount 128), only `install` is replaced by
`build_ext --inplace` and a copy of the resulting package `lettuce_<hash>/` next to the installed reference
(`baseline/_ref/`, git-ignored, travels with the gpurun snapshot).  The reference then finds the module through its
normal `Generator.resolve()` (`importlib.import_module("lettuce_<hash>")`, _generator.py:86-98) once `baseline/_ref`
is on `sys.path`.  `<hash>` covers the reference's version string, which its versioneer derives from `git describe`
of the enclosing repository -- different here and on the box -- so a manifest (`baseline/_ref/native_manifest.json`)
maps stencil / strategy / operators to the package, and baseline/reference.py:native_simulation aliases it.

    python baseline/build_native.py            # D3Q19 BGK, PRE_ and POST_STREAMING (the only natively generated
                                               # collision of the reference besides NoCollision)

The generated kernel addresses with 32-bit `int` (`using index_t = int`, _template.py:62), so q * nodes must stay
below 2^31: 384^3 is the largest bench cube for D3Q19.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def generators():
    import torch
    from baseline import reference
    lt = reference.load()
    from lettuce.cuda_native import Generator, StreamingStrategy
    ctx = lt.Context("cpu", dtype=torch.float32, use_native=False)
    flow = lt.TaylorGreenVortex(ctx, [8] * 3, 1600, 0.05, stencil=lt.D3Q19())
    collision = lt.BGKCollision(flow.units.relaxation_parameter_lu)
    for strategy in (StreamingStrategy.PRE_STREAMING, StreamingStrategy.POST_STREAMING):
        # the construction of lettuce/_simulation.py:191-214 for a flow without boundaries
        yield Generator(flow.stencil, collision=collision.native_generator(0), pre_boundaries=[],
                        post_boundaries=[], equilibrium=flow.equilibrium.native_generator(),
                        streaming_strategy=strategy)


def main():
    from baseline import reference
    assert os.path.isdir(os.path.join(reference.INSTALLED, "lettuce")), "install the reference first (reference.py)"
    env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a", MAX_JOBS=str(os.cpu_count() or 4))
    import json
    manifest = reference.native_modules()
    for gen in generators():
        key = reference.native_key(gen)
        if key in manifest:
            print(f"{manifest[key]}: present ({key})")
            continue
        package = f"lettuce_{gen.name}"
        target = os.path.join(reference.INSTALLED, package)
        directory = gen.format(tempfile.mkdtemp(prefix="lettuce_native_"))
        log = os.path.join(directory, "build.log")
        with open(log, "wb") as fh:
            subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=directory, env=env,
                           stdout=fh, stderr=fh, check=True)
        shutil.rmtree(target, ignore_errors=True)
        shutil.copytree(os.path.join(directory, package), target)
        manifest[key] = package
        with open(reference.MANIFEST, "w") as fh:
            json.dump(manifest, fh, indent=1)
        print(f"{package}: built for sm_100a ({key}), log {log}")


if __name__ == "__main__":
    main()
