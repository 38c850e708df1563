"""Ahead-of-time build of liblbm_b200.so (sm_100a) with plain nvcc.

No per-configuration JIT (the reference generates and compiles one extension per
(stencil, strategy, operator list), lettuce/cuda_native/_generator.py:99-127):
every (stencil, dtype, collision, streaming, masked) variant is template-instantiated
once and selected at run time from the descriptor.  The 40 (stencil, dtype, collision)
triples are separate translation units and compile in parallel.

    python -m lettuce_b200.build [--force] [--verbose]

Nothing but the default `liblbm_b200.so` (+ its `.sha256` stamp) is ever written inside the repository: objects
go to `$LBM_B200_BUILD_DIR` (default `$TMPDIR/lbm_b200_build/<suffix>`), so that the tree that `gpurun` snapshots
stays small.  The default library is built without `-lineinfo` (a third of the size); `LBM_B200_LINEINFO=1` adds it
for ncu source-page sessions.

A/B experiments: `LBM_B200_NVCC_DEFINES="-DFOO=1" LBM_B200_BUILD_SUFFIX=foo python -m lettuce_b200.build` writes
`liblbm_b200_foo.so` into the build directory (or into `$LBM_B200_VARIANT_DIR`, e.g. `gpurun_out/../variants`, when it
has to travel); select it at run time with `LBM_B200_LIB=<path>` (lettuce_b200/native.py).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
_SUFFIX = os.environ.get("LBM_B200_BUILD_SUFFIX", "")
OBJ = os.environ.get("LBM_B200_BUILD_DIR") or os.path.join(tempfile.gettempdir(), "lbm_b200_build",
                                                         _SUFFIX or "default")
if _SUFFIX:
    LIB = os.path.join(os.environ.get("LBM_B200_VARIANT_DIR") or OBJ, "liblbm_b200_" + _SUFFIX + ".so")
else:
    LIB = os.path.join(PKG, "liblbm_b200.so")
STAMP = LIB + ".sha256"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++20", "-O3", *(["-lineinfo"] if os.environ.get("LBM_B200_LINEINFO") == "1" else []),
         "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"),
         *os.environ.get("LBM_B200_NVCC_DEFINES", "").split()]

STENCILS = ("D2Q9", "D3Q19", "D3Q27")
REALS = ("float", "double")


def _units():
    units = [("lbm_api", os.path.join(CSRC, "lbm_api.cu"), []),
             ("lbm_moments", os.path.join(CSRC, "lbm_moments.cu"), []),
             ("lbm_slab", os.path.join(CSRC, "lbm_slab.cu"), [])]
    # heaviest units first so the pool's tail is short
    for c, cname in ((3, "kbc"), (5, "smagorinsky"), (4, "regularized"), (6, "bgk_forced"), (2, "trt"), (1, "bgk"),
                     (0, "none")):
        for s in reversed(STENCILS):
            if cname == "kbc" and s == "D3Q19":
                continue            # KBC exists for D2Q9 and D3Q27 only
            for r in reversed(REALS):
                units.append((f"lbm_step_{s}_{r}_{cname}", os.path.join(CSRC, "lbm_step_inst.cu"),
                              [f"-DLBM_INST_STENCIL={s}", f"-DLBM_INST_REAL={r}", f"-DLBM_INST_COLL={c}"]))
    return units


def _headers_hash() -> "hashlib._Hash":
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(d)):
            if name.endswith((".cuh", ".h")):
                with open(os.path.join(d, name), "rb") as fh:
                    h.update(name.encode() + b"\0" + fh.read())
    # (without the absolute include path: the digest must not depend on where the checkout lives)
    h.update(" ".join(f for f in FLAGS if not f.startswith(ROOT)).encode())
    return h


def _unit_hash(unit) -> str:
    name, src, defs = unit
    h = _headers_hash()
    with open(src, "rb") as fh:
        h.update(fh.read())
    h.update(" ".join(defs).encode())
    return h.hexdigest()


def source_digest(digests=None) -> str:
    """sha256 over every translation unit's (source, headers, flags) hash: identifies the sources a library was
    built from (written next to the library as `<lib>.sha256`)."""
    if digests is None:
        digests = {u[0]: _unit_hash(u) for u in _units()}
    return hashlib.sha256("".join(sorted(digests.values())).encode()).hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library."""
    units = _units()
    digests = {u[0]: _unit_hash(u) for u in units}
    stamp = STAMP
    digest = source_digest(digests)
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read().strip() == digest):
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)

    def compile_one(unit):
        name, src, defs = unit
        obj = os.path.join(OBJ, name + ".o")
        ustamp = obj + ".sha256"
        if (not force and os.path.exists(obj) and os.path.exists(ustamp)
                and open(ustamp).read().strip() == digests[name]):
            return obj
        cmd = [NVCC, *FLAGS, *defs, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(ustamp, "w") as fh:
            fh.write(digests[name])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, units))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
