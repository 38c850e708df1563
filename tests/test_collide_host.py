"""The collision operators of csrc/lbm_core.cuh, compiled for the HOST (tests/csrc/collide_host.cu; the operators are
written once for host and device), against the reference's golden outputs and the oracle -- no GPU needed.  Covers the
one-node float / double instantiations and the two-node float2 instantiation of the packed kernels, which must
agree with the float instantiation bit for bit."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, load_golden, max_rel
from oracle import lbm_oracle as lo

STENCIL_ID = {"D2Q9": 0, "D3Q19": 1, "D3Q27": 2}
KIND = {"none": 0, "bgk": 1, "trt": 2, "kbc": 3, "regularized": 4, "smagorinsky": 5, "bgk_forced": 6}


@pytest.fixture(scope="module")
def host_lib():
    src = os.path.join(ROOT, "tests", "csrc", "collide_host.cu")
    out = os.path.join(tempfile.gettempdir(), "lbm_b200_build", "libcollide_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "lettuce_b200", "csrc", h) for h in ("lbm_core.cuh", "lbm_vec.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-std=c++20", "-O2", "--expt-relaxed-constexpr", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                        "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.collide_host.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_double, C.c_double,
                                 C.c_void_p, C.c_double, C.c_double]
    lib.collide_host.restype = C.c_int
    return lib


def collide(lib, stencil, coll, f, dtype, packed=False, p0=1.0, p1=1.0, force=None, ueq=0.0, src=0.0):
    """f: [q, *res] float64 -> collided populations in `dtype` arithmetic (the operators are node-local, so the
    node order does not matter)"""
    q = f.shape[0]
    assert lo.stencil(stencil)["q"] == q
    g = np.ascontiguousarray(f.reshape(q, -1).astype(dtype))
    fptr = None
    if force is not None:
        farr = (C.c_double * 3)(*(list(force) + [0.0] * (3 - len(force))))
        fptr = C.cast(farr, C.c_void_p)
    rc = lib.collide_host(STENCIL_ID[stencil], 0 if dtype == np.float32 else 1, KIND[coll], int(packed),
                          g.ctypes.data, g.shape[1], p0, p1, fptr, ueq, src)
    assert rc == 0, rc
    return g.reshape(f.shape)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_host_operators_match_reference_golden(host_lib, dtype):
    g = load_golden("random_collisions")
    tol = 1e-12 if dtype == np.float64 else 1e-5
    for stencil, res in (("D2Q9", [6, 5]), ("D3Q19", [4, 5, 6]), ("D3Q27", [4, 5, 6])):
        units = lo.Units(50.0, 0.1, characteristic_length_lu=res[0])
        f0 = g[f"{stencil}_f0"]
        for coll in ("bgk", "trt", "kbc", "regularized", "smagorinsky"):
            if coll == "kbc" and stencil == "D3Q19":
                continue
            p0, p1 = {"bgk": (units.tau, 0), "trt": (units.tau, 0.8), "kbc": (units.tau, 0),
                      "regularized": (units.tau, 0), "smagorinsky": (units.tau, 0.17)}[coll]
            out = collide(host_lib, stencil, coll, f0, dtype, p0=p0, p1=p1)
            assert max_rel(out, g[f"{stencil}_{coll}"]) < tol, (stencil, coll, dtype)


def test_packed_instantiation_equals_scalar_bit_for_bit(host_lib):
    rng = np.random.default_rng(3)
    for stencil in ("D2Q9", "D3Q19", "D3Q27"):
        st = lo.stencil(stencil)
        f0 = st["w"][:, None] * (1.0 + 0.2 * (rng.random((st["q"], 64)) - 0.5))
        for coll, p0, p1 in (("bgk", 0.61, 0), ("trt", 0.55, 0.9), ("kbc", 0.52, 0), ("regularized", 0.7, 0),
                             ("smagorinsky", 0.51, 0.17), ("none", 1, 1)):
            if coll == "kbc" and stencil == "D3Q19":
                continue
            a = collide(host_lib, stencil, coll, f0, np.float32, packed=False, p0=p0, p1=p1)
            b = collide(host_lib, stencil, coll, f0, np.float32, packed=True, p0=p0, p1=p1)
            assert np.array_equal(a, b), (stencil, coll)
        force = [1e-4, -2e-4, 5e-5][:st["d"]]
        a = collide(host_lib, stencil, "bgk_forced", f0, np.float32, False, p0=0.8, force=force, ueq=0.5, src=0.6)
        b = collide(host_lib, stencil, "bgk_forced", f0, np.float32, True, p0=0.8, force=force, ueq=0.5, src=0.6)
        assert np.array_equal(a, b), stencil


@pytest.mark.parametrize("scheme", ["guo", "shan_chen"])
def test_forced_bgk_matches_oracle(host_lib, scheme):
    rng = np.random.default_rng(5)
    for stencil in ("D2Q9", "D3Q19", "D3Q27"):
        st = lo.stencil(stencil)
        d = st["d"]
        shape = (8, 5) if d == 2 else (2, 4, 5)          # the oracle's force broadcast wants d spatial axes
        f0 = st["w"].reshape((-1,) + (1,) * d) * (1.0 + 0.2 * (rng.random((st["q"], *shape)) - 0.5))
        tau = 0.83
        acc = np.array([2e-4, -1e-4, 3e-4][:d])
        ref = lo.collide_bgk_forced(st, f0, tau, acc, scheme=scheme)
        ueq = 0.5 if scheme == "guo" else tau
        src = 1.0 - 1.0 / (2.0 * tau) if scheme == "guo" else 0.0
        out = collide(host_lib, stencil, "bgk_forced", f0, np.float64, p0=tau, force=list(acc), ueq=ueq, src=src)
        assert max_rel(out, ref) < 1e-12, (stencil, scheme)
        out32 = collide(host_lib, stencil, "bgk_forced", f0, np.float32, p0=tau, force=list(acc), ueq=ueq, src=src)
        assert max_rel(out32, ref) < 1e-5, (stencil, scheme)


@pytest.mark.parametrize("coll", ["bgk", "trt", "kbc", "regularized", "smagorinsky"])
def test_collisions_conserve_mass_and_momentum(host_lib, coll):
    """tests/collision/test_collision_conserves_mass.py / _momentum.py of the reference"""
    rng = np.random.default_rng(8)
    for stencil in ("D2Q9", "D3Q27"):
        st = lo.stencil(stencil)
        f0 = st["w"][:, None] * (1.0 + 0.3 * (rng.random((st["q"], 50)) - 0.5))
        out = collide(host_lib, stencil, coll, f0, np.float64, p0=0.6, p1=0.17 if coll == "smagorinsky" else 1.1)
        assert np.allclose(lo.rho(out), lo.rho(f0), rtol=1e-13, atol=0)
        assert np.allclose(lo.j(st, out), lo.j(st, f0), rtol=0, atol=1e-14)
