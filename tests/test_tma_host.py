"""Tile geometry of the TMA-staged kernels (csrc/lbm_launch.cuh, lbm_tma.cuh) checked on the host: which lattices the
kernels take, how rows are cut into boxes, and that the stages fit the SM's shared memory.  (The kernels themselves are
pinned bit for bit against the LDG kernel by the GPU tests.)"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(ROOT, "tests", "csrc", "tma_host.cu")
    out = os.path.join(tempfile.gettempdir(), "lbm_b200_build", "libtma_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src] + [os.path.join(ROOT, "lettuce_b200", "csrc", h)
                    for h in ("lbm_launch.cuh", "lbm_tma.cuh", "lbm_step.cuh", "lbm_core.cuh", "lbm_vec.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        subprocess.run([nvcc, "-std=c++20", "-O1", "--expt-relaxed-constexpr", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
                        "-shared", "-o", out, src], check=True)
    L = C.CDLL(out)
    L.tma_host_available.argtypes = [C.c_int, C.c_longlong, C.c_int]
    return L


def test_row_boxes_tile_the_contiguous_axis(lib):
    T, max_rows = lib.tma_host_tile_nodes(), lib.tma_host_max_rows()
    assert T == 512
    for n2 in range(64, 4097, 64):
        tz, rows = lib.tma_host_row_extent(n2), lib.tma_host_tile_rows(n2)
        assert tz & (tz - 1) == 0 and 64 <= tz <= 256          # a power of two: box extents are at most 256 elements
        assert n2 % tz == 0                                     # no partial box along z: a shifted store could not clip
        assert tz == 256 or n2 % (2 * tz) != 0                  # the largest such box
        assert rows * tz == T and 2 <= rows <= max_rows         # one tile = 512 nodes = 256 consumer threads x 2


def test_which_lattices_the_staged_kernels_take(lib):
    F32, F64 = 0, 1
    assert lib.tma_host_available(F32, 512 ** 3, 512)
    assert not lib.tma_host_available(F64, 512 ** 3, 512)                  # fp32 only
    assert not lib.tma_host_available(F32, 512 * 512 * 100, 100)           # contiguous extent must be a multiple of 64
    assert not lib.tma_host_available(F32, 512 * 512 * 96, 96)
    assert lib.tma_host_available(F32, 4096 * 1024, 1024)                  # C4
    assert not lib.tma_host_available(F32, 64 * 64, 64)                    # too small to fill a persistent grid
    # full tiles never straddle two planes (3-D) / the row count divides nx (2-D): one box per tile and population
    assert lib.tma_host_rows_boxable(512, 512, 512) and lib.tma_host_rows_boxable(4096, 1, 1024)
    assert not lib.tma_host_rows_boxable(37, 9, 256) and not lib.tma_host_rows_boxable(300, 1, 320)


def test_stages_fit_the_shared_memory_of_an_sm(lib):
    limit = 227 * 1024
    for q, ctas, stages in ((27, 1, 4), (19, 1, 5), (9, 2, 5)):
        assert ctas * (stages * lib.tma_host_stage_bytes(q, 0) + 2048) <= limit, q      # pulling kernel
    for q, ctas, stages in ((27, 1, 3), (19, 1, 5), (9, 2, 4)):
        assert ctas * (stages * lib.tma_host_stage_bytes(q, 1) + 2048) <= limit, q      # pushing kernel
    for q in (9, 19, 27):
        for push in (0, 1):
            assert lib.tma_host_stage_bytes(q, push) % 128 == 0          # TMA boxes in shared memory: 128-byte aligned


def test_header_documents_the_staged_variant():
    with open(os.path.join(ROOT, "include", "lbm_b200.h")) as fh:
        text = fh.read()
    assert re.search(r"3 = TMA-staged kernel", text) and "multiple of 64" in text
