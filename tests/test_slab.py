"""Multi-rank x-slab tests.  CPU (gloo, world_size 2 and 3): decomposition, slab initial state and the
ring-exchange logic with the oracle as local stepper.  GPU (needs >= 2 devices): the peer-mapped CUDA
path must reproduce the single-GPU result bit for bit (run with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT
from lettuce_b200.slab import SlabDecomposition

sys.path.insert(0, os.path.join(ROOT, "tests"))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_decomposition_ranges():
    for nx, world in ((16, 1), (16, 2), (17, 4), (23, 8), (8, 8)):
        decs = [SlabDecomposition(nx, world, r) for r in range(world)]
        assert sum(d.nx_local for d in decs) == nx
        assert decs[0].x0 == 0 and decs[-1].x1 == nx
        for a, b in zip(decs, decs[1:]):
            assert a.x1 == b.x0
        assert max(d.nx_local for d in decs) - min(d.nx_local for d in decs) <= 1
        for d in decs:
            assert d.owner_of(d.x0) == d.rank and d.owner_of(d.x0 - 1) == d.lo and d.owner_of(d.x1) == d.hi
            assert d.halo_indices(1)[0] == (d.x0 - 1) % nx and d.halo_indices(1)[-1] == d.x1 % nx
    with pytest.raises(ValueError):
        SlabDecomposition(3, 4, 0)
    with pytest.raises(ValueError):
        SlabDecomposition(8, 2, 2)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_initial_state_and_ring_exchange_gloo(world):
    import slab_worker
    port = free_port()
    with mp.get_context("spawn").Manager() as mgr:
        results = mgr.dict()
        mp.spawn(slab_worker.cpu_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        for rank, out in results.items():
            for key, err in out.items():
                tol = 1e-14 if key.endswith("_init") else (1e-15 if key.endswith("_step") else 0.5)
                assert err < tol, (rank, key, err)


@pytest.mark.gpu
def test_slab_cuda_path_matches_single_gpu_bit_exact():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stderr[-4000:]
    assert "bit-exact=False" not in r.stdout and "bit-exact=True" in r.stdout
    assert r.stdout.count("[slab-obstacle]") >= 6
