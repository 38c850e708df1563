"""Benchmark of the stream+collide hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one lattice-Boltzmann time step of BASELINE.json's config 2: Taylor-Green vortex 3-D,
D3Q19, BGK, fp32, 256^3 nodes per GPU (weak scaling: the global lattice is [256*N, 256, 256], split
into x-slabs).  Prints ONE JSON line:

  value     MLUPS with the populations resident in HBM, CUDA-event timed, max over ranks
  e2e       MLUPS through the C ABI's host-buffer entry (lbm_run_host): pinned host populations are
            uploaded, K steps run with the kinetic energy read back to the host after every step,
            the final populations are downloaded -- all inside the timed region
  roofline  algorithmic bytes (2*q*4 B per node update) / measured launch duration vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json
  torch_gpu_port the reference's torch path restated op for op (oracle/torch_port.py), timed on the same GPU
  cpu_baseline   the NumPy oracle port of the reference's algorithm timed on this box's host cores
                 on a bounded sample (smaller lattice, same workload)

`--impl reference` times only the CPU arm (the oracle port of the reference's torch algorithm; the
reference itself is Python source that cannot travel to the GPU box) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

Q, BYTES_PER_NODE = 19, 2 * 19 * 4       # D3Q19 fp32: every population read once and written once
RE, MA = 1600.0, 0.05
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed
# `ncu --set full` captures (profiles/r1_step_d3q19_bgk_{pre,post}_256.csv); algorithmic = 2.550e9
NCU_DRAM_BYTES_PER_LAUNCH = {(256, "PRE_STREAMING"): 1.275082e9 + 1.224746e9,
                             (256, "POST_STREAMING"): 1.276132e9 + 1.225850e9,
                             # profiles/r1_step_d3q19_bgk_pre_512_dram.csv; algorithmic = 20.401e9
                             (512, "PRE_STREAMING"): 10.200856e9 + 10.150541e9}


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores
# ----------------------------------------------------------------------------------------------
def cpu_arm(steps: int, warmup: int, budget_s: float):
    """Time the NumPy oracle (oracle/lbm_oracle.py) on a bounded sample of the workload, on all host cores
    (x-chunks of the collide phase and the per-population rolls run in a thread pool)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import lbm_oracle as lo
    st = lo.stencil("D3Q19")
    cores = max(1, min(os.cpu_count() or 1, 32))

    with ThreadPoolExecutor(cores) as pool:
        def run(n, k):
            f, units = lo.tgv_initial(st, [n] * 3, RE, MA, dtype=np.float32)
            coll = dict(kind="bgk", tau=np.float32(units.tau))
            t0 = time.perf_counter()
            for _ in range(k):
                f = lo.step_parallel(st, f, coll, strategy="PRE_STREAMING", pool=pool, chunks=4 * cores)
            assert f.dtype == np.float32 and np.isfinite(f).all()
            return time.perf_counter() - t0

        probe = run(64, 2) / 2 / 64 ** 3                      # seconds per node update
        n = 64
        for cand in (96, 128, 160, 192, 256):
            if probe * cand ** 3 * (steps + warmup) <= budget_s:
                n = cand
        if warmup:
            run(n, warmup)
        dt = run(n, steps)
    mlups = steps * n ** 3 / 1e6 / dt
    return {"value": mlups, "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"TGV3D D3Q19 BGK fp32 {n}^3, {steps} steps, PRE_STREAMING, NumPy oracle with a {cores}-thread "
                      f"pool (host has {os.cpu_count()} cores)"}, dt / steps * 1e3


def torch_gpu_port(f0, tau, steps, strategy, dev):
    """MLUPS of oracle/torch_port.py (lettuce's sequence of full-size torch ops) on the GPU, explicit syncs."""
    import torch
    from oracle.torch_port import TorchBGK
    port = TorchBGK("D3Q19", tau, dev, torch.float32)
    f = f0.clone()
    f = port.step(f, strategy)                                  # warm-up (allocator, cuBLAS handles)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        f = port.step(f, strategy)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    nodes = f0[0].numel()
    assert torch.isfinite(f).all()
    return {"value": steps * nodes / 1e6 / dt, "unit": "MLUPS", "kind": "port", "steps": steps,
            "what": "oracle/torch_port.py: the reference's torch path restated op for op (sum, einsum, elementwise "
                    "equilibrium temporaries, one torch.roll per population), fp32 on this GPU, device-synchronised"}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_arm(args.steps, args.warmup, budget_s=150.0)
    line = {"impl": "reference", "metric": "MLUPS (TGV3D D3Q19 BGK fp32)", "value": base["value"], "unit": "MLUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, args.size, args.strategy), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, size, strategy="PRE_STREAMING"):
    return {"workload": f"TaylorGreenVortex3D D3Q19 BGK fp32, {size}^3 nodes per GPU (BASELINE.json configs[1]), "
                        f"Re {RE:g} Ma {MA:g}, tau from units, f_neq initialisation",
            "global_lattice": [size * n_gpus, size, size], "parallelism": f"x-slab x{n_gpus}",
            "streaming": strategy + (" (default of the reference's `lettuce benchmark`, lettuce/cli.py:82-85)"
                                     if strategy == "PRE_STREAMING" else ""),
            "l2": "working set 2.56 GB per GPU >> 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_main(args):
    import numpy as np
    import torch
    import lettuce_b200 as lt
    from lettuce_b200 import native

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n = args.size
    strategy = lt.StreamingStrategy[args.strategy]
    ctx = lt.Context(dev, dtype=torch.float32)
    if world == 1 and not args.slab:
        flow = lt.TaylorGreenVortex(ctx, [n] * 3, RE, MA, stencil=lt.D3Q19())
        sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], strategy)
        stepper = lambda k: native.invoke_n(sim, k)
    else:
        from lettuce_b200 import slab
        flow, sim, stepper = slab.make_tgv_slab_simulation(ctx, [n * world, n, n], RE, MA, lt.D3Q19(), strategy)
    nodes_local = n ** 3
    nodes_total = nodes_local * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stepper(max(args.warmup, 3))
    barrier()
    launches0 = native.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        start.record()
        stepper(args.steps)
        stop.record()
        barrier()
        launches = native.launch_count() - launches0
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        # The timed region is only tens of milliseconds: keep the same kernel running (untimed) for about half
        # a second so that the 50 ms clock samples are taken under this load.  The number of extra batches is
        # derived from the all-reduced time, i.e. identical on every rank (slabs advance in lock step).
        batch = 20 if args.steps >= 20 else 2 * ((args.steps + 1) // 2)
        for _ in range(int(600.0 / max(ms / args.steps * batch, 1e-3)) + 1):
            stepper(batch)
        torch.cuda.synchronize(dev)
        barrier()
    mlups = args.steps * nodes_total / 1e6 / (ms * 1e-3)
    assert torch.isfinite(flow.f).all()

    kernel_name = (native.engine_of(sim).variant_name if world == 1 and not args.slab
                   else "step_sync (slab, in-kernel lock step)")
    # ---- e2e: HOST populations in, HOST populations out, every step's kinetic energy read back
    e2e = None
    if not args.no_e2e:
        f_host = torch.empty(flow.f.shape, dtype=torch.float32).pin_memory()
        f_host.copy_(flow.f)
        out_host = torch.empty_like(f_host).pin_memory()
        fbytes = f_host.numel() * 4
        if world == 1:
            # N = 1: one call of the C ABI's host-buffer entry
            import ctypes as C
            energy = torch.zeros(args.steps, dtype=torch.float64).pin_memory()
            eng = native.engine_of(sim)
            torch.cuda.synchronize(dev)
            native.check(native.lib().lbm_run_host(C.byref(eng.desc), f_host.data_ptr(), out_host.data_ptr(), 1, None))
            dt = float("inf")
            for _ in range(2):      # host-side noise (page placement, PCIe contention) is large: best of two calls
                t0 = time.perf_counter()
                native.check(native.lib().lbm_run_host(C.byref(eng.desc), f_host.data_ptr(), out_host.data_ptr(),
                                                       args.steps, energy.data_ptr()))
                dt = min(dt, time.perf_counter() - t0)
            assert torch.isfinite(energy).all() and float(energy[-1]) > 0
            how = ("lbm_run_host (C ABI): pinned host f uploaded, K steps (fused step + energy kernel), kinetic energy read "
                   "back every step, final f downloaded; wall clock of the faster of two calls")
        else:
            # N > 1: the public Python API on every rank's slab -- upload, Simulation(K) with a global
            # kinetic-energy reporter of interval 1 (reduce kernel + all-reduce + D2H per step), download
            from lettuce_b200 import slab
            # defer=False: every step's energy is read back to the host inside the timed region (the contract's
            # per-step device-to-host read), not collected on the device and fetched at the end
            rep = lt.ObservableReporter(slab.GlobalSum(lt.IncompressibleKineticEnergy(flow)), interval=1, out=None,
                                        defer=False)
            sim.reporter.append(rep)
            flow.i = 1                      # skip the step-0 report so exactly K reports fall in the timed region
            barrier()
            t0 = time.perf_counter()
            native.engine_of(sim).load(f_host.to(dev, non_blocking=True))
            sim(args.steps)
            out_host.copy_(flow.f, non_blocking=True)
            barrier()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            assert len(rep.out) == args.steps and rep.out[-1][2] > 0
            sim.reporter.pop()
            how = ("public API per rank: pinned host slab uploaded, Simulation(K) with a global kinetic-energy reporter "
                   "(interval 1: reduce + all-reduce + D2H), final slab downloaded; wall clock, max over ranks")
        e2e = {"value": args.steps * nodes_total / 1e6 / dt, "unit": "MLUPS",
               "h2d_bytes_per_step": fbytes * world / args.steps, "d2h_bytes_per_step": fbytes * world / args.steps + 8,
               "note": how + "; transfers amortised over K steps"}

    torch_port = None
    if world == 1 and not args.slab and not args.no_e2e:
        # the reference's torch GPU path (restated, see oracle/torch_port.py) on the same lattice, for context
        try:
            torch_port = torch_gpu_port(flow.f, flow.units.relaxation_parameter_lu, 10, args.strategy, dev)
        except torch.OutOfMemoryError:
            torch_port = {"unavailable": "out of memory"}
    if world > 1:
        # orderly teardown on every rank: unmap the neighbours' buffers, then leave the process group together
        sim.close()
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    per_launch_ms = ms / args.steps
    achieved = nodes_local * BYTES_PER_NODE / (per_launch_ms * 1e-3) / 1e9
    line = {"metric": "MLUPS (TGV3D D3Q19 BGK fp32)", "value": mlups, "unit": "MLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_launch_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, n, args.strategy),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get((n, args.strategy)),
                         "traffic_unit": "bytes per launch (ncu, profiles/r1_step_d3q19_bgk_*.csv)",
                         "algorithmic_bytes_per_launch": nodes_local * BYTES_PER_NODE, "peak_source": peak_src,
                         "kernel": kernel_name,
                         "bytes_per_node": BYTES_PER_NODE},
            "clocks": clocks.summary(), "gpu_launches": int(launches), "e2e": e2e}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"], _ = cpu_arm(steps=3, warmup=1, budget_s=20.0)
    if torch_port is not None:
        line["torch_gpu_port"] = torch_port
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=256, help="nodes per axis per GPU")
    ap.add_argument("--strategy", default="PRE_STREAMING",
                    choices=["NO_STREAMING", "PRE_STREAMING", "POST_STREAMING", "DOUBLE_STREAMING"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (large --size)")
    ap.add_argument("--slab", action="store_true",
                    help="with --gpus 1: run the multi-GPU slab kernel (in-kernel lock step) with the rank as its own "
                         "neighbour, e.g. to profile it under ncu")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
