"""Benchmark of the stream+collide hot path (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2|c3] [--size n]

A "step" is one lattice-Boltzmann time step of a Taylor-Green vortex, fp32, n^3 nodes PER GPU (weak scaling: the
global lattice is [n*N, n, n], split into x-slabs):

    --config c2 (default)  D3Q19 BGK   BASELINE.json configs[1]; default n = 512, the size north_star's target
                                       sentence is quoted on ("fused D3Q19 BGK TGV 512^3 fp32 >= 80 % of the HBM
                                       roofline"); the same run at configs[1]'s own 256^3 is reported under "at_256"
    --config c3            D3Q27 KBC   BASELINE.json configs[2] (512^3)

Prints ONE JSON line:

  value      MLUPS with the populations resident in HBM, CUDA-event timed, max over ranks
  e2e        MLUPS through the C ABI's host-buffer entry (lbm_run_host, N = 1) or the public Python API on every
             rank's slab (N > 1): pinned host populations uploaded, K steps with the kinetic energy of EVERY step
             read back to the host, final populations downloaded -- all inside the timed region
  roofline   algorithmic bytes (2*q*4 B per node update) / measured launch duration vs the measured HBM copy
             bandwidth in MEASURED_PEAKS.json; traffic = ncu DRAM bytes per launch read from the committed capture
             under profiles/ (null when there is none for this configuration)
  cpu_baseline        the UNMODIFIED reference (baseline/_ref, lettuce.Simulation on Context('cpu',
                      use_native=False)) timed on this box's host cores on a bounded sample of the workload
  torch_gpu_reference the same unmodified reference on Context('cuda', use_native=False), device-synchronised
  native_gpu_reference the same unmodified reference on its own generated CUDA kernel (use_native=True), prebuilt
                      for sm_100a by baseline/build_native.py
  c3_512              (default configuration only) value / roofline / clocks of `bench.py --config c3` (D3Q27 KBC
                      512^3, BASELINE.json configs[2]) run in a process of its own

`--impl reference` times only the reference's own CPU implementation (stock lettuce.Simulation, all host threads)
and prints the same line shape; under torchrun only rank 0 works.
"""
from __future__ import annotations

import argparse
import csv
import glob
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

RE, MA = 1600.0, 0.05
CONFIGS = {
    "c2": dict(stencil="D3Q19", q=19, collision="BGK", default_size=512,
               name="TaylorGreenVortex3D D3Q19 BGK fp32", baseline="BASELINE.json configs[1] (256^3; north_star's "
               "roofline target is quoted at 512^3)"),
    "c3": dict(stencil="D3Q27", q=27, collision="KBC", default_size=512,
               name="TaylorGreenVortex3D D3Q27 KBC fp32", baseline="BASELINE.json configs[2] (512^3)"),
}


def measured_hbm_peak():
    """HBM copy bandwidth in GB/s from the driver-written MEASURED_PEAKS.json (the sustained figure where the file
    distinguishes one: the step kernel is timed inside a long run of back-to-back launches), else the fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            data = json.load(fh)
        found = []

        def walk(node, trail):
            if isinstance(node, dict):
                for k, v in node.items():
                    walk(v, trail + [str(k).lower()])
            elif isinstance(node, (int, float)) and not isinstance(node, bool):
                name = ".".join(trail)
                if "hbm" in name and 1000.0 <= float(node) <= 9000.0:
                    found.append((name, float(node)))
        walk(data, [])
        for want in ("sustain", "hbm_gbs", "burst", ""):
            for name, value in found:
                if want in name:
                    return value, f"measured (MEASURED_PEAKS.json: {name})"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(config: str, size: int, strategy: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed `ncu --set
    full` capture of this configuration (profiles/r*_dram_<config>_<size>_<pre|post>.csv, written by
    profiles/summarize.py); (None, None) when no capture of this configuration is committed."""
    tag = {"PRE_STREAMING": "pre", "POST_STREAMING": "post"}.get(strategy)
    if tag is None:
        return None, None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_dram_{config}_{size}_{tag}.csv")), reverse=True):
        try:
            rows = {}
            with open(path) as fh:
                for r in csv.DictReader(fh):
                    try:
                        rows[r["metric"]] = float(r["value"])
                    except ValueError:
                        pass                                   # the kernel-name rows
            return rows["dram__bytes_read.sum"] + rows["dram__bytes_write.sum"], os.path.relpath(path, ROOT)
        except Exception:
            continue
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
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
        # samples under load only (the sampler also sees the idle moments around the timed region)
        busy = [c for c in sm if c < max(smax) - 1] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(config: str, n_gpus: int, size: int, strategy: str):
    c = CONFIGS[config]
    f_gb = c["q"] * size ** 3 * 4 / 1e9
    return {"workload": f"{c['name']}, {size}^3 nodes per GPU ({c['baseline']}), Re {RE:g} Ma {MA:g}, "
                        f"tau from units, f_neq initialisation",
            "global_lattice": [size * n_gpus, size, size], "parallelism": f"x-slab x{n_gpus}",
            "streaming": strategy + (" (default of the reference's `lettuce benchmark`, lettuce/cli.py:82-85)"
                                     if strategy == "PRE_STREAMING" else ""),
            "l2": f"working set {2 * f_gb:.2f} GB per GPU (two population buffers) >> 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------
# reference arm: the unmodified reference (baseline/_ref) through its own public API
# ----------------------------------------------------------------------------------------------
def _reference_simulation(lt, device, config, n, strategy):
    import torch
    c = CONFIGS[config]
    ctx = lt.Context(device=device, dtype=torch.float32, use_native=False)      # stock torch path (cli.py:98-118)
    flow = lt.TaylorGreenVortex(ctx, [n] * 3, RE, MA, stencil=getattr(lt, c["stencil"])())
    collision = (lt.BGKCollision(flow.units.relaxation_parameter_lu) if c["collision"] == "BGK"
                 else lt.KBCCollision())
    return flow, lt.Simulation(flow, collision, [], lt.StreamingStrategy[strategy])


def reference_cpu(config: str, strategy: str, steps: int, warmup: int, budget_s: float, want_size: int):
    """MLUPS of the reference's torch CPU path (lettuce.Simulation.__call__, Context('cpu', use_native=False)) on
    all host threads.  The lattice is `want_size`^3 when steps+warmup of it fit `budget_s` (and host RAM), else the
    largest smaller cube that does.  Falls back to the NumPy oracle port when baseline/_ref is absent."""
    import torch
    from baseline import reference
    cores = os.cpu_count() or 1
    if not reference.available():
        return _oracle_port_cpu(config, strategy, steps, warmup, budget_s)
    lt = reference.load()
    # all host cores: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the reference's torch CPU
    # path on ONE thread (measured: 1.7 instead of 12 MLUPS) -- the environment variable only sets the default
    if torch.get_num_threads() < cores:
        torch.set_num_threads(cores)
    threads = torch.get_num_threads()

    def run(n, k, w):
        flow, sim = _reference_simulation(lt, "cpu", config, n, strategy)
        if w:
            sim(w)
        t0 = time.perf_counter()
        sim(k)
        dt = time.perf_counter() - t0
        assert torch.isfinite(flow.f).all()
        return dt

    per_node = run(64, 2, 1) / 2 / 64 ** 3                      # probe: seconds per node update
    try:
        import psutil
        ram = psutil.virtual_memory().available
    except Exception:
        ram = 32e9
    q = CONFIGS[config]["q"]
    n = 64
    for cand in (96, 128, 160, 192, 256, 320, 384):      # 512^3 would need ~140 GB of torch temporaries
        fits_time = per_node * cand ** 3 * (steps + warmup) <= budget_s
        fits_ram = 14 * q * cand ** 3 * 4 <= 0.6 * ram          # the torch path holds ~10 full-size temporaries
        if cand <= want_size and fits_time and fits_ram:
            n = cand
    dt = run(n, steps, warmup)
    mlups = steps * n ** 3 / 1e6 / dt
    return ({"value": mlups, "unit": "MLUPS", "cores": threads, "kind": "reference",
             "sample": f"{CONFIGS[config]['name']} {n}^3, {steps} steps after {warmup} warm-up, {strategy}: unmodified "
                       f"lettuce.Simulation from baseline/_ref on Context('cpu', use_native=False), "
                       f"{threads} torch threads (host has {cores} cores)",
             "lattice": [n] * 3}, dt / steps * 1e3, n)


def _oracle_port_cpu(config, strategy, steps, warmup, budget_s):
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import lbm_oracle as lo
    c = CONFIGS[config]
    st = lo.stencil(c["stencil"])
    cores = max(1, min(os.cpu_count() or 1, 32))
    with ThreadPoolExecutor(cores) as pool:
        def run(n, k):
            f, units = lo.tgv_initial(st, [n] * 3, RE, MA, dtype=np.float32)
            coll = dict(kind=c["collision"].lower(), tau=np.float32(units.tau))
            t0 = time.perf_counter()
            for _ in range(k):
                f = lo.step_parallel(st, f, coll, strategy=strategy, pool=pool, chunks=4 * cores)
            assert np.isfinite(f).all()
            return time.perf_counter() - t0
        probe = run(64, 2) / 2 / 64 ** 3
        n = 64
        for cand in (96, 128, 160, 192, 256):
            if probe * cand ** 3 * (steps + warmup) <= budget_s:
                n = cand
        if warmup:
            run(n, warmup)
        dt = run(n, steps)
    return ({"value": steps * n ** 3 / 1e6 / dt, "unit": "MLUPS", "cores": cores, "kind": "port",
             "sample": f"{c['name']} {n}^3, {steps} steps, {strategy}: NumPy oracle port (baseline/_ref is missing) with a "
                       f"{cores}-thread pool", "lattice": [n] * 3}, dt / steps * 1e3, n)


def reference_torch_gpu(config, strategy, n, steps, dev):
    """The unmodified reference on the GPU through its stock torch path, device-synchronised around the timed
    region (the reference's own MLUPS is not, lettuce/_simulation.py:311-323).  Falls back to smaller cubes when
    its full-size temporaries do not fit."""
    import torch
    from baseline import reference
    if not reference.available():
        return {"unavailable": "baseline/_ref is missing"}
    lt = reference.load()
    for size in [s for s in (n, 384, 256, 128) if s <= n]:
        try:
            flow, sim = _reference_simulation(lt, dev, config, size, strategy)
            sim(2)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            sim(steps)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            assert torch.isfinite(flow.f).all()
            return {"value": steps * size ** 3 / 1e6 / dt, "unit": "MLUPS", "kind": "reference", "steps": steps,
                    "lattice": [size] * 3, "same_config": size == n,
                    "what": "unmodified lettuce.Simulation (baseline/_ref) on Context('cuda', use_native=False), "
                            "fp32, this GPU, torch.cuda.synchronize() on both sides"}
        except torch.OutOfMemoryError:
            flow = sim = None
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    return {"unavailable": "out of memory at every size"}


def reference_native_gpu(config, strategy, n, steps, dev):
    """The unmodified reference on ITS OWN generated CUDA kernel (Context('cuda', use_native=True),
    lettuce/_simulation.py:172-229, lettuce/cuda_native/_template.py:64-81), prebuilt for sm_100a by
    baseline/build_native.py.  The generated kernel exists for BGK only and indexes with 32-bit ints, so the cube
    is the largest of (n, 384, 256) with q * nodes < 2^31; it synchronises the device after every launch itself."""
    import torch
    from baseline import reference
    c = CONFIGS[config]
    if not reference.available():
        return {"unavailable": "baseline/_ref is missing"}
    if c["collision"] != "BGK":
        return {"unavailable": f"the reference generates no native kernel for {c['collision']} "
                               "(lettuce/cuda_native/ext/_collision: BGK and NoCollision only)"}
    lt = reference.load()
    size = max(s for s in (n, 384, 256, 128) if s <= n and c["q"] * s ** 3 < 2 ** 31)
    ctx = lt.Context(device=dev, dtype=torch.float32, use_native=True)
    flow = lt.TaylorGreenVortex(ctx, [size] * 3, RE, MA, stencil=getattr(lt, c["stencil"])())
    sim = reference.native_simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [],
                                      lt.StreamingStrategy[strategy])
    if sim is None:
        return {"unavailable": "generated module not prebuilt (python baseline/build_native.py)"}
    sim(3)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    sim(steps)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    assert torch.isfinite(flow.f).all()
    v = steps * size ** 3 / 1e6 / dt
    return {"value": v, "unit": "MLUPS", "kind": "reference", "steps": steps, "lattice": [size] * 3,
            "same_config": size == n, "roofline_frac": v * 1e6 * 2 * c["q"] * 4 / 1e9 / measured_hbm_peak()[0],
            "what": "unmodified lettuce.Simulation (baseline/_ref) on Context('cuda', use_native=True): the "
                    "reference's generated CUDA kernel (8x8x8 threads, one node per thread, --use_fast_math), "
                    "compiled for sm_100a from the reference's own generator and setup.py, fp32, this GPU; "
                    "32-bit indices limit it to q * nodes < 2^31"}


def leg_in_subprocess(cmd_args, timeout=240):
    """One comparison leg in a process of its own (its memory is returned before the next leg; a fault inside the
    reference's generated kernel -- a sticky CUDA error -- or a module that does not load cannot cost the contract
    line).  Returns the last JSON object the child printed."""
    cmd = [sys.executable, os.path.abspath(__file__)] + [str(a) for a in cmd_args]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")}
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
        lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
        if out.returncode == 0 and lines:
            return json.loads(lines[-1])
        return {"unavailable": f"leg exited {out.returncode}: {out.stderr.strip()[-300:]}"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def native_leg_in_subprocess(args, n):
    return leg_in_subprocess(["--leg", "native_gpu_reference", "--config", args.config, "--strategy", args.strategy,
                              "--size", n])


def second_config_leg(args):
    """The D3Q27 KBC configuration (BASELINE.json configs[2], `bench.py --config c3`) as a secondary key of the
    default line: device-resident value and roofline of its own run, in its own process."""
    line = leg_in_subprocess(["--config", "c3", "--steps", args.steps, "--warmup", args.warmup, "--quick", "--no-cpu",
                              "--no-e2e"], timeout=300)
    if "unavailable" in line:
        return line
    return {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"],
            "lattice": line["config"]["global_lattice"], "roofline": line["roofline"], "clocks": line["clocks"],
            "gpu_launches": line["gpu_launches"], "note": "python bench.py --config c3 (device-resident leg only)"}


def reference_main(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    base, ms, n = reference_cpu(args.config, args.strategy, args.steps, args.warmup, budget_s=150.0,
                                want_size=args.size)
    cfg = workload_config(args.config, args.gpus, args.size, args.strategy)
    cfg["reference_lattice"] = [n] * 3
    cfg["same_config"] = n == args.size and args.gpus == 1
    line = {"impl": "reference", "metric": f"MLUPS ({CONFIGS[args.config]['name']})", "value": base["value"],
            "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def gpu_main(args):
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
        # stdout carries the JSON line only: whatever NCCL prints while the communicator comes up (its version line,
        # debug output when NCCL_DEBUG is set on the box) is sent to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    c = CONFIGS[args.config]
    bytes_per_node = 2 * c["q"] * 4          # every population read once and written once
    strategy = lt.StreamingStrategy[args.strategy]
    stencil_cls = getattr(lt, c["stencil"])

    def collision_of(flow):
        return lt.BGKCollision(flow.units.relaxation_parameter_lu) if c["collision"] == "BGK" else lt.KBCCollision()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def build(n):
        ctx = lt.Context(dev, dtype=torch.float32)
        if world == 1 and not args.slab:
            flow = lt.TaylorGreenVortex(ctx, [n] * 3, RE, MA, stencil=stencil_cls())
            sim = lt.Simulation(flow, collision_of(flow), [], strategy)
        else:
            from lettuce_b200 import slab
            flow, sim, _ = slab.make_tgv_slab_simulation(ctx, [n * world, n, n], RE, MA, stencil_cls(), strategy,
                                                         collision_factory=collision_of)
        return flow, sim

    def device_timed(sim, steps, warmup):
        """(ms for `steps` steps [max over ranks], launches): CUDA events on the launching stream"""
        native.invoke_n(sim, max(warmup, 3))
        barrier()
        launches0 = native.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        native.invoke_n(sim, steps)
        stop.record()
        barrier()
        launches = native.launch_count() - launches0
        ms = start.elapsed_time(stop)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    n = args.size
    flow, sim = build(n)
    nodes_local, nodes_total = n ** 3, n ** 3 * world
    with ClockSampler(local) as clocks:
        ms, launches = device_timed(sim, args.steps, args.warmup)
        # The timed region may be short: keep the same kernel running (untimed) for about half a second so that the
        # 50 ms clock samples are taken under this load.  The number of extra batches is derived from the
        # all-reduced time, i.e. identical on every rank (slabs advance in lock step).
        batch = 20 if args.steps >= 20 else 2 * ((args.steps + 1) // 2)
        for _ in range(int(600.0 / max(ms / args.steps * batch, 1e-3)) + 1):
            native.invoke_n(sim, batch)
        barrier()
    mlups = args.steps * nodes_total / 1e6 / (ms * 1e-3)
    assert torch.isfinite(flow.f).all()
    kernel_name = (native.engine_of(sim).variant_name if world == 1 and not args.slab
                   else "step_sync (slab, in-kernel lock step)")

    # ---- e2e: HOST populations in, HOST populations out, every step's kinetic energy read back
    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(args, lt, native, flow, sim, dev, world, barrier, nodes_total)

    if world > 1:
        # orderly teardown on every rank: unmap the neighbours' buffers, then leave the process group together
        sim.close()
        dist.barrier()
        dist.destroy_process_group()
    del flow, sim
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    extra = {}
    if world == 1 and not args.slab and not args.quick:
        if args.config == "c2" and n != 256:
            # BASELINE.json configs[1]'s own lattice, same kernel
            flow2, sim2 = build(256)
            ms2, _ = device_timed(sim2, args.steps, args.warmup)
            v2 = args.steps * 256 ** 3 / 1e6 / (ms2 * 1e-3)
            peak, _ = measured_hbm_peak()
            extra["at_256"] = {"value": v2, "unit": "MLUPS", "ms_per_step": ms2 / args.steps,
                               "roofline_frac": v2 * 1e6 * bytes_per_node / 1e9 / peak,
                               "lattice": [256] * 3, "note": "BASELINE.json configs[1] (256^3), device-resident"}
            del flow2, sim2
            gc.collect()
            torch.cuda.empty_cache()
        extra["torch_gpu_reference"] = reference_torch_gpu(args.config, args.strategy, n, 5, dev)
        gc.collect()
        torch.cuda.empty_cache()                      # the legs below run in processes of their own
        extra["native_gpu_reference"] = native_leg_in_subprocess(args, n)
        if args.config == "c2":
            extra["c3_512"] = second_config_leg(args)
    if rank != 0:
        return
    peak, peak_src = measured_hbm_peak()
    per_launch_ms = ms / args.steps
    achieved = nodes_local * bytes_per_node / (per_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.config, n, args.strategy)
    line = {"metric": f"MLUPS ({c['name']})", "value": mlups, "unit": "MLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per_launch_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, world, n, args.strategy),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": nodes_local * bytes_per_node, "peak_source": peak_src,
                         "kernel": kernel_name, "bytes_per_node": bytes_per_node},
            "clocks": clocks.summary(), "gpu_launches": int(launches), "e2e": e2e}
    line.update(extra)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"], _, _ = reference_cpu(args.config, args.strategy, steps=3, warmup=1, budget_s=20.0,
                                                   want_size=min(n, 256))
    print(json.dumps(line))


def e2e_leg(args, lt, native, flow, sim, dev, world, barrier, nodes_total):
    import torch
    f_host = torch.empty(flow.f.shape, dtype=torch.float32).pin_memory()
    f_host.copy_(flow.f)
    fbytes = f_host.numel() * 4
    if world == 1 and not args.slab:
        # N = 1: one call of the C ABI's host-buffer entry; the result replaces the input in the same pinned buffer
        import ctypes as C
        energy = torch.zeros(args.steps, dtype=torch.float64).pin_memory()
        eng = native.engine_of(sim)
        torch.cuda.synchronize(dev)
        L = native.lib()
        native.check(L.lbm_run_host(C.byref(eng.desc), f_host.data_ptr(), f_host.data_ptr(), 1, None))
        dt = float("inf")
        for _ in range(2):      # host-side noise (page placement, PCIe contention) is large: best of two calls
            f_host.copy_(flow.f)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            native.check(L.lbm_run_host(C.byref(eng.desc), f_host.data_ptr(), f_host.data_ptr(), args.steps,
                                        energy.data_ptr()))
            dt = min(dt, time.perf_counter() - t0)
        assert torch.isfinite(energy).all() and float(energy[-1]) > 0
        how = ("lbm_run_host (C ABI): pinned host f uploaded, K steps with the kinetic energy of every step reduced "
               "inside the step kernel and read back, final f downloaded; wall clock of the faster of two calls")
    else:
        # N > 1: the public Python API on every rank's slab -- upload, Simulation(K) with a global kinetic-energy
        # reporter of interval 1, download.  defer=False: every step's energy is read back to the host inside the
        # timed region (the contract's per-step device-to-host read)
        import torch.distributed as dist
        from lettuce_b200 import slab
        rep = lt.ObservableReporter(slab.GlobalSum(lt.IncompressibleKineticEnergy(flow)), interval=1, out=None,
                                    defer=False)
        sim.reporter.append(rep)
        flow.i = 1                      # skip the step-0 report so exactly K reports fall in the timed region
        barrier()
        t0 = time.perf_counter()
        native.engine_of(sim).load(f_host.to(dev, non_blocking=True))
        sim(args.steps)
        f_host.copy_(flow.f, non_blocking=True)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        assert len(rep.out) == args.steps and rep.out[-1][2] > 0
        sim.reporter.pop()
        how = ("public API per rank: pinned host slab uploaded, Simulation(K) with a global kinetic-energy reporter "
               "(interval 1), final slab downloaded; wall clock, max over ranks")
    return {"value": args.steps * nodes_total / 1e6 / dt, "unit": "MLUPS",
            "h2d_bytes_per_step": fbytes * world / args.steps, "d2h_bytes_per_step": fbytes * world / args.steps + 8,
            "seconds": dt, "note": how + "; the two population transfers are amortised over K steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--size", type=int, default=None, help="nodes per axis per GPU (default: the config's)")
    ap.add_argument("--strategy", default="PRE_STREAMING",
                    choices=["NO_STREAMING", "PRE_STREAMING", "POST_STREAMING", "DOUBLE_STREAMING"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg")
    ap.add_argument("--quick", action="store_true",
                    help="skip the at_256, torch_gpu_reference, native_gpu_reference and c3_512 legs")
    ap.add_argument("--slab", action="store_true",
                    help="with --gpus 1: run the multi-GPU slab kernel (in-kernel lock step) with the rank as its own "
                         "neighbour, e.g. to profile it under ncu")
    ap.add_argument("--leg", default=None, choices=["native_gpu_reference"],
                    help="run one comparison leg alone and print its JSON object (used by the main run)")
    args = ap.parse_args()
    if args.size is None:
        args.size = CONFIGS[args.config]["default_size"]
    if args.leg == "native_gpu_reference":
        print(json.dumps(reference_native_gpu(args.config, args.strategy, args.size, 20, "cuda:0")))
        return
    if args.impl == "reference":
        reference_main(args)
    else:
        gpu_main(args)


if __name__ == "__main__":
    main()
