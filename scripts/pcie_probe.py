"""Host <-> device copy bandwidth of every visible GPU, one at a time and all together (one process, one stream per
device, pinned host buffers): explains the `e2e` figures of bench.py at N > 1, where every rank moves its whole slab
over PCIe twice.  Prints one JSON line.

    python scripts/pcie_probe.py [GiB per buffer, default 2]
"""
import json
import subprocess
import sys
import time

import torch


def timed(copies, devices):
    for d in devices:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for c in copies:
        c()
    for d in devices:
        torch.cuda.synchronize(d)
    return time.perf_counter() - t0


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
    n = int(gib * 2 ** 30)
    devices = list(range(torch.cuda.device_count()))
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in devices]
    dev = [torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}") for d in devices]
    streams = [torch.cuda.Stream(device=d) for d in devices]

    def h2d(i):
        def run():
            with torch.cuda.stream(streams[i]):
                dev[i].copy_(host[i], non_blocking=True)
        return run

    def d2h(i):
        def run():
            with torch.cuda.stream(streams[i]):
                host[i].copy_(dev[i], non_blocking=True)
        return run

    out = {"gib_per_buffer": gib, "devices": len(devices), "h2d_alone_gbs": [], "d2h_alone_gbs": []}
    for i in devices:
        timed([h2d(i)], [i])                                           # warm-up
        out["h2d_alone_gbs"].append(round(n / 1e9 / min(timed([h2d(i)], [i]) for _ in range(3)), 1))
        out["d2h_alone_gbs"].append(round(n / 1e9 / min(timed([d2h(i)], [i]) for _ in range(3)), 1))
    if len(devices) > 1:
        t = min(timed([h2d(i) for i in devices], devices) for _ in range(3))
        out["h2d_all_together_gbs_total"] = round(len(devices) * n / 1e9 / t, 1)
        t = min(timed([d2h(i) for i in devices], devices) for _ in range(3))
        out["d2h_all_together_gbs_total"] = round(len(devices) * n / 1e9 / t, 1)
    try:
        out["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True,
                                         timeout=30).stdout.splitlines()[:len(devices) + 1]
    except Exception as e:
        out["topology"] = str(e)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
