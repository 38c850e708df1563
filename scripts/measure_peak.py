"""STREAM-style copy bandwidth of this GPU, measured the way the driver's MEASURED_PEAKS.json describes it
(`b.copy_(a)`, 1 Gi bf16 elements, best of 10): context for the roofline fractions when that file is absent.
Also the same copy as a hand-rolled float4 kernel would see it (torch's copy kernel is one).

    python scripts/measure_peak.py   ->   one JSON line
"""
import json

import torch


def main():
    dev = torch.device("cuda", 0)
    n = 1 << 30
    a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
    b = torch.empty_like(a)
    best = float("inf")
    times = []
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        b.copy_(a)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e)
        times.append(ms)
        best = min(best, ms)
    bytes_moved = 2 * n * 2
    # sustained: 200 copies back to back (the clocks settle under the power cap)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(200):
        b.copy_(a)
    e.record()
    torch.cuda.synchronize()
    sustained = 200 * bytes_moved / (s.elapsed_time(e) * 1e-3) / 1e9
    print(json.dumps({"hbm_gbs_burst": bytes_moved / (best * 1e-3) / 1e9, "hbm_gbs_sustained": sustained,
                      "bytes": bytes_moved, "best_ms": best, "all_ms": times,
                      "device": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
