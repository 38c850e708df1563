"""What does a kinetic-energy report after EVERY step cost on top of the bare step?  (VERDICT round 1, item 9:
`IncompressibleKineticEnergy` with interval 1 on C2 POST and C4.)  The report rides on the step kernels
(`lbm_step_moments`: sum of 0.5 |u|^2 and max |u|^2 reduced inside the step, masked runs included), values stay on the
device until `reporter.out` is read.  Prints one JSON line per case.

    python scripts/bench_reporter_overhead.py [--small]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import lettuce_b200 as lt  # noqa: E402
from bench_configs import make_obstacle  # noqa: E402


def timed(sim, steps):
    sim(10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    sim(steps)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    small = "--small" in sys.argv
    S = lt.StreamingStrategy
    ctx = lt.Context("cuda", dtype=torch.float32)
    n = 256 if small else 512
    cases = []
    for strat in (S.POST_STREAMING, S.PRE_STREAMING):
        cases.append((f"C2 TGV3D D3Q19 BGK {n}^3", strat, 100,
                      lambda: lt.TaylorGreenVortex(ctx, [n] * 3, 1600.0, 0.05, stencil=lt.D3Q19()),
                      lambda f: lt.BGKCollision(f.units.relaxation_parameter_lu)))
        cases.append(("C4 cylinder D2Q9 BGK 4096x1024", strat, 400,
                      lambda: make_obstacle(ctx, [4096, 1024], lt.D2Q9()),
                      lambda f: lt.BGKCollision(f.units.relaxation_parameter_lu)))
        cases.append((f"C3 TGV3D D3Q27 KBC {n // 2}^3", strat, 100,
                      lambda: lt.TaylorGreenVortex(ctx, [n // 2] * 3, 1600.0, 0.05, stencil=lt.D3Q27()),
                      lambda f: lt.KBCCollision()))
    for name, strat, steps, make_flow, make_coll in cases:
        flow = make_flow()
        bare = timed(lt.Simulation(flow, make_coll(flow), [], strat), steps)
        del flow
        torch.cuda.empty_cache()
        flow = make_flow()
        rep = lt.ObservableReporter(lt.IncompressibleKineticEnergy(flow), interval=1, out=None)
        sim = lt.Simulation(flow, make_coll(flow), [rep], strat)
        with_report = timed(sim, steps)
        t0 = time.perf_counter()
        values = rep.out
        fetch = time.perf_counter() - t0
        print(json.dumps(dict(case=name, streaming=strat.name, steps=steps, ms_bare=bare, ms_with_energy_every_step=with_report,
                              overhead=with_report / bare - 1.0, reports=len(values), last=values[-1][2],
                              fetch_all_values_s=fetch)), flush=True)
        del flow, sim, rep
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
