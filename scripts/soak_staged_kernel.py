"""Soak test of the TMA-staged kernel: thousands of steps on BASELINE-sized lattices, compared bit for bit with the
LDG kernel at the end (rare races in the barrier ring / tile claiming would show up as a single differing element).

    python scripts/soak_staged_kernel.py [steps]
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
from lettuce_b200 import native as nv  # noqa: E402
from bench_configs import make_obstacle  # noqa: E402


def run(make, variant, steps, batch):
    flow, sim = make()
    eng = nv.engine_of(sim)
    eng.desc.variant = variant
    name = eng.lib.lbm_step_variant_name(eng.desc).decode()
    done = 0
    while done < steps:
        k = min(batch, steps - done)
        nv.invoke_n(sim, k)
        done += k
    torch.cuda.synchronize()
    return flow.f, name


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    ctx = lt.Context("cuda", dtype=torch.float32)
    S = lt.StreamingStrategy

    def tgv_kbc():
        flow = lt.TaylorGreenVortex(ctx, [256] * 3, 1600.0, 0.05, stencil=lt.D3Q27())
        return flow, lt.Simulation(flow, lt.KBCCollision(), [], S.PRE_STREAMING)

    def sphere_trt():
        flow = make_obstacle(ctx, [256, 128, 128], lt.D3Q27())
        return flow, lt.Simulation(flow, lt.TRTCollision(flow.units.relaxation_parameter_lu), [], S.PRE_STREAMING)

    def tgv2d_bgk():
        flow = lt.TaylorGreenVortex(ctx, [2048, 1024], 1600.0, 0.05, stencil=lt.D2Q9())
        return flow, lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], S.PRE_STREAMING)

    def tgv_kbc_post():
        flow = lt.TaylorGreenVortex(ctx, [256] * 3, 1600.0, 0.05, stencil=lt.D3Q27())
        return flow, lt.Simulation(flow, lt.KBCCollision(), [], S.POST_STREAMING)

    def sphere_trt_post():
        flow = make_obstacle(ctx, [256, 128, 128], lt.D3Q27())
        return flow, lt.Simulation(flow, lt.TRTCollision(flow.units.relaxation_parameter_lu), [], S.POST_STREAMING)

    ok = True
    for name, make, n in (("TGV3D D3Q27 KBC 256^3", tgv_kbc, steps), ("sphere D3Q27 TRT 256x128x128", sphere_trt, steps),
                          ("TGV2D D2Q9 BGK 2048x1024", tgv2d_bgk, steps),
                          ("TGV3D D3Q27 KBC 256^3 POST (pushing kernel)", tgv_kbc_post, steps),
                          ("sphere D3Q27 TRT 256x128x128 POST (pushing kernel)", sphere_trt_post, steps)):
        t0 = time.perf_counter()
        ref, ref_name = run(make, 2, n, 13)
        got, got_name = run(make, 3, n, 11)              # different batch lengths: different launch chaining
        same = bool(torch.equal(ref, got))
        finite = bool(torch.isfinite(got).all())
        ok = ok and same and finite
        print(json.dumps(dict(case=name, steps=n, bit_identical=same, finite=finite, reference=ref_name, staged=got_name,
                              seconds=round(time.perf_counter() - t0, 1))), flush=True)
        del ref, got
        torch.cuda.empty_cache()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
