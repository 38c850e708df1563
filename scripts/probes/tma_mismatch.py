"""Debug probe: where does the TMA-staged kernel differ from the one-node LDG kernel?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import lettuce_b200 as lt
from lettuce_b200 import native as nv

def run(res, variant, steps, stencil=lt.D3Q19):
    c = lt.Context("cuda:0", dtype=torch.float32)
    flow = lt.TaylorGreenVortex(c, res, 1600.0, 0.05, stencil=stencil())
    gen = torch.Generator(device=flow.f.device).manual_seed(11)
    flow.f.mul_(1.0 + 1e-2 * (torch.rand(flow.f.shape, generator=gen, device=flow.f.device) - 0.5))
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], lt.StreamingStrategy.PRE_STREAMING)
    eng = nv.engine_of(sim)
    eng.desc.variant = variant
    for _ in range(steps):
        nv.invoke_n(sim, 1) if os.environ.get("SINGLE") else None
    if not os.environ.get("SINGLE"):
        nv.invoke_n(sim, steps)
    torch.cuda.synchronize()
    return flow.f.clone()

res = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [12, 32, 256]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ref = run(res, 1, steps)
bad_runs = 0
for trial in range(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    got = run(res, 3, steps)
    diff = (got != ref).nonzero()
    if len(diff):
        bad_runs += 1
        q, x, y, z = diff.unbind(1)
        print(f"trial {trial}: {len(diff)} mismatches; q {sorted(set(q.tolist()))[:30]} x {sorted(set(x.tolist()))[:20]} "
              f"y {sorted(set(y.tolist()))[:40]} z [{int(z.min())}..{int(z.max())}] distinct z {len(set(z.tolist()))}")
        rows = sorted(set((int(a), int(b)) for a, b in zip(x.tolist(), y.tolist())))
        print("   rows (x,y):", rows[:24], "n =", len(rows))
        i = diff[0]
        print("   first:", i.tolist(), float(got[tuple(i)]), float(ref[tuple(i)]))
print("bad runs", bad_runs)
