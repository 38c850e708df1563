"""C1 (TGV2D D2Q9 BGK 256^2 fp64) step rate with plain launches vs CUDA-graph replay (LBM_B200_GRAPH_MAX_NODES).

    python scripts/graph_c1.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lettuce_b200 as lt  # noqa: E402
from lettuce_b200 import native  # noqa: E402

ctx = lt.Context("cuda:0", dtype=torch.float64)
for res, stencil in (([256, 256], lt.D2Q9), ([64, 64, 64], lt.D3Q19)):
    flow = lt.TaylorGreenVortex(ctx, res, 1.0, 0.05, stencil=stencil())
    sim = lt.Simulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [], lt.StreamingStrategy.PRE_STREAMING)
    nodes = flow.f[0].numel()
    for limit in ("0", "1000000", "0", "1000000"):
        os.environ["LBM_B200_GRAPH_MAX_NODES"] = limit
        native.invoke_n(sim, 128)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        native.invoke_n(sim, 2048)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / 2048
        print(f"{res} graph_max_nodes={limit:>8}: {us:6.2f} us/step, {nodes / us:9.1f} MLUPS", flush=True)
