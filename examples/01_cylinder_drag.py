"""Flow around a circular cylinder with link-wise bounce-back applied after streaming and the drag / lift
coefficients from the momentum exchange on the boundary links (counterpart of the reference's
examples/advanced_projects/efficient_bounce_back_obstacle/01_script_cylinder_simulation.py).

    python examples/01_cylinder_drag.py [--bc ibb1|hwbb|fwbb] [--diameter 20] [--steps 20000] [--dry]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lettuce_b200 as lt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bc", default="ibb1", choices=["ibb1", "hwbb", "fwbb"])
    ap.add_argument("--diameter", type=int, default=20, help="cylinder diameter in lattice nodes")
    ap.add_argument("--domain", type=int, nargs=2, default=[30, 10], help="domain length and height in diameters")
    ap.add_argument("--reynolds", type=float, default=100.0)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--dry", action="store_true", help="build everything on the CPU and stop before the first step")
    args = ap.parse_args()

    ctx = lt.Context("cpu" if args.dry else "cuda", dtype=torch.float64)
    resolution = [args.domain[0] * args.diameter, args.domain[1] * args.diameter]
    flow = lt.ObstacleCylinder(ctx, resolution, args.reynolds, 0.05, char_length_pu=1.0,
                               char_length_lu=float(args.diameter), bc_type=args.bc, lateral_walls="periodic",
                               u_init=1, perturb_init=True, calc_force_coefficients=True, stencil=lt.D2Q9())
    simulation = lt.EbbSimulation(flow, lt.BGKCollision(flow.units.relaxation_parameter_lu), [])
    cylinder = simulation.post_streaming_boundaries[-1]
    drag = lt.ObservableReporter(lt.DragCoefficient(flow, cylinder, flow.solid_mask, area_pu=1.0), interval=100, out=None)
    lift = lt.ObservableReporter(lt.LiftCoefficient(flow, cylinder, flow.solid_mask, area_pu=1.0), interval=100, out=None)
    simulation.reporter += [drag, lift]
    print(f"{type(cylinder).__name__}: {cylinder.n_links} links, tau = {flow.units.relaxation_parameter_lu:.4f}")
    if args.dry:
        return
    mlups = simulation(args.steps)
    print(f"{mlups:.0f} MLUPS")
    tail = drag.out[len(drag.out) // 2:]
    print("mean drag coefficient over the second half:", sum(r[2] for r in tail) / len(tail))
    print("last lift coefficients:", [round(r[2], 4) for r in lift.out[-5:]])


if __name__ == "__main__":
    main()
