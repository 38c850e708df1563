"""Taylor-Green vortex 3-D, D3Q19 BGK fp32, with kinetic-energy and enstrophy reporters
(counterpart of the reference's examples/00_simplest_TGV.py and 03_outputs_TGV.py).

    python examples/00_taylor_green.py [--resolution 256] [--steps 2000] [--dry]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lettuce_b200 as lt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolution", type=int, default=256)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--vtk", default=None, help="directory for .vtr output every 500 steps")
    ap.add_argument("--dry", action="store_true", help="build everything on the CPU and stop before the first step")
    args = ap.parse_args()

    ctx = lt.Context("cpu" if args.dry else "cuda", dtype=torch.float32)
    flow = lt.TaylorGreenVortex(ctx, [args.resolution] * 3, reynolds_number=1600, mach_number=0.05, stencil=lt.D3Q19())
    collision = lt.BGKCollision(tau=flow.units.relaxation_parameter_lu)
    energy = lt.ObservableReporter(lt.IncompressibleKineticEnergy(flow), interval=100, out=None)
    enstrophy = lt.ObservableReporter(lt.Enstrophy(flow), interval=100, out=None)
    reporters = [energy, enstrophy]
    if args.vtk:
        reporters.append(lt.VTKReporter(interval=500, filename_base=os.path.join(args.vtk, "tgv")))
    # PRE_STREAMING: the step before every energy report reduces the energy inside the step kernel
    simulation = lt.Simulation(flow, collision, reporters, lt.StreamingStrategy.PRE_STREAMING)
    if args.dry:
        print("dry run: built", type(flow).__name__, flow.resolution, "tau =", flow.units.relaxation_parameter_lu)
        return
    mlups = simulation(args.steps)
    print(f"{mlups:.0f} MLUPS")
    for (step, t, e), (_, _, w) in zip(energy.out, enstrophy.out):
        print(f"step {step:6d}  t = {t:7.3f}  E_kin = {e:.6f}  enstrophy = {w:.6f}")
    spectrum = lt.EnergySpectrum(flow)()
    print("E(k), k = 0..9:", [f"{v:.3e}" for v in spectrum[:10].tolist()])


if __name__ == "__main__":
    main()
