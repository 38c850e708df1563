"""`python -m lettuce_b200.cli {benchmark,convergence}` -- the two commands of the reference's console script
(lettuce/cli.py:57-186) on the B200 engine.  `benchmark`: build a flow, BGK with tau from the units, run
`steps` time steps, print MLUPS (same option names and defaults: precision double, PRE_STREAMING; only the
flows and stencils of the hot-path scope are offered).  `convergence`: Taylor-Green 2-D in diffusive scaling,
checks second-order convergence of the velocity and first-order of the pressure."""
from __future__ import annotations

import click
import torch

import sys

import numpy as np

from . import (BGKCollision, Context, D2Q9, D3Q19, D3Q27, ErrorReporter, Simulation, StreamingStrategy,
               TaylorGreenVortex, __version__, flow_by_name)

# the reference's registry (lettuce/ext/_flows/_flow_by_name.py) plus a short name for the D3Q27 vortex
FLOWS = {**flow_by_name, "taylor3d": (TaylorGreenVortex, D3Q27)}


@click.group()
@click.version_option(version=__version__)
@click.option("-i", "--gpu-id", type=int, default=0, help="Device ID of the GPU (default=0).")
@click.option("-p", "--precision", type=click.Choice(["single", "double"]), default="double",
              help="Numerical precision, 32 or 64 bit per float (default=double).")
@click.pass_context
def main(ctx, gpu_id, precision):
    """B200-native lattice Boltzmann stream+collide engine (CUDA only)."""
    if not torch.cuda.is_available():
        click.echo("CUDA not found; lettuce_b200 has no CPU path.")
        raise click.Abort
    ctx.obj = {"device": torch.device(f"cuda:{gpu_id}"),
               "dtype": {"single": torch.float32, "double": torch.float64}[precision]}


@main.command()
@click.option("-s", "--steps", type=int, default=10, help="Number of time steps.")
@click.option("-r", "--resolution", type=int, default=1024, help="Grid resolution per axis.")
@click.option("-f", "--flow", type=click.Choice(list(FLOWS)), default="taylor2d")
@click.option("--streaming-strategy", type=click.Choice(["PRE_STREAMING", "POST_STREAMING"]),
              default="PRE_STREAMING", help="Streaming strategy (default=PRE_STREAMING).")
@click.pass_context
def benchmark(ctx, steps, resolution, flow, streaming_strategy):
    """Run a short simulation and print performance in MLUPS."""
    flow_class, stencil = FLOWS[flow]
    context = Context(ctx.obj["device"], ctx.obj["dtype"])
    fl = flow_class(context, resolution=[resolution] * stencil().d, reynolds_number=1, mach_number=0.05,
                    stencil=stencil)
    collision = BGKCollision(tau=fl.units.relaxation_parameter_lu)
    simulation = Simulation(fl, collision, [], streaming_strategy=getattr(StreamingStrategy, streaming_strategy))
    simulation(min(steps, 64))                    # warm-up: context creation, first launches, graph capture
    mlups = simulation(steps)
    click.echo("Finished {} ({}, {}) for {} steps in {} bit precision with {}. MLUPS: {:10.2f}".format(
        fl.__class__.__name__, fl.stencil.__class__.__name__, ctx.obj["device"], steps,
        str(ctx.obj["dtype"]).replace("torch.float", ""), streaming_strategy, mlups))
    return 0


def run_convergence(context, exponents=range(4, 9), echo=print):
    """(order_u, order_p) of the last refinement; lettuce/cli.py:134-186"""
    echo(("{:>15} " * 6).format("resolution", "error (u)", "order (u)", "error (p)", "order (p)", "MLUPS"))
    old_u = old_p = None
    factor_u = factor_p = 0.0
    for i in exponents:
        resolution = 2 ** i
        flow = TaylorGreenVortex(context, [resolution] * 2, reynolds_number=10000, mach_number=8 / resolution,
                                 stencil=D2Q9())
        reporter = ErrorReporter(flow.analytic_solution, interval=1, out=None)
        simulation = Simulation(flow, BGKCollision(tau=flow.units.relaxation_parameter_lu), [reporter])
        mlups = simulation(10 * resolution)
        error_u, error_p = np.mean(np.abs(reporter.out), axis=0).tolist()
        factor_u = 0 if old_u is None else old_u / error_u
        factor_p = 0 if old_p is None else old_p / error_p
        old_u, old_p = error_u, error_p
        echo(f"{resolution:15} {error_u:15.2e} {factor_u / 2:15.2f} {error_p:15.2e} {factor_p / 2:15.2f} {mlups:15.2f}")
    return factor_u / 2, factor_p / 2


@main.command()
@click.pass_context
def convergence(ctx):
    """Use Taylor Green 2D for convergence test in diffusive scaling."""
    order_u, order_p = run_convergence(Context(ctx.obj["device"], ctx.obj["dtype"]), echo=click.echo)
    tol = 1e-1
    if not (2 - tol) < order_u < (2 + tol):
        click.echo(f"FAILED: Velocity convergence order {order_u} is not in [1.9, 2.1]")
        sys.exit(1)
    if not (1 - tol) < order_p < (1 + tol):
        click.echo(f"FAILED: Pressure convergence order {order_p} is not in [0.9, 1.1].")
        sys.exit(1)
    return 0


if __name__ == "__main__":
    main()
