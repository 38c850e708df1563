"""`python -m lettuce_b200.cli benchmark` -- the measurement harness of the reference's console script
(`lettuce benchmark`, lettuce/cli.py:57-131) on the B200 engine: build a flow, BGK with tau from the
units, run `steps` time steps, print MLUPS.  Same option names and defaults (precision double,
PRE_STREAMING); only the flows and stencils of the hot-path scope are offered."""
from __future__ import annotations

import click
import torch

from . import (BGKCollision, Context, D2Q9, D3Q19, D3Q27, Simulation, StreamingStrategy, TaylorGreenVortex,
               __version__)

FLOWS = {"taylor2d": (TaylorGreenVortex, D2Q9), "taylor3d": (TaylorGreenVortex, D3Q27),
         "taylor3d_d3q19": (TaylorGreenVortex, D3Q19)}


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
    simulation(min(steps, 10))                    # warm-up: context creation, first launches
    mlups = simulation(steps)
    click.echo("Finished {} ({}, {}) for {} steps in {} bit precision with {}. MLUPS: {:10.2f}".format(
        fl.__class__.__name__, fl.stencil.__class__.__name__, ctx.obj["device"], steps,
        str(ctx.obj["dtype"]).replace("torch.float", ""), streaming_strategy, mlups))
    return 0


if __name__ == "__main__":
    main()
