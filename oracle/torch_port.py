"""Torch restatement of the reference's torch path for ONE configuration family (BGK, no boundaries), written
op for op the way lettuce executes a step (TEST / BASELINE INFRASTRUCTURE ONLY -- see oracle/lbm_oracle.py).

Purpose: bench.py times it on the B200 as the stand-in for "the reference's torch GPU run" (north_star asks for
the numbers side by side; the reference's sources cannot travel to the GPU box).  It performs the same sequence of
full-size tensor operations as lettuce: density by `sum`, momentum by `einsum`, the quadratic equilibrium as a
chain of elementwise temporaries (lettuce/ext/_equilibrium/quadratic_equilibrium.py:11-24), BGK relaxation
(lettuce/ext/_collision/bgk_collision.py:17-22) and streaming as one `torch.roll` per population written back
into the population tensor (lettuce/_simulation.py:241-256).  Checked against the NumPy oracle in
tests/test_oracle_golden.py.
"""
import torch

from . import lbm_oracle as lo


class TorchBGK:
    def __init__(self, stencil_name: str, tau: float, device, dtype):
        st = lo.stencil(stencil_name)
        self.d, self.q = st["d"], st["q"]
        self.e_list = [tuple(int(c) for c in v) for v in st["e"]]
        self.e = torch.tensor(st["e"], device=device, dtype=dtype)
        self.w = torch.tensor(st["w"], device=device, dtype=dtype)
        self.tau = tau
        self.cs2 = lo.CS2

    def equilibrium(self, rho, u):
        exu = torch.tensordot(self.e, u, dims=1)
        uxu = torch.einsum("d...,d...->...", [u, u])
        return torch.einsum("q...,q...->q...",
                            [self.w, rho * ((2 * exu - uxu) / (2 * self.cs2) + 0.5 * (exu / self.cs2) ** 2 + 1)])

    def collide(self, f):
        rho = torch.sum(f, dim=0)[None, ...]
        u = torch.einsum("qd,q...->d...", [self.e, f]) / rho
        feq = self.equilibrium(torch.sum(f, dim=0)[None, ...], u)     # the reference recomputes rho here
        return f - 1.0 / self.tau * (f - feq)

    def stream(self, f):
        dims = tuple(range(self.d))
        for i in range(1, self.q):
            f[i] = torch.roll(f[i], shifts=self.e_list[i], dims=dims)
        return f

    def step(self, f, strategy="PRE_STREAMING"):
        pre, post = lo.STRATEGIES[strategy]
        if pre:
            f = self.stream(f)
        f = self.collide(f)
        if post:
            f = self.stream(f)
        return f
