"""Velocity sets (API of lettuce/_stencil.py:14-50 and lettuce/ext/_stencil/*).

The tables are generated from the construction rule of each set and are identical, entry by
entry, to the CUDA-side tables in csrc/lbm_core.cuh (tests/test_host_logic.py checks both
against the oracle's).
"""
from __future__ import annotations

from typing import List

import numpy as np

__all__ = ["Stencil", "TorchStencil", "D2Q9", "D3Q19", "D3Q27"]


class Stencil:
    e: List[List[int]]
    w: List[float]
    opposite: List[int]
    cs: float = 1 / np.sqrt(3.0)

    @property
    def d(self) -> int:
        return len(self.e[0])

    @property
    def q(self) -> int:
        return len(self.e)

    def _finish(self):
        self.opposite = [self.e.index([-c for c in v]) for v in self.e]


def _pairs(vectors):
    out = []
    for v in vectors:
        out += [list(v), [-c for c in v]]
    return out


class D2Q9(Stencil):
    """lettuce/ext/_stencil/d2q9.py:8-10"""

    def __init__(self):
        self.e = [[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1], [1, 1], [-1, 1], [-1, -1], [1, -1]]
        self.w = [4.0 / 9.0] + [1.0 / 9.0] * 4 + [1.0 / 36.0] * 4
        self._finish()


_FACES = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
_EDGES = [(0, 1, 1), (0, 1, -1), (1, 0, 1), (1, 0, -1), (1, 1, 0), (1, -1, 0)]
_CORNERS = [(1, 1, 1), (1, 1, -1), (1, -1, 1), (1, -1, -1)]


class D3Q19(Stencil):
    """lettuce/ext/_stencil/d3q19.py:8-13"""

    def __init__(self):
        self.e = [[0, 0, 0]] + _pairs(_FACES) + _pairs(_EDGES)
        self.w = [1.0 / 3.0] + [1.0 / 18.0] * 6 + [1.0 / 36.0] * 12
        self._finish()


class D3Q27(Stencil):
    """lettuce/ext/_stencil/d3q27.py:8-12"""

    def __init__(self):
        self.e = [[0, 0, 0]] + _pairs(_FACES) + _pairs(_EDGES) + _pairs(_CORNERS)
        self.w = [8.0 / 27.0] + [2.0 / 27.0] * 6 + [1.0 / 54.0] * 12 + [1.0 / 216.0] * 8
        self._finish()


class TorchStencil:
    """The stencil as tensors of the context dtype (lettuce/_stencil.py:31-50; note that, as in
    the reference, `opposite` becomes a float tensor -- indexing uses `Stencil.opposite`)."""
    cs: float = 1 / np.sqrt(3.0)

    def __init__(self, stencil: Stencil, context):
        self.e = context.convert_to_tensor(stencil.e)
        self.w = context.convert_to_tensor(stencil.w)
        self.opposite = context.convert_to_tensor(stencil.opposite)

    @property
    def d(self):
        return self.e.shape[1]

    @property
    def q(self):
        return self.e.shape[0]
