"""Simulation driver (API of lettuce/_simulation.py:17-344).

The step itself is `native.invoke` -- one fused CUDA kernel launch per time step -- installed
on `Simulation._collide_and_stream`, the attribute through which the reference plugs in its
generated extension (lettuce/_simulation.py:229).  User subclasses that re-assign that
attribute keep working.  There is no torch implementation of collide/stream here.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import Enum
from timeit import default_timer as timer
from typing import List, Optional

import torch

from . import native

__all__ = ["Collision", "Reporter", "Simulation", "BreakableSimulation", "StreamingStrategy"]


class StreamingStrategy(Enum):
    """bit 1: stream before the collide phase, bit 0: after it
    (lettuce/cuda_native/_default_code_gen.py:13-25)"""
    NO_STREAMING = 0b00
    PRE_STREAMING = 0b10
    POST_STREAMING = 0b01
    DOUBLE_STREAMING = 0b11

    def pre_streaming(self) -> bool:
        return bool(self.value & 0b10)

    def post_streaming(self) -> bool:
        return bool(self.value & 0b01)


class Collision(ABC):
    """Collision operators are parameter holders for the fused kernel; `collision(flow)` and `native_available`
    keep the reference's operator contract (lettuce/_simulation.py:17-28)."""

    def __call__(self, flow) -> torch.Tensor:
        """post-collision populations of every node of `flow.f` as a new tensor (e.g.
        lettuce/ext/_collision/bgk_collision.py:17-22): one collide-only launch of the step kernel"""
        return native.apply_operator(self, flow, is_collision=True)

    def native_available(self) -> bool:
        try:
            native.op_kind(self)
            return True
        except NotImplementedError:
            return False


class Reporter(ABC):
    interval: int
    # True when the reporter only acts on steps with flow.i % interval == 0, so the driver may run
    # the steps in between without calling it
    batchable = False

    def __init__(self, interval: int):
        self.interval = interval

    @abstractmethod
    def __call__(self, simulation: "Simulation"):
        ...


class Simulation:
    def __init__(self, flow, collision, reporter: List[Reporter],
                 streaming_strategy: StreamingStrategy = StreamingStrategy.POST_STREAMING):
        self.flow = flow
        self.flow.collision = collision
        self.context = flow.context
        self.collision = collision
        self.reporter = reporter
        self.streaming_strategy = streaming_strategy
        # boundaries are read once each; flows build fresh objects per access (obstacle.py:107-122)
        self.pre_boundaries = list(flow.pre_boundaries or [])
        self.post_boundaries = list(flow.post_boundaries or [])
        self.collision_index = len(self.pre_boundaries)
        self.transformer = self.pre_boundaries + [collision] + self.post_boundaries
        self.no_collision_mask: Optional[torch.Tensor] = None
        self.no_streaming_mask: Optional[torch.Tensor] = None
        self._build_masks()
        self._collide_and_stream = native.invoke

    def _build_masks(self):
        """Label field and no-stream mask (lettuce/_simulation.py:100-146): `no_collision_mask`
        holds, per node, the index of the transformer entry acting there (later boundaries win);
        `no_streaming_mask` is the OR of the boundaries' masks.  As in the reference both start
        out filled with `collision_index`."""
        boundaries = self.pre_boundaries + self.post_boundaries
        if not boundaries:
            return
        flow, ctx = self.flow, self.context
        shape = [int(s) for s in flow.f.shape]
        self.no_collision_mask = ctx.full_tensor(shape[1:], self.collision_index, dtype=torch.uint8)
        self.no_streaming_mask = ctx.full_tensor(shape, self.collision_index, dtype=torch.uint8)
        labels = list(range(len(self.pre_boundaries))) + \
            list(range(self.collision_index + 1, self.collision_index + 1 + len(self.post_boundaries)))
        for label, boundary in zip(labels, boundaries):
            ncm, nsm = self._boundary_masks(boundary, shape)
            if ncm is not None:
                self.no_collision_mask[ncm.to(device=ctx.device, dtype=torch.bool)] = label
            if nsm is not None:
                self.no_streaming_mask |= nsm.to(device=ctx.device, dtype=torch.uint8)

    def _boundary_masks(self, boundary, shape):
        """(no_collision_mask, no_streaming_mask) of one boundary; either may be None"""
        return (boundary.make_no_collision_mask(shape[1:], context=self.context),
                boundary.make_no_streaming_mask(shape, context=self.context))

    @property
    def units(self):
        return self.flow.units

    def step(self, num_steps: int):
        return self(num_steps)

    def _report(self):
        for reporter in self.reporter:
            reporter(self)

    def _flush_reporters(self):
        """reporters that keep device values until they are read (ObservableReporter) complete their rows"""
        for reporter in self.reporter:
            flush = getattr(reporter, "flush", None)
            if callable(flush):
                flush()

    def _batch_length(self, remaining: int) -> int:
        """Number of steps that can run before any reporter has to see the state."""
        if self._collide_and_stream is not native.invoke:
            return 1
        k = remaining
        for r in self.reporter:
            if not getattr(r, "batchable", False):
                return 1
            interval = max(int(r.interval), 1)
            k = min(k, interval - self.flow.i % interval)
        return max(k, 1)

    def _reports_ride_on_every_step(self) -> bool:
        """all reporters have interval 1 and accept the step kernels' own reductions for whole batches"""
        if self._collide_and_stream is not native.invoke or not self.reporter or not self.flow.f.is_cuda:
            return False
        if not all(callable(getattr(r, "accepts_fused_batches", None)) and r.accepts_fused_batches(self)
                   for r in self.reporter):
            return False
        engine = native.engine_of(self)
        return type(engine) is native.Engine and engine.moments_state() != native.MOMENTS_UNAVAILABLE

    def _fused_reporters_due(self, k: int):
        """(any, all): `k` steps from now, does ANY due reporter evaluate an observable that the step kernels can
        reduce on the fly (`fused_with_step`: IncompressibleKineticEnergy / MaximumVelocity of this flow), and are
        ALL due reporters of that kind?"""
        if self._collide_and_stream is not native.invoke or not self.flow.f.is_cuda:
            return False, False
        due = [r for r in self.reporter if (self.flow.i + k) % max(int(r.interval), 1) == 0]
        fused = [r for r in due
                 if getattr(getattr(r, "observable", None), "fused_with_step", False)
                 and getattr(r.observable, "flow", None) is self.flow and getattr(r, "batchable", False)]
        return bool(fused), bool(due) and len(fused) == len(due)

    def __call__(self, num_steps: int) -> float:
        """Run `num_steps` time steps; returns MLUPS (lettuce/_simulation.py:311-323).  The
        device is synchronised before the clock is read.

        Steps between two due reports run as one library call (`lbm_step_n`).  Reports of observables the step
        kernels can reduce themselves cost no second pass over the populations: with NO / PRE streaming the step
        that writes the reported state reduces it; with POST streaming the step FOLLOWING the reported state does
        (its input node is the reported node), so when only such reporters are due and more steps follow, that next
        step is launched first and the reporters then receive the values it produced (`flow.i` still names the
        reported step; `flow.f` is one step ahead for the duration of these reporter calls)."""
        self.context.synchronize()
        beg = timer()
        if self.flow.i == 0:
            self._report()
        remaining = int(num_steps)
        ahead = 0                       # 1 when flow.f is already one step ahead of flow.i (see above)
        if remaining >= 2 and self._reports_ride_on_every_step():
            # every reporter wants every step and nothing but moments the step kernels reduce themselves: the whole
            # run is ONE library call (lbm_step_moments_n), the reporters take their rows from its result
            engine = native.engine_of(self)
            moments = engine.steps_with_moments(remaining)
            for reporter in self.reporter:
                reporter.ingest_fused_batch(self, self.flow.i + 1, moments)
            self.flow.i += remaining
            if len(moments) < remaining:            # POST streaming: the last state is reduced on its own
                self._report()
            remaining = 0
        while remaining > 0:
            k = self._batch_length(remaining)
            todo = k - ahead
            engine = native.engine_of(self) if self._collide_and_stream is native.invoke else None
            state = engine.moments_state() if engine is not None else native.MOMENTS_UNAVAILABLE
            any_fused, all_fused = self._fused_reporters_due(k)
            if todo > 0 and any_fused and state == native.MOMENTS_OF_OUTPUT:
                if todo > 1:
                    native.invoke_n(self, todo - 1)
                engine.step_with_moments()
            elif todo == 1 and engine is None:
                self._collide_and_stream(self)
            elif todo > 0:
                if engine is None:
                    for _ in range(todo):
                        self._collide_and_stream(self)
                else:
                    native.invoke_n(self, todo)
            self.flow.i += k
            remaining -= k
            ahead = 0
            if all_fused and state == native.MOMENTS_OF_INPUT and remaining > 0:
                # the next step reduces the moments of the state being reported while it reads it
                self.flow._b200_reported_moments = engine.step_with_moments()
                ahead = 1
                try:
                    self._report()
                finally:
                    self.flow._b200_reported_moments = None
            else:
                self._report()
        self._flush_reporters()
        self.context.synchronize()
        end = timer()
        nodes = 1
        for n in self.flow.resolution:
            nodes *= int(n)
        return num_steps * nodes / 1e6 / (end - beg)


class BreakableSimulation(Simulation):
    """Runs until `flow.i` reaches `num_steps`; reporters may push `flow.i` past it to abort
    (lettuce/_simulation.py:325-344)."""

    def __init__(self, flow, collision, reporter: List[Reporter]):
        super().__init__(flow, collision, reporter)

    def __call__(self, num_steps: int) -> float:
        self.context.synchronize()
        beg = timer()
        if self.flow.i == 0:
            self._report()
        while self.flow.i < num_steps:
            self._collide_and_stream(self)
            self.flow.i += 1
            self._report()
        self._flush_reporters()
        self.context.synchronize()
        end = timer()
        nodes = 1
        for n in self.flow.resolution:
            nodes *= int(n)
        return num_steps * nodes / 1e6 / (end - beg)
