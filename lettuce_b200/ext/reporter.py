"""Observables and the observable reporter (API of lettuce/ext/_reporter/observable_reporter.py).

Every observable is ONE pass over the populations on the CUDA engine (`lbm_reduce`: warp-shuffle
block reductions, deterministic two-stage fold) instead of the reference's chain of full-size
torch temporaries; results stay on the device as 0-d float64 tensors until a reporter reads them.
"""
from __future__ import annotations

import sys
from abc import ABC, abstractmethod
from typing import Optional

import torch

from .. import native
from .._simulation import Reporter

__all__ = ["Observable", "ObservableReporter", "MaximumVelocity", "IncompressibleKineticEnergy",
           "Enstrophy", "EnergySpectrum", "Mass", "ProgressReporter", "write_image", "FailureReporterBase", "NaNReporter", "HighMaReporter", "ErrorReporter"]


class Observable(ABC):
    def __init__(self, flow):
        self.context = flow.context
        self.flow = flow

    @abstractmethod
    def __call__(self, f: Optional[torch.Tensor] = None):
        ...


class MaximumVelocity(Observable):
    """max |u| in physical units (observable_reporter.py:27-31).  `fused_with_step`: see IncompressibleKineticEnergy."""
    fused_with_step = True

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        fused = native.fused_moments(self.flow, f)            # reduced inside a step kernel?
        u_max = torch.sqrt(fused[1]) if fused is not None else native.reduce(self.flow.stencil, native.MAX_U, f)
        return self.flow.units.convert_velocity_to_pu(u_max)

    def from_fused_moments(self, moments):
        """values of a whole batch of steps from the rows (sum 0.5|u|^2, max |u|^2) the step kernels reduced"""
        return self.flow.units.convert_velocity_to_pu(torch.sqrt(moments[:, 1]))


class IncompressibleKineticEnergy(Observable):
    """sum 0.5 |u|^2 dx^d in physical units (observable_reporter.py:34-42).  `fused_with_step`: when a reporter
    of this observable is due, `Simulation.__call__` lets a step kernel reduce it (`lbm_step_moments`: the step that
    writes the reported state, or with POST_STREAMING the step that follows it) -- no second pass over the
    populations."""
    fused_with_step = True

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        units = self.flow.units
        fused = native.fused_moments(self.flow, f)            # reduced inside a step kernel?
        e_lu = fused[0] if fused is not None else native.reduce(self.flow.stencil, native.SUM_HALF_U2, f)
        return units.convert_incompressible_energy_to_pu(e_lu) * units.convert_length_to_pu(1.0) ** self.flow.stencil.d

    def from_fused_moments(self, moments):
        units = self.flow.units
        return (units.convert_incompressible_energy_to_pu(moments[:, 0])
                * units.convert_length_to_pu(1.0) ** self.flow.stencil.d)


class Enstrophy(Observable):
    """sum |curl u|^2 dx^d in physical units with 6th-order periodic differences
    (observable_reporter.py:45-68).  Only meaningful on periodic domains."""

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        units, st = self.flow.units, self.flow.stencil
        _, u = native.moments(st, f, want_rho=False)
        w2_lu = native.reduce(st, native.ENSTROPHY, u)          # lattice units, dx = 1
        dx = units.convert_length_to_pu(1.0)
        scale = units.convert_velocity_to_pu(1.0) / dx           # d(u_pu)/d(x_pu) per d(u_lu)/d(x_lu)
        return w2_lu * scale ** 2 * dx ** st.d


class EnergySpectrum(Observable):
    """Kinetic energy spectrum E(k), k = 0 .. int(max|k|)-1: shell sums of 0.5 |fft(u_pu) / norm|^2 over
    k-0.5 < |k| <= k+0.5 (observable_reporter.py:71-137).  The velocity field comes from the engine's moment
    kernel and the transform from cuFFT (`torch.fft.fftn`; a library call, not part of the step path).  Instead of
    the reference's boolean shell mask `[*resolution, kmax]` (3.7 GB at 256^3) every node stores one shell index
    and the shells are summed with one `bincount` -- O(N) memory."""

    def __init__(self, flow):
        super().__init__(flow)
        self.dx = flow.units.convert_length_to_pu(1.0)
        self.dimensions = [int(n) for n in flow.resolution]
        dev = flow.context.device
        k2 = None
        for axis, n in enumerate(self.dimensions):
            k = torch.fft.fftfreq(n, d=1.0 / n, dtype=torch.float64, device=dev)
            shape = [1] * len(self.dimensions)
            shape[axis] = n
            k2 = (k * k).reshape(shape) if k2 is None else k2 + (k * k).reshape(shape)
        wavenorms = torch.sqrt(k2)
        if flow.stencil.d == 3:                      # normalisation by the FIRST resolution entry, as in :86-89
            self.norm = self.dimensions[0] * (2 * torch.pi) ** 0.5 / self.dx ** 2
        else:
            self.norm = self.dimensions[0] / self.dx
        self.n_shells = int(torch.max(wavenorms))
        self.wavenumbers = torch.arange(self.n_shells)
        # k - 0.5 < |k| <= k + 0.5  <=>  k = ceil(|k| - 0.5); shells >= n_shells fall into a discarded last bin
        self.shell = torch.clamp(torch.ceil(wavenorms - 0.5), 0, self.n_shells).to(torch.int64).flatten()

    def __call__(self, f=None):
        return self.spectrum_from_u(self.flow.u())

    def spectrum_from_u(self, u):
        u = self.flow.units.convert_velocity_to_pu(u)
        dims = tuple(range(1, u.dim()))
        uh = torch.fft.fftn(u, dim=dims) / self.norm
        ekin = 0.5 * (uh.real ** 2 + uh.imag ** 2).sum(dim=0)
        ek = torch.bincount(self.shell.to(ekin.device), weights=ekin.flatten(), minlength=self.n_shells + 1)
        return ek[:self.n_shells]


class Mass(Observable):
    """total mass in lattice units; like the reference it skips the first and last index of the
    last two axes and subtracts the populations on `no_mass_mask` (observable_reporter.py:140-158)"""

    def __init__(self, flow, no_mass_mask=None):
        super().__init__(flow)
        self.mask = no_mass_mask

    def __call__(self, f=None):
        f = self.flow.f if f is None else f
        mass = native.reduce(self.flow.stencil, native.SUM_F_INNER, f)
        if self.mask is not None:
            mass = mass - native.reduce(self.flow.stencil, native.SUM_F_MASKED, f, self.mask)
        return mass


class ObservableReporter(Reporter):
    """Evaluates `observable` every `interval` steps and prints or stores
    `[step, time_pu, value...]` (observable_reporter.py:161-200).

    The reference converts every value to NumPy at once, which on a GPU is a device synchronisation per report
    (`convert_to_ndarray`, observable_reporter.py:189-190).  When the rows are collected in a list (`out=None`)
    and the value lives on a CUDA device, the device tensor is kept instead and all pending values are brought to
    the host in ONE transfer when `reporter.out` is read or the running `Simulation.__call__` returns, so a reporter
    with a short interval no longer stalls the launch queue.  `defer` forces that behaviour on (True) or off (False)."""
    batchable = True

    def __init__(self, observable, interval=1, out=sys.stdout, defer: Optional[bool] = None):
        super().__init__(interval)
        self.observable = observable
        self._rows = [] if out is None else out
        self._pending = []                   # (row index, device tensor) of rows whose values are still on the device
        self._defer = defer
        self._parameter_name = observable.__class__.__name__
        if out is not None:
            print("steps    ", "time    ", self._parameter_name)

    @property
    def out(self):
        """the list of rows (all values on the host) or the stream that was passed in"""
        self.flush()
        return self._rows

    @out.setter
    def out(self, value):
        self.flush()
        self._rows = value

    def flush(self):
        """bring the values that are still on the device to the host and complete their rows (in place).  Called
        when `out` is read and by `Simulation.__call__` before it returns, so a list obtained from `reporter.out`
        earlier is complete after every run."""
        if not self._pending:
            return
        pending, self._pending = self._pending, []
        flat = torch.cat([v.reshape(-1).to(torch.float64) for _, v, _ in pending]).cpu().tolist()
        pos = 0
        for index, v, per_row in pending:
            n = v.numel()
            if per_row:                       # a batch: one value for each of n consecutive rows
                for k in range(n):
                    self._rows[index + k] = self._rows[index + k][:2] + [flat[pos + k]]
            else:
                self._rows[index] = self._rows[index][:2] + flat[pos:pos + n]
            pos += n

    def accepts_fused_batches(self, simulation) -> bool:
        """can `Simulation.__call__` hand this reporter the values of a whole batch of steps at once
        (`ingest_fused_batch`)?  Every step is reported, the observable is one the step kernels reduce themselves, and
        the rows are collected in a list with their values left on the device until they are read."""
        obs = self.observable
        return (int(self.interval) == 1 and isinstance(self._rows, list) and self._defer is not False
                and getattr(obs, "fused_with_step", False) and getattr(obs, "flow", None) is simulation.flow
                and hasattr(obs, "from_fused_moments") and simulation.flow.f.is_cuda)

    def ingest_fused_batch(self, simulation, first_step: int, moments):
        """rows for steps first_step .. first_step + len(moments) - 1 from the moments the step kernels reduced"""
        values = self.observable.from_fused_moments(moments).detach()
        units = simulation.units
        start = len(self._rows)
        self._rows.extend([i, units.convert_time_to_pu(i)] for i in range(first_step, first_step + len(moments)))
        self._pending.append((start, values, True))

    def __call__(self, simulation):
        if simulation.flow.i % self.interval != 0:
            return
        value = self.observable(simulation.flow.f)
        head = [simulation.flow.i, simulation.units.convert_time_to_pu(simulation.flow.i)]
        defer = self._defer if self._defer is not None else (torch.is_tensor(value) and value.is_cuda)
        if defer and isinstance(self._rows, list) and torch.is_tensor(value):
            assert value.dim() < 2
            self._rows.append(head)
            self._pending.append((len(self._rows) - 1, value.detach().clone(), False))
            return
        observed = self.observable.context.convert_to_ndarray(value)
        assert len(observed.shape) < 2
        observed = [observed.item()] if len(observed.shape) == 0 else observed.tolist()
        entry = head + observed
        if isinstance(self._rows, list):
            self._rows.append(entry)
        else:
            print(*entry, file=self._rows)


class FailureReporterBase(Reporter):
    """Detects a failing simulation every `interval` steps and aborts a `BreakableSimulation` by pushing
    `flow.i` past any step target (lettuce/ext/_reporter/failure_reporter.py:12-56).  The test itself is one
    reduction on the engine; the list of the `k` worst locations is only assembled after a failure.  A log
    file is written when `outdir` is given (VTK output is outside the hot path)."""
    batchable = True
    name = "Failure"

    def __init__(self, interval: int, k: int = 100, outdir: Optional[str] = None):
        super().__init__(interval)
        self.k = k
        self.outdir = outdir
        self.failed_iteration = None
        self.results = None

    def __call__(self, simulation):
        if simulation.flow.i % self.interval != 0 or not self.is_failed(simulation):
            return
        self.results = self.get_results(simulation)
        self.failed_iteration = simulation.flow.i
        if self.outdir is not None:
            import os
            os.makedirs(self.outdir, exist_ok=True)
            with open(os.path.join(self.outdir, f"{self.name}_reporter.log"), "w") as fh:
                fh.write(f"{self.name} detected at iteration {simulation.flow.i}\n")
                for pos, val in self.results:
                    fh.write(f"{pos} {val}\n")
        print(f"(!) ABORT MESSAGE: {self.name}Reporter detected {self.name} (reporter-interval = {self.interval}) "
              f"at iteration {simulation.flow.i}.")
        simulation.flow.i = int(simulation.flow.i + 1e10)

    def _top(self, mask: torch.Tensor, values: torch.Tensor):
        bad = values[mask]
        coords = torch.nonzero(mask)
        n = min(self.k, bad.numel())
        idx = torch.arange(n, device=values.device) if torch.isnan(bad).any() else torch.topk(bad, k=n).indices
        return [(list(map(int, c)), float(v)) for c, v in zip(coords[idx].cpu().numpy(), bad[idx].cpu().numpy())]

    @abstractmethod
    def is_failed(self, simulation) -> bool:
        ...

    @abstractmethod
    def get_results(self, simulation):
        ...


class NaNReporter(FailureReporterBase):
    """aborts when any population is NaN (failure_reporter.py:131-153).  NaN propagates through the sum of
    all populations, so the test is the engine's SUM_F reduction."""
    name = "NaN"

    def is_failed(self, simulation) -> bool:
        total = native.reduce(simulation.flow.stencil, native.SUM_F, simulation.flow.f)
        return bool(torch.isnan(total).cpu())

    def get_results(self, simulation):
        return self._top(torch.isnan(simulation.flow.f), simulation.flow.f)


class HighMaReporter(FailureReporterBase):
    """aborts when the local Mach number |u|/cs exceeds `threshold` anywhere (failure_reporter.py:156-184)"""
    name = "HighMa"

    def __init__(self, interval: int, threshold: float = 0.3, k: int = 100, outdir: Optional[str] = None):
        super().__init__(interval, k, outdir)
        self.threshold = threshold

    def is_failed(self, simulation) -> bool:
        flow = simulation.flow
        umax = native.reduce(flow.stencil, native.MAX_U, flow.f)
        return bool((umax / flow.stencil.cs > self.threshold).cpu()) or bool(torch.isnan(umax).cpu())

    def get_results(self, simulation):
        flow = simulation.flow
        ma = torch.norm(flow.u(), dim=0) / flow.stencil.cs
        return self._top(ma > self.threshold, ma)


class ErrorReporter(Reporter):
    """L2 error of velocity and pressure (physical units) against an analytic solution, normalised by
    resolution^(d/2) (lettuce/ext/_reporter/error_reporter.py:9-45).  The fields come from the engine's
    moment kernel; the norms are two small torch reductions."""
    batchable = True

    def __init__(self, analytical_solution, interval=1, out=sys.stdout):
        super().__init__(interval)
        self.analytical_solution = analytical_solution
        self.out = [] if out is None else out
        if not isinstance(self.out, list):
            print("#error_u         error_p", file=self.out)

    def __call__(self, simulation):
        flow = simulation.flow
        if flow.i % self.interval != 0:
            return
        pref, uref = self.analytical_solution(t=simulation.units.convert_time_to_pu(flow.i))
        pref, uref = flow.context.convert_to_tensor(pref), flow.context.convert_to_tensor(uref)
        rho, u = native.moments(flow.stencil, flow.f)
        p = flow.units.convert_density_lu_to_pressure_pu(rho)
        u = flow.units.convert_velocity_to_pu(u)
        norm = float(p.numel()) ** 0.5        # resolution^(d/2) with resolution = (#nodes)^(1/d)
        err_u = (torch.norm(u - uref) / norm).item()
        err_p = (torch.norm(p - pref) / norm).item()
        if isinstance(self.out, list):
            self.out.append([err_u, err_p])
        else:
            print(err_u, err_p, file=self.out)


class ProgressReporter(Reporter):
    """Logs wall time, time per step and the estimated time of completion every `interval` steps to
    `<outdir>/progress_reporter_log.txt` and optionally to stdout
    (lettuce/ext/_reporter/progress_reporter.py:12-137).  Reads no device data, so it never synchronises."""
    batchable = True

    def __init__(self, interval=1000, t_max=0, i_target=0, i_start=0, outdir=None, print_message=False,
                 checkpoint=False):
        super().__init__(interval)
        self.t_max, self.i_start, self.i_target = t_max, i_start, i_target
        self.outdir = None if outdir is None else str(outdir)
        self.print_message, self.checkpoint = print_message, checkpoint
        if self.outdir is not None:
            import os
            os.makedirs(self.outdir, exist_ok=True)
        self.running = False
        self.t_start = 0.0
        self.t_elapsed = 0.0

    def _log(self, line: str):
        if self.outdir is not None:
            import os
            with open(os.path.join(self.outdir, "progress_reporter_log.txt"), "a") as fh:
                fh.write(line + "\n")
        if self.print_message:
            print(line)

    def start_timer(self):
        from timeit import default_timer as timer
        self.running = True
        self.t_start = timer()
        self.t_elapsed = 0.0
        self._log(f"t_start: {self.t_start}, interval: {self.interval}, i_target: {self.i_target}")
        self._log("|".join(["timestamp ".center(13), "step".center(10), "t_now".center(10), "t_elapsed".center(10),
                            "t_per_step".center(10), "t_remain(est)".center(15), "t_total(est)".center(15),
                            "DATE_FINISH(est)".center(20), " T WARNING"]))

    def __call__(self, simulation):
        import datetime
        from timeit import default_timer as timer
        if not self.running:
            self.start_timer()
            return
        i = simulation.flow.i
        if i % self.interval != 0:
            return
        now = datetime.datetime.now()
        t_now = timer()
        self.t_elapsed = t_now - self.t_start
        t_per_step = self.t_elapsed / (self.interval if i == self.i_start else (i - self.i_start))
        t_remaining = t_per_step * (self.i_target - i)
        t_total = self.t_elapsed + t_remaining
        finish = now + datetime.timedelta(seconds=t_remaining)
        line = " ".join([now.strftime("%y%m%d_%H%M%S").ljust(13), str(i).rjust(10), f"{t_now:.2f}".rjust(10),
                         f"{self.t_elapsed:.2f}".rjust(10), f"{t_per_step:.6f}".rjust(10),
                         f"{t_remaining:.2f}".rjust(15), f"{t_total:.2f}".rjust(15),
                         " " + finish.strftime("%Y-%m-%d %H:%M:%S").ljust(20)])
        if t_total > self.t_max:
            line += f" WARNING t_total>t_max={self.t_max}"
        self._log(line)
        if self.checkpoint and self.t_elapsed > self.t_max and self.outdir is not None:
            import os
            simulation.flow.dump(os.path.join(self.outdir, f"{now.strftime('%y%m%d_%H%M%S')}_f_{i}.cpt"))


def write_image(filename, array2d):
    """2-D field as an image without axes (lettuce/ext/_reporter/write_image.py:4-13); needs matplotlib"""
    from matplotlib import pyplot as plt
    fig, ax = plt.subplots()
    plt.tight_layout()
    ax.imshow(array2d)
    ax.get_xaxis().set_visible(False)
    ax.get_yaxis().set_visible(False)
    plt.savefig(filename)
    plt.close(fig)
