"""HDF5 checkpoint series (API of lettuce/util/datautils.py:17-80).  `h5py` is imported when the reporter is built:
it is an optional dependency (absent from the image this package is developed in).

Every due step appends `flow.f` to the extendable dataset "f" of `<filebase>.h5`; the populations come off the
device through one pinned host buffer that is reused for every snapshot."""
from __future__ import annotations

import io
import pickle

import numpy as np
import torch

from .._simulation import Reporter

__all__ = ["HDF5Reporter"]


class HDF5Reporter(Reporter):
    batchable = True

    def __init__(self, flow, collision, interval, filebase="./output", metadata=None):
        import h5py
        super().__init__(interval)
        from .. import __version__
        self._h5py = h5py
        self.context = flow.context
        self.filebase = filebase
        self.shape = (flow.stencil.q, *[int(n) for n in flow.resolution])
        self._host = None
        with h5py.File(self.filebase + ".h5", "w") as fs:
            fs.attrs["lettuce_version"] = __version__
            fs.attrs["flow"] = self._pickle_to_h5(self._describe(flow))
            fs.attrs["_collision"] = self._pickle_to_h5(self._describe(collision))
            for key, value in (metadata or {}).items():
                fs.attrs[key] = value
            fs.create_dataset(name="f", shape=(0, *self.shape), maxshape=(None, *self.shape),
                              dtype=np.float64 if flow.f.dtype == torch.float64 else np.float32)

    @staticmethod
    def _describe(obj):
        """what is stored about the flow / collision: class name and plain parameters (the reference pickles the
        objects themselves, device tensors included; a restart needs the class and its scalars)"""
        plain = {k: v for k, v in vars(obj).items() if isinstance(v, (int, float, str, bool, list, tuple, type(None)))}
        return dict(cls=type(obj).__name__, parameters=plain)

    def __call__(self, simulation):
        flow = simulation.flow
        if flow.i % self.interval != 0:
            return
        f = flow.f
        if f.is_cuda:
            if self._host is None:
                self._host = torch.empty(f.shape, dtype=f.dtype, pin_memory=True)
            self._host.copy_(f, non_blocking=True)
            torch.cuda.current_stream(f.device).synchronize()
            data = self._host.numpy()
        else:
            data = f.detach().numpy()
        with self._h5py.File(self.filebase + ".h5", "r+") as fs:
            fs["f"].resize(fs["f"].shape[0] + 1, axis=0)
            fs["f"][-1, ...] = data
            fs.attrs["data"] = str(fs["f"].shape[0])
            fs.attrs["steps"] = str(flow.i)

    @staticmethod
    def _pickle_to_h5(instance):
        buffer = io.BytesIO()
        pickle.dump(instance, buffer)
        return np.void(buffer.getvalue())
