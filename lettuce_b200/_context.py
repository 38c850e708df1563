"""Device / dtype context (API of lettuce/_context.py:9-122).

lettuce_b200 is a CUDA-only engine: a Context may be created on the CPU to build
flows, masks and descriptors (host logic), but stepping or reducing a CPU flow
raises -- there is no CPU code path.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

__all__ = ["Context"]

_FLOATS = (torch.float16, torch.float32, torch.float64)


class Context:
    """`Context(device, dtype, use_native)` with the reference's defaulting rules
    (lettuce/_context.py:17-62): dtype defaults to float32, `use_native` defaults to True on
    CUDA devices and False on the CPU, and requesting the native engine on a CPU device is an
    error."""

    def __init__(self, device: Optional[Union[torch.device, str]] = None,
                 dtype: Optional[torch.dtype] = None, use_native: Optional[bool] = None):
        have_cuda = torch.cuda.is_available()
        if device is None:
            if use_native:
                assert have_cuda, "native engine requested but cuda is not available!"
            device = "cuda:0" if (have_cuda or use_native) else "cpu"
        device = torch.device(device)
        if device.type == "cuda":
            assert have_cuda, "cuda device explicitly requested but cuda is not available!"
            if device.index is None:
                device = torch.device("cuda", torch.cuda.current_device())
            use_native = True if use_native is None else bool(use_native)
        else:
            assert device.type == "cpu", f"lettuce works on cpu or cuda devices; {device} is not supported!"
            assert not use_native, "can not use the native engine on an explicitly requested cpu device!"
            use_native = False
        dtype = dtype or torch.float32
        assert dtype in _FLOATS, f"lettuce works with 16/32/64 bit floats; {dtype} is not supported!"
        self.device = device
        self.dtype = dtype
        self.use_native = use_native

    def synchronize(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    # tensor factories (lettuce/_context.py:72-92)
    def empty_tensor(self, size, *args, dtype=None, **kwargs) -> torch.Tensor:
        return torch.empty(size, *args, **kwargs, device=self.device, dtype=dtype or self.dtype)

    def zero_tensor(self, size, *args, dtype=None, **kwargs) -> torch.Tensor:
        return torch.zeros(size, *args, **kwargs, device=self.device, dtype=dtype or self.dtype)

    def one_tensor(self, size, *args, dtype=None, **kwargs) -> torch.Tensor:
        return torch.ones(size, *args, **kwargs, device=self.device, dtype=dtype or self.dtype)

    def full_tensor(self, size, value, *args, dtype=None, **kwargs) -> torch.Tensor:
        return torch.full(size, value, *args, **kwargs, device=self.device, dtype=dtype or self.dtype)

    def convert_to_tensor(self, array, *args, dtype: Optional[torch.dtype] = None, **kwargs) -> torch.Tensor:
        """Boolean-like inputs become uint8, everything else the context dtype
        (lettuce/_context.py:94-116)."""
        if dtype is None:
            boolish = (bool, torch.bool, torch.uint8, np.uint8, np.dtype("bool"), np.dtype("uint8"))
            dtype = torch.uint8 if getattr(array, "dtype", None) in boolish else self.dtype
        if not isinstance(array, torch.Tensor):
            array = torch.from_numpy(np.ascontiguousarray(np.array(array)))
        return array.to(*args, **kwargs, device=self.device, dtype=dtype)

    @staticmethod
    def convert_to_ndarray(tensor: Union[torch.Tensor, List]) -> np.ndarray:
        if isinstance(tensor, torch.Tensor):
            return tensor.detach().cpu().numpy()
        return np.array(tensor)
