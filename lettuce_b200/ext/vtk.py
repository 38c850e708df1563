"""VTK output (API of lettuce/ext/_reporter/vtk_reporter.py:1-69) without pyevtk.

`write_vtk` writes the same file the reference gets from `pyevtk.hl.gridToVTK`: an XML RectilinearGrid (`.vtr`) with
integer node coordinates 0..n-1, point data in appended raw binary blocks (UInt64 byte count + data, x fastest).
`VTKReporter` takes pressure and velocity from the engine's moment kernel (one pass over the populations), copies
them to pinned host memory without blocking the stream, and writes the file on a background thread while the
simulation carries on; `wait()` (also called before the next write and at interpreter exit) joins it.
"""
from __future__ import annotations

import atexit
import os
import struct
from concurrent.futures import ThreadPoolExecutor
from typing import Dict

import numpy as np
import torch

from .. import native
from .._simulation import Reporter

__all__ = ["VTKReporter", "write_vtk", "read_vtr"]

_VTK_TYPE = {"float32": "Float32", "float64": "Float64", "int64": "Int64", "int32": "Int32", "uint8": "UInt8"}


def _as_3d(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim == 2:
        a = a[..., None]
    if a.ndim != 3:
        raise ValueError(f"point data must be [nx, ny(, nz)], got shape {a.shape}")
    return a


def write_vtk(point_dict: Dict[str, np.ndarray], id=0, filename_base="./data/output") -> str:
    """`{filename_base}_{id:08d}.vtr` with every array of `point_dict` as point data (vtk_reporter.py:10-15).
    Returns the path."""
    return _write_vtr(f"{filename_base}_{id:08d}.vtr", point_dict)


def _write_vtr(path: str, point_dict: Dict[str, np.ndarray]) -> str:
    arrays = {k: _as_3d(v) for k, v in point_dict.items()}
    if not arrays:
        raise ValueError("no point data")
    shape = next(iter(arrays.values())).shape
    for k, a in arrays.items():
        if a.shape != shape:
            raise ValueError(f"point data '{k}' has shape {a.shape}, expected {shape}")
        if a.dtype.name not in _VTK_TYPE:
            arrays[k] = a.astype(np.float64 if a.dtype.kind == "f" else np.int64)
    nx, ny, nz = shape
    extent = f"0 {nx - 1} 0 {ny - 1} 0 {nz - 1}"
    coords = [np.arange(0, n, dtype=np.int64) for n in shape]
    blocks, offset = [], 0

    def entry(name, a):
        nonlocal offset
        line = (f'<DataArray Name="{name}" NumberOfComponents="1" type="{_VTK_TYPE[a.dtype.name]}" '
                f'format="appended" offset="{offset}"/>\n')
        blocks.append(a)
        offset += 8 + a.nbytes
        return line

    head = ['<?xml version="1.0"?>\n',
            '<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">\n',
            f'<RectilinearGrid WholeExtent="{extent}">\n', f'<Piece Extent="{extent}">\n',
            f'<PointData scalars="{next(iter(arrays))}">\n']
    # VTK orders points with x fastest: the transposed C-order array
    head += [entry(k, np.ascontiguousarray(a.transpose(2, 1, 0)).astype(a.dtype.newbyteorder("<"), copy=False))
             for k, a in arrays.items()]
    head += ['</PointData>\n', '<CellData>\n', '</CellData>\n', '<Coordinates>\n']
    head += [entry(f"{ax}_coordinates", c) for ax, c in zip("xyz", coords)]
    head += ['</Coordinates>\n', '</Piece>\n', '</RectilinearGrid>\n', '<AppendedData encoding="raw">\n_']
    tmp = str(path) + ".part"
    with open(tmp, "wb") as fh:
        fh.write("".join(head).encode())
        for b in blocks:
            fh.write(struct.pack("<Q", b.nbytes))
            fh.write(b.tobytes())
        fh.write(b'\n</AppendedData>\n</VTKFile>\n')
    os.replace(tmp, path)           # readers never see a half-written file
    return str(path)


def read_vtr(path) -> Dict[str, np.ndarray]:
    """Point data and coordinates of a `.vtr` file written by `write_vtk`, as `[nx, ny, nz]` arrays (used by the
    tests and handy for restart / post-processing scripts)."""
    import re
    with open(path, "rb") as fh:
        raw = fh.read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    header = raw[:marker].decode()
    data = raw[raw.index(b"_", marker) + 1:]
    ext = [int(v) for v in re.search(r'WholeExtent="([^"]+)"', header).group(1).split()]
    shape = (ext[1] + 1, ext[3] + 1, ext[5] + 1)
    inv = {v: k for k, v in _VTK_TYPE.items()}
    out = {}
    for name, typ, off in re.findall(r'<DataArray Name="([^"]+)" NumberOfComponents="1" type="(\w+)" '
                                     r'format="appended" offset="(\d+)"/>', header):
        off = int(off)
        nbytes = struct.unpack("<Q", data[off:off + 8])[0]
        a = np.frombuffer(data[off + 8:off + 8 + nbytes], dtype=np.dtype(inv[typ]).newbyteorder("<"))
        out[name] = a.copy() if name.endswith("_coordinates") else a.reshape(shape[::-1]).transpose(2, 1, 0).copy()
    return out


class VTKReporter(Reporter):
    """Writes pressure `p` and velocity `ux, uy(, uz)` in physical units every `interval` steps to
    `{filename_base}_{step:08d}.vtr` (vtk_reporter.py:18-52).  The fields are produced by `lbm_moments`, copied
    to pinned host memory asynchronously and written by a background thread."""
    batchable = True

    def __init__(self, interval=50, filename_base="./data/output"):
        super().__init__(interval)
        self.filename_base = str(filename_base)
        directory = os.path.dirname(self.filename_base)
        if directory and not os.path.isdir(directory):
            os.makedirs(directory, exist_ok=True)
        self.point_dict = dict()
        self._pool = ThreadPoolExecutor(max_workers=1)
        self._pending = None
        atexit.register(self.wait)

    def flush(self):
        """called by Simulation.__call__ before it returns: the last due file is on disk (the reference writes
        synchronously, vtk_reporter.py) and a failed write raises here"""
        self.wait()

    def wait(self):
        """block until the file of the last due step is on disk (re-raises a failed write)"""
        pending, self._pending = self._pending, None
        if pending is not None:
            pending.result()

    def __call__(self, simulation):
        flow = simulation.flow
        if flow.i % self.interval != 0:
            return
        self.wait()                                   # one file in flight: bounded pinned memory
        rho, u = native.moments(flow.stencil, flow.f)
        p = flow.units.convert_density_lu_to_pressure_pu(rho)
        u = flow.units.convert_velocity_to_pu(u)
        fields = {"p": p[0]}
        for d in range(flow.stencil.d):
            fields[f"u{'xyz'[d]}"] = u[d]
        host, event = {}, None
        for name, t in fields.items():
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=t.is_cuda)
            buf.copy_(t, non_blocking=True)
            host[name] = buf
        if flow.f.is_cuda:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(flow.f.device))
        step, base = int(flow.i), self.filename_base

        def job():
            if event is not None:
                event.synchronize()
            self.point_dict = {k: v.numpy() for k, v in host.items()}
            return write_vtk(self.point_dict, step, base)

        self._pending = self._pool.submit(job)

    def output_mask(self, simulation):
        """`{filename_base}_mask.vtr` with the simulation's no_collision_mask (vtk_reporter.py:54-69)"""
        mask = simulation.no_collision_mask
        if mask is None:
            raise ValueError("the simulation has no boundaries, hence no no_collision_mask")
        return _write_vtr(self.filename_base + "_mask.vtr",
                          {"mask": simulation.flow.context.convert_to_ndarray(mask).astype(np.int64)})
