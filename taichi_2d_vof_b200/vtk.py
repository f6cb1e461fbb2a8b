"""Minimal VTK XML writer for the 3-D driver's export (3dvof.py:624-627).

The reference calls ``pyevtk.hl.gridToVTK(path, x, y, z, pointData={"VOF": F})`` which writes ``path.vtr``: a VTK
RectilinearGrid with the three coordinate arrays and one value per grid point, appended as raw binary.  pyevtk is a
third-party dependency that is not needed for that: this module writes the same kind of file with NumPy only (host I/O,
no computation on the fields)."""
from __future__ import annotations

import struct

import numpy as np

_VTK_TYPE = {np.dtype(np.float32): "Float32", np.dtype(np.float64): "Float64", np.dtype(np.int32): "Int32"}


def grid_to_vtk(path: str, x, y, z, pointData: dict) -> str:
    """Write ``path + '.vtr'`` (same argument meaning as pyevtk's gridToVTK) and return the file name.
    x, y, z: 1-D coordinate arrays of lengths (n1, n2, n3); every pointData array has shape (n1, n2, n3)."""
    coords = [np.ascontiguousarray(a) for a in (x, y, z)]
    n = tuple(len(c) for c in coords)
    blocks, offset = [], 0

    def add(arr):
        nonlocal offset
        raw = np.ascontiguousarray(arr).tobytes()
        off = offset
        blocks.append(struct.pack("<Q", len(raw)) + raw)        # UInt64 byte count, then the data (appended, raw)
        offset += 8 + len(raw)
        return off

    lines = ['<?xml version="1.0"?>',
             '<VTKFile type="RectilinearGrid" version="1.0" byte_order="LittleEndian" header_type="UInt64">',
             f'  <RectilinearGrid WholeExtent="0 {n[0] - 1} 0 {n[1] - 1} 0 {n[2] - 1}">',
             f'    <Piece Extent="0 {n[0] - 1} 0 {n[1] - 1} 0 {n[2] - 1}">',
             f'      <PointData Scalars="{next(iter(pointData))}">' if pointData else '      <PointData>']
    for name, a in pointData.items():
        a = np.asarray(a)
        if a.shape != n:
            raise ValueError(f"pointData[{name!r}] has shape {a.shape}, the grid has {n} points")
        # VTK's point index runs fastest along x: Fortran order of an array indexed [i, j, k]
        off = add(np.asfortranarray(a).ravel(order="F"))
        lines.append(f'        <DataArray type="{_VTK_TYPE[a.dtype]}" Name="{name}" NumberOfComponents="1" format="appended" offset="{off}"/>')
    lines += ['      </PointData>', '      <Coordinates>']
    for name, c in zip(("x_coordinates", "y_coordinates", "z_coordinates"), coords):
        off = add(c)
        lines.append(f'        <DataArray type="{_VTK_TYPE[c.dtype]}" Name="{name}" NumberOfComponents="1" format="appended" offset="{off}"/>')
    lines += ['      </Coordinates>', '    </Piece>', '  </RectilinearGrid>', '  <AppendedData encoding="raw">']
    fname = path + ".vtr"
    with open(fname, "wb") as f:
        f.write(("\n".join(lines) + "\n_").encode("ascii"))
        for b in blocks:
            f.write(b)
        f.write(b"\n  </AppendedData>\n</VTKFile>\n")
    return fname


def read_vtr_arrays(fname: str) -> dict:
    """Read back the appended arrays of a file written by grid_to_vtk (tests; not a general VTK reader)."""
    import re
    raw = open(fname, "rb").read()
    head, _, rest = raw.partition(b'<AppendedData encoding="raw">')
    data = rest[rest.index(b"_") + 1:]
    ext = [int(v) for v in re.search(rb'WholeExtent="([^"]+)"', head).group(1).split()]
    n = (ext[1] + 1, ext[3] + 1, ext[5] + 1)
    out = {"shape": n}
    for m in re.finditer(rb'<DataArray type="(\w+)" Name="([^"]+)"[^>]*offset="(\d+)"', head):
        dt = {v: k for k, v in _VTK_TYPE.items()}[m.group(1).decode()]
        off = int(m.group(3))
        nbytes = struct.unpack("<Q", data[off:off + 8])[0]
        a = np.frombuffer(data[off + 8:off + 8 + nbytes], dtype=dt)
        name = m.group(2).decode()
        out[name] = a.reshape(n, order="F") if a.size == n[0] * n[1] * n[2] and not name.endswith("_coordinates") else a
    return out
