"""Minimal readers for the reference's mesh fixtures (TEST INFRASTRUCTURE ONLY; the product keeps the
reference's MeshImport, which stays as it is per BASELINE.json north_star).

  read_vtu   : .vtu with raw appended data, zlib blocks, UInt64 headers (what ReadVTK reads in
               MeshImport.jl:47-98); returns points (n,3) float64 and 0-based cells of the dominant type.
  read_msh41 : Gmsh .msh 4.1 ASCII (what FerriteGmsh reads, MeshImport.jl:28-37); returns points and the
               0-based 4-node tetrahedra (element type 4), nodes numbered in tag order.
"""
import re
import struct
import zlib

import numpy as np

_DT = {"Float64": "<f8", "Float32": "<f4", "Int64": "<i8", "Int32": "<i4", "UInt8": "u1", "UInt64": "<u8"}


def _appended_array(blob, offset, dtype):
    nb, bs, last = struct.unpack_from("<3Q", blob, offset)
    sizes = struct.unpack_from(f"<{nb}Q", blob, offset + 24)
    pos = offset + 24 + 8 * nb
    out = bytearray()
    for s in sizes:
        out += zlib.decompress(blob[pos:pos + s])
        pos += s
    return np.frombuffer(bytes(out), dtype=_DT[dtype])


def read_vtu_arrays(path):
    """Every appended DataArray of the file by name (Points, connectivity, ..., and any cell / point data)."""
    raw = open(path, "rb").read()
    head_end = raw.index(b"<AppendedData")
    head = raw[:head_end].decode("utf-8", "replace")
    assert 'header_type="UInt64"' in head and "vtkZLibDataCompressor" in head
    start = raw.index(b"_", raw.index(b">", head_end)) + 1
    blob = raw[start:]
    arrays = {}
    for m in re.finditer(r'<DataArray type="(\w+)" Name="(\w+)"[^>]*offset="(\d+)"', head):
        arrays[m.group(2)] = _appended_array(blob, int(m.group(3)), m.group(1))
    return arrays


def read_vtu(path):
    arrays = read_vtu_arrays(path)
    pts = arrays["Points"].reshape(-1, 3).astype(np.float64)
    conn, offs, types = arrays["connectivity"].astype(np.int64), arrays["offsets"].astype(np.int64), arrays["types"]
    starts = np.concatenate([[0], offs[:-1]])
    vtk_type = int(np.bincount(types).argmax())  # dominant cell type (MeshImport.jl:100-130)
    npe = {12: 8, 10: 4}[vtk_type]
    sel = np.nonzero(types == vtk_type)[0]
    cells = np.stack([conn[starts[sel] + a] for a in range(npe)], axis=1)
    return pts, cells


def read_msh41(path):
    lines = open(path).read().split("\n")
    i = lines.index("$Nodes") + 1
    nblocks, nnodes = (int(v) for v in lines[i].split()[:2])
    i += 1
    tags, xyz = [], []
    for _ in range(nblocks):
        n = int(lines[i].split()[3])
        i += 1
        tags += [int(t) for t in lines[i:i + n]]
        i += n
        xyz += [[float(v) for v in l.split()] for l in lines[i:i + n]]
        i += n
    order = np.argsort(tags)
    tags = np.asarray(tags)[order]
    pts = np.asarray(xyz, dtype=np.float64)[order]
    i = lines.index("$Elements") + 1
    nblocks = int(lines[i].split()[0])
    i += 1
    tets = []
    for _ in range(nblocks):
        _, _, etype, n = (int(v) for v in lines[i].split())
        i += 1
        if etype == 4:
            tets += [[int(v) for v in l.split()[1:5]] for l in lines[i:i + n]]
        i += n
    cells = np.searchsorted(tags, np.asarray(tets, dtype=np.int64))
    return pts, cells
