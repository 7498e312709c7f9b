"""Minimal reader of the raw-appended VTK XML pieces cracks_b200/host/vtu_writer.cc writes (tests only)."""
import re

import numpy as np

_DTYPES = {"Float64": np.float64, "Float32": np.float32, "Int64": np.int64, "UInt8": np.uint8}


def read_vtu(path):
    raw = open(path, "rb").read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", marker) + 1
    head = raw[:marker].decode()
    piece = re.search(r'<Piece NumberOfPoints="(\d+)" NumberOfCells="(\d+)">', head)
    out = {"n_points": int(piece.group(1)), "n_cells": int(piece.group(2)), "point_data": [], "cell_data": []}
    section = None
    for line in head.splitlines():
        for tag in ("PointData", "CellData", "Points", "Cells"):
            if "<" + tag + ">" in line:
                section = tag
        m = re.search(r'<DataArray type="(\w+)" Name="(\w+)" NumberOfComponents="(\d+)" format="appended" offset="(\d+)"/>', line)
        if not m:
            continue
        dtype, name, ncomp, offset = _DTYPES[m.group(1)], m.group(2), int(m.group(3)), int(m.group(4))
        nbytes = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=start + offset)[0])
        a = np.frombuffer(raw, dtype=dtype, count=nbytes // np.dtype(dtype).itemsize, offset=start + offset + 8)
        out[name] = a.reshape(-1, ncomp) if ncomp > 1 else a
        if section == "PointData":
            out["point_data"].append(name)
        elif section == "CellData":
            out["cell_data"].append(name)
    assert raw[start + offset + 8 + nbytes:].strip().startswith(b"</AppendedData>")
    return out
