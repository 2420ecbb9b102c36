"""Minimal stand-in for the `plyfile` package (gsplat/gau_io.py:2), which is not installed in
this image.  Only what the reference's load_ply touches: PlyData.read(path).elements[0][name]
(a column) and .elements[0][0] (first row; its len() is the property count).  Handles
binary_little_endian PLY with scalar properties -- the layout official 3DGS checkpoints use.
Test infrastructure only (tests/golden/make_golden_density.py runs the REFERENCE's loader on
top of it); the product's own reader is easygaussiansplatting_b200/gau_io.py."""
import numpy as np

_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
          "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
          "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


class PlyData:
    def __init__(self, elements):
        self.elements = elements

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, count, props, in_vertex = None, 0, [], False
            while True:
                line = f.readline().decode("ascii").strip()
                if line == "end_header":
                    break
                tok = line.split()
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    in_vertex = tok[1] == "vertex"
                    if in_vertex:
                        count = int(tok[2])
                elif tok[0] == "property" and in_vertex:
                    props.append((tok[2], _TYPES[tok[1]]))
            if fmt != "binary_little_endian":
                raise ValueError("only binary_little_endian PLY is supported")
            data = np.fromfile(f, dtype=np.dtype(props), count=count)
        return PlyData([data])
