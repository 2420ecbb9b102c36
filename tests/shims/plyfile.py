"""Import stub for `plyfile` (gsplat/gau_io.py:2); .ply loading is out of scope."""


class PlyData:
    @staticmethod
    def read(path):
        raise ImportError("plyfile is not installed in this image")
