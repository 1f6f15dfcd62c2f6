"""ctypes binding of oracle/_ref/libsim3dr_ref.so = the reference's own Sim3DR rasteriser compiled from
/root/reference/head_detector/Sim3DR/lib/rasterize_kernel.cpp (recipe: oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
used to validate oracle/pncc_oracle.py and to generate tests/golden/pncc_ref.npz."""
import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libsim3dr_ref.so")


def available() -> bool:
    return os.path.exists(LIB)


def rasterize(vertices: np.ndarray, triangles: np.ndarray, colors: np.ndarray, bg: np.ndarray) -> np.ndarray:
    """Sim3DR.rasterize (head_detector/Sim3DR/Sim3DR.py:15-41): paints `bg` in place with alpha = 1, depth buffer -1e8."""
    lib = C.CDLL(LIB)
    fn = getattr(lib, "_Z10_rasterizePhPfPiS0_S0_iiiifb")   # _rasterize(uchar*, float*, int*, float*, float*, int, int, int, int, float, bool)
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_float, C.c_bool]
    h, w, c = bg.shape
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    col = np.ascontiguousarray(colors, dtype=np.float32)
    depth = np.zeros((h, w), dtype=np.float32) - 1e8
    assert bg.dtype == np.uint8 and bg.flags.c_contiguous
    fn(bg.ctypes.data, v.ctypes.data, t.ctypes.data, col.ctypes.data, depth.ctypes.data, t.shape[0], h, w, c, 1.0, False)
    return bg
