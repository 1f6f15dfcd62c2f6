"""Stand-in for the reference's Cython extension `Sim3DR_Cython` (head_detector/Sim3DR/lib/rasterize.pyx, built by the
reference's setup.py:50-61) so that `import head_detector` works here without running the reference's build.
TEST INFRASTRUCTURE ONLY.  `rasterize` forwards to the SAME C++ (`_rasterize`, rasterize_kernel.cpp:219-293) compiled in
place into oracle/_ref/libsim3dr_ref.so (oracle/Makefile), with the argument list of rasterize.pyx:90-102 - so the
reference's own `PNCCProcessor` / `PredictionResult.get_pncc()` run unmodified for golden vectors."""
import ctypes as C
import os

_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref", "libsim3dr_ref.so")


def rasterize(image, vertices, triangles, colors, depth_buffer, ntri, h, w, c, alpha=1, reverse=False):
    if not os.path.exists(_LIB):
        raise RuntimeError("oracle/_ref/libsim3dr_ref.so is missing: run `make -C oracle _ref/libsim3dr_ref.so` (needs /root/reference)")
    fn = getattr(C.CDLL(_LIB), "_Z10_rasterizePhPfPiS0_S0_iiiifb")
    fn.restype = None
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_float, C.c_bool]
    for a in (image, vertices, triangles, colors, depth_buffer):
        assert a.flags.c_contiguous
    fn(image.ctypes.data, vertices.ctypes.data, triangles.ctypes.data, colors.ctypes.data, depth_buffer.ctypes.data, ntri, h, w, c, float(alpha), bool(reverse))


def get_normal(*a, **k):
    raise RuntimeError("Sim3DR_Cython stand-in: get_normal is not on any path this repo exercises")
