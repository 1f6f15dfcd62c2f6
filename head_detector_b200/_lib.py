"""ctypes binding of libvggheads_b200.so (include/vggheads_b200.h).  There is no fallback: if the
library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvggheads_b200.so")

NUM_VERTS = 5023
NUM_PARAMS = 413

(OUT_BOXES, OUT_SCORES, OUT_KEEP_IDX, OUT_KEEP_CNT, OUT_KEEP_BOXES, OUT_KEEP_SCORES, OUT_HEAD_OFFSETS, OUT_HEAD_PARAMS,
 OUT_HEAD_VERTS, OUT_HEAD_ROT, OUT_INPUT) = range(11)
OP_STEM, OP_CONV, OP_SPP, OP_PATCH_GATHER, OP_PATCH_MASK, OP_STEM_CONV = 0, 1, 2, 3, 4, 5


class BufDesc(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("fp32", C.c_int32), ("stack", C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("in_buf", C.c_int32), ("in_coff", C.c_int32), ("cin", C.c_int32),
        ("out_buf", C.c_int32), ("out_coff", C.c_int32), ("cout", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("relu", C.c_int32), ("up", C.c_int32), ("up_cout", C.c_int32),
        ("res_buf", C.c_int32), ("res_coff", C.c_int32), ("res_alpha", C.c_float),
        ("n_pad", C.c_int32), ("k_total", C.c_int32), ("block_n", C.c_int32),
        ("w_off", C.c_int64), ("b_off", C.c_int64),
        ("lane", C.c_int32), ("level", C.c_int32),
    ]


class NetDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("image_size", C.c_int32), ("n_bufs", C.c_int32), ("n_ops", C.c_int32),
        ("bufs", C.POINTER(BufDesc)), ("ops", C.POINTER(OpDesc)),
        ("weights_host", C.c_void_p), ("n_weights", C.c_int64),
        ("bias_host", C.c_void_p), ("n_bias", C.c_int64),
        ("reg_buf", C.c_int32 * 3), ("flame_buf", C.c_int32 * 3),
        ("keep_k", C.c_int32), ("n_dense_ops", C.c_int32), ("split", C.c_int32), ("act_f16", C.c_int32),
    ]


_SIGS = {
    "vgh_version": (C.c_int, []),
    "vgh_last_error": (C.c_char_p, []),
    "vgh_flame_create": (C.c_int, [C.c_void_p] * 5 + [C.POINTER(C.c_void_p)]),
    "vgh_flame_destroy": (None, [C.c_void_p]),
    "vgh_flame_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "vgh_select_nms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_letterbox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_pncc_render": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_head_bbox": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vgh_detector_create": (C.c_int, [C.POINTER(NetDesc), C.c_void_p, C.POINTER(C.c_void_p)]),
    "vgh_detector_destroy": (None, [C.c_void_p]),
    "vgh_detector_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_detector_postprocess": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "vgh_detector_dense_flame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_detector_output": (C.c_void_p, [C.c_void_p, C.c_int]),
    "vgh_detector_num_anchors": (C.c_int, [C.c_void_p]),
    "vgh_detector_read_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "vgh_detector_write_buffer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "vgh_detector_forward_from": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vgh_detector_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int]),
    "vgh_detector_collect_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_submit_device": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "vgh_detector_record_layout": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgh_detector_arm_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]),
    "vgh_detector_push_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vgh_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "vgh_peer_free": (C.c_int, [C.c_void_p]),
    "vgh_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vgh_peer_close": (C.c_int, [C.c_void_p]),
    "vgh_gather_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_run_device": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]),
    "vgh_detector_set_override": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vgh_detector_profile": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_autotune": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_op_config": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "vgh_detector_launch_count": (C.c_int, [C.c_void_p]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(head_detector_b200 has no CPU or PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().vgh_last_error()
        raise RuntimeError(f"libvggheads_b200 {what} failed (status {rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
