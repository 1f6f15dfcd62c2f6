"""Device engine: owns a vgh_detector handle (conv network plan + CUDA graph) for one (batch, size)."""
import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, arch
from .flame import FLAMELayer


class _DevView:
    """Zero-copy torch view of library-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class Engine:
    def __init__(self, weights: Dict[str, torch.Tensor], batch: int, image_size: int = 640, keep_top_k: int = 100,
                 flame: Optional[FLAMELayer] = None, sparse_heads: Optional[bool] = None, parity: bool = False,
                 act_dtype: Optional[str] = None):
        """`sparse_heads`: run the FLAME branch of the three head levels on 8x8 windows around the NMS survivors
        instead of the whole feature maps (same values at the survivors; the dense [B,A,413] tensor of the
        reference's model boundary then does not exist).  Default: $VGGHEADS_B200_SPARSE_HEADS == "1".
        `parity`: fp32-class conv arithmetic (activations and weights as three bf16 terms, six partial products per MAC
        accumulated in fp32 on the same tensor-core kernels; ~6x the work) for end-to-end checks against the reference's
        fp32 path.  Dense heads only.
        `act_dtype`: "fp16" | "bf16" storage of activations and weights in the throughput mode (arch.ACT_DTYPES; default
        $VGGHEADS_B200_ACT or "fp16").  The parity mode always stores bf16 terms."""
        if not torch.cuda.is_available():
            raise RuntimeError("head_detector_b200.Engine needs a CUDA device (no CPU fallback)")
        self.B, self.S, self.keep_k = int(batch), int(image_size), int(keep_top_k)
        self.flame = flame if flame is not None else FLAMELayer()
        if sparse_heads is None:
            import os

            sparse_heads = os.environ.get("VGGHEADS_B200_SPARSE_HEADS", "0") == "1"
        self.parity = bool(parity)
        self.sparse_heads = bool(sparse_heads) and not self.parity
        self.act_dtype = "bf16" if self.parity else (act_dtype or arch.default_act_dtype())
        self._act_torch = arch.ACT_DTYPES[self.act_dtype]
        self.plan = arch.build_plan(self.S, sparse_heads=(self.B, self.keep_k) if self.sparse_heads else None, fused_stem=not self.parity,
                                    act_dtype=self.act_dtype)
        if self.parity:
            self.plan = arch.split_plan(self.plan)
        self.packed = arch.pack(self.plan, weights)
        pk = self.packed
        bufs = (_lib.BufDesc * len(self.plan.bufs))(*[_lib.BufDesc(*b, 1 if i in self.plan.stack_bufs else 0)
                                                       for i, b in enumerate(self.plan.bufs)])
        ops = (_lib.OpDesc * len(self.plan.ops))()
        for o, op, m in zip(ops, self.plan.ops, pk.op_meta):
            o.kind = op.kind
            o.in_buf, o.in_coff, o.cin = op.src
            o.out_buf, o.out_coff = op.dst
            o.cout, o.ksize, o.stride, o.relu, o.up, o.up_cout = op.cout, op.k, op.stride, op.relu, op.up, op.up_cout
            o.res_buf, o.res_coff = (op.res[0], op.res[1]) if op.res is not None else (-1, 0)
            o.lane = op.lane
            o.level = op.level
            if m:
                o.res_alpha, o.n_pad, o.k_total, o.block_n = m["alpha"], m["n_pad"], m["k_total"], m["block_n"]
                o.w_off, o.b_off = m["w_off"], m["b_off"]
        nd = _lib.NetDesc()
        nd.batch, nd.image_size, nd.n_bufs, nd.n_ops = self.B, self.S, len(self.plan.bufs), len(self.plan.ops)
        nd.bufs, nd.ops = bufs, ops
        nd.weights_host, nd.n_weights = pk.weights.ctypes.data, pk.weights.size
        nd.bias_host, nd.n_bias = pk.bias.ctypes.data, pk.bias.size
        nd.reg_buf = (C.c_int32 * 3)(*self.plan.reg_buf)
        nd.flame_buf = (C.c_int32 * 3)(*self.plan.flame_buf)
        nd.keep_k = self.keep_k
        nd.n_dense_ops = self.plan.n_dense_ops if self.plan.n_dense_ops is not None else len(self.plan.ops)
        nd.split = 1 if self.parity else 0
        nd.act_f16 = 1 if self.act_dtype == "fp16" else 0
        h = C.c_void_p()
        _lib.check(_lib.lib().vgh_detector_create(C.byref(nd), self.flame.handle(), C.byref(h)), "vgh_detector_create")
        self._h = h
        self.A = _lib.lib().vgh_detector_num_anchors(h)
        self._override = None

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib._lib is not None:
            _lib.lib().vgh_detector_destroy(self._h)
            self._h = None

    # -- zero-copy views of the library's result buffers
    def view(self, which, shape, typestr="<f4"):
        return torch.as_tensor(_DevView(_lib.lib().vgh_detector_output(self._h, which), shape, typestr), device="cuda")

    @property
    def input(self):
        return self.view(_lib.OUT_INPUT, (self.B, self.S, self.S, 3), "|u1")

    @property
    def boxes(self):
        return self.view(_lib.OUT_BOXES, (self.B, self.A, 4))

    @property
    def scores(self):
        return self.view(_lib.OUT_SCORES, (self.B, self.A))

    @property
    def keep_cnt(self):
        return self.view(_lib.OUT_KEEP_CNT, (self.B,), "<i4")

    @property
    def keep_idx(self):
        return self.view(_lib.OUT_KEEP_IDX, (self.B, self.keep_k), "<i4")

    @property
    def keep_boxes(self):
        return self.view(_lib.OUT_KEEP_BOXES, (self.B, self.keep_k, 4))

    @property
    def keep_scores(self):
        return self.view(_lib.OUT_KEEP_SCORES, (self.B, self.keep_k))

    @property
    def head_offsets(self):
        return self.view(_lib.OUT_HEAD_OFFSETS, (self.B + 1,), "<i4")

    def head_params(self, n):
        return self.view(_lib.OUT_HEAD_PARAMS, (n, _lib.NUM_PARAMS))

    def head_verts(self, n):
        return self.view(_lib.OUT_HEAD_VERTS, (n, _lib.NUM_VERTS, 3))

    def head_rot(self, n):
        return self.view(_lib.OUT_HEAD_ROT, (n, 3, 3))

    # -- stage calls (the reference seams)
    def forward(self, images_u8: torch.Tensor):
        """uint8 [B,S,S,3] cuda -> (boxes [B,A,4], scores [B,A]) views; replaces `self.model(image)`."""
        assert images_u8.dtype == torch.uint8 and tuple(images_u8.shape) == (self.B, self.S, self.S, 3) and images_u8.is_cuda
        img = images_u8.contiguous()
        _lib.check(_lib.lib().vgh_detector_forward(self._h, img.data_ptr(), _lib.stream_ptr()), "vgh_detector_forward")
        return self.boxes, self.scores

    def dense_flame(self):
        out = torch.empty(self.B, self.A, _lib.NUM_PARAMS, device="cuda")
        _lib.check(_lib.lib().vgh_detector_dense_flame(self._h, out.data_ptr(), _lib.stream_ptr()), "dense_flame")
        return out

    def postprocess(self, conf=0.5, iou=0.5, top_k=1000, img_xform: Optional[torch.Tensor] = None):
        xf = None if img_xform is None else img_xform.to(device="cuda", dtype=torch.float32).contiguous()
        _lib.check(_lib.lib().vgh_detector_postprocess(self._h, conf, iou, top_k, None if xf is None else xf.data_ptr(),
                                                       _lib.stream_ptr()), "vgh_detector_postprocess")

    def set_override(self, boxes: Optional[torch.Tensor], scores: Optional[torch.Tensor]):
        self._override = None if boxes is None else (boxes.float().contiguous(), scores.float().contiguous())
        b, s = (None, None) if self._override is None else (self._override[0].data_ptr(), self._override[1].data_ptr())
        _lib.check(_lib.lib().vgh_detector_set_override(self._h, b, s), "set_override")

    def run_device(self, conf=0.5, iou=0.5, top_k=1000):
        _lib.check(_lib.lib().vgh_detector_run_device(self._h, conf, iou, top_k, _lib.stream_ptr()), "run_device")

    def run_host(self, images_host: torch.Tensor, out: dict, conf=0.5, iou=0.5, top_k=1000, img_xform_host=None):
        """Host-buffer end-to-end call; `out` holds pinned host tensors (see alloc_host_outputs)."""
        total = out["total"]
        _lib.check(_lib.lib().vgh_detector_run_host(
            self._h, images_host.data_ptr(), None if img_xform_host is None else img_xform_host.data_ptr(), conf, iou, top_k,
            out["keep_cnt"].data_ptr(), out["keep_boxes"].data_ptr(), out["keep_scores"].data_ptr(), out["params"].data_ptr(),
            out["verts"].data_ptr(), out["verts"].shape[0], total.data_ptr(), _lib.stream_ptr()), "run_host")
        return int(total[0])

    # -- multi-GPU: packed result records (include/vggheads_b200.h "multi-GPU gather")
    def record_layout(self) -> dict:
        out = (C.c_int64 * 4)()
        _lib.check(_lib.lib().vgh_detector_record_layout(self._h, out), "record_layout")
        return {"B": self.B, "K": self.keep_k, "fixed_words": out[0], "param_words": out[1], "vert_words": out[2], "capacity_words": out[3]}

    def arm_push(self, dst_ptr: int, wait_flag: int = 0, wait_val: int = 0, done_flag: int = 0, done_val: int = 0):
        """The next submit_device / submit_host packs its result record into `dst_ptr` (local or peer-mapped device
        memory) after waiting on the device for *wait_flag >= wait_val, then sets *done_flag = done_val."""
        _lib.check(_lib.lib().vgh_detector_arm_push(self._h, C.c_void_p(dst_ptr), C.c_void_p(wait_flag or None), wait_val,
                                                    C.c_void_p(done_flag or None), done_val), "arm_push")

    def submit_device(self, conf=0.5, iou=0.5, top_k=1000):
        """Graph replay over the staging input on the current stream (+ the armed record push)."""
        _lib.check(_lib.lib().vgh_detector_submit_device(self._h, conf, iou, top_k, _lib.stream_ptr()), "submit_device")

    def push_status(self) -> int:
        st = C.c_int32(0)
        _lib.check(_lib.lib().vgh_detector_push_status(self._h, C.byref(st)), "push_status")
        return st.value

    def submit_host(self, images_host: torch.Tensor, conf=0.5, iou=0.5, top_k=1000):
        """Pipelined end-to-end: enqueue upload + compute + result staging of one batch (max 2 in flight)."""
        _lib.check(_lib.lib().vgh_detector_submit_host(self._h, images_host.data_ptr(), conf, iou, top_k), "submit_host")

    def collect_host(self, out: dict):
        """Download the oldest outstanding batch into `out` (see alloc_host_outputs); returns total heads."""
        _lib.check(_lib.lib().vgh_detector_collect_host(
            self._h, out["keep_cnt"].data_ptr(), out["keep_boxes"].data_ptr(), out["keep_scores"].data_ptr(), out["params"].data_ptr(),
            out["verts"].data_ptr(), out["verts"].shape[0], out["total"].data_ptr()), "collect_host")
        return int(out["total"][0])

    def alloc_host_outputs(self, max_heads: int):
        pin = dict(pin_memory=True)
        return {
            "keep_cnt": torch.zeros(self.B, dtype=torch.int32, **pin),
            "keep_boxes": torch.zeros(self.B, self.keep_k, 4, **pin),
            "keep_scores": torch.zeros(self.B, self.keep_k, **pin),
            "params": torch.zeros(max_heads, _lib.NUM_PARAMS, **pin),
            "verts": torch.zeros(max_heads, _lib.NUM_VERTS, 3, **pin),
            "total": torch.zeros(1, dtype=torch.int32, **pin),
        }

    def read_buffer(self, name: str) -> torch.Tensor:
        """Activation buffer by plan name -> float32 NHWC cpu tensor (debug / layer-wise parity)."""
        i = self.plan.buf_names[name]
        h, w, c, fp32 = self.plan.bufs[i]
        nb = 1 if i in self.plan.stack_bufs else self.B
        arr = np.empty(nb * h * w * c, dtype=np.float32 if fp32 else np.uint16)
        _lib.check(_lib.lib().vgh_detector_read_buffer(self._h, i, arr.ctypes.data, arr.nbytes), "read_buffer")
        t = torch.from_numpy(arr) if fp32 else torch.from_numpy(arr.view(np.int16)).view(self._act_torch).float()
        t = t.reshape(nb, h, w, c)
        if self.parity and not fp32:   # six planes [h|m|h|m|h|l] per 32-channel granule -> y = (l + m) + h
            g = t.reshape(nb, h, w, c // 192, 6, 32)
            t = ((g[..., 5, :] + g[..., 1, :]) + g[..., 0, :]).reshape(nb, h, w, c // 6)
        return t

    def write_buffer(self, name: str, value: torch.Tensor):
        """float32 NHWC cpu tensor -> activation buffer `name` (rounded to the activation dtype unless the buffer is fp32); parity tests."""
        i = self.plan.buf_names[name]
        h, w, c, fp32 = self.plan.bufs[i]
        nb = 1 if i in self.plan.stack_bufs else self.B
        t = value.detach().cpu().contiguous().float()
        if self.parity and not fp32:
            hh, mm, ll = arch.split_terms(t.reshape(nb, h, w, c // 192, 1, 32))
            t = torch.cat([hh, mm, hh, mm, hh, ll], dim=4).reshape(nb, h, w, c)
        assert tuple(t.shape) == (nb, h, w, c), (name, tuple(t.shape), (nb, h, w, c))
        arr = t.numpy() if fp32 else t.to(self._act_torch).view(torch.int16).numpy()
        _lib.check(_lib.lib().vgh_detector_write_buffer(self._h, i, arr.ctypes.data, arr.nbytes), "write_buffer")

    def forward_from(self, first_label: Optional[str]):
        """Run the dense plan from the op labelled `first_label` (None: box decode only) over the current buffers."""
        n_dense = self.plan.n_dense_ops if self.plan.n_dense_ops is not None else len(self.plan.ops)
        first = n_dense if first_label is None else next(i for i, op in enumerate(self.plan.ops) if op.label == first_label)
        _lib.check(_lib.lib().vgh_detector_forward_from(self._h, first, _lib.stream_ptr()), "forward_from")
        return self.boxes, self.scores

    def autotune(self, iters=3):
        _lib.check(_lib.lib().vgh_detector_autotune(self._h, iters, _lib.stream_ptr()), "autotune")

    def op_config(self, i):
        out = (C.c_int32 * 6)()
        _lib.check(_lib.lib().vgh_detector_op_config(self._h, i, out), "op_config")
        return dict(zip(("mt", "stages", "block_n", "bk", "tw", "th"), out))

    def profile(self, iters=5, conf=0.5, iou=0.5, top_k=1000):
        """Per-op device times (ms) over the staging input: list of (label, ms, flops) + post stages."""
        n = len(self.plan.ops) + 4
        ms = np.zeros(n, dtype=np.float32)
        _lib.check(_lib.lib().vgh_detector_profile(self._h, iters, conf, iou, top_k, ms.ctypes.data, n, _lib.stream_ptr()), "profile")
        rows = []
        for op, t in zip(self.plan.ops, ms):
            flops = 0
            if op.kind == _lib.OP_CONV and not op.level:   # patch-level convs: work depends on the survivors, not counted
                src_res = self.plan.bufs[op.src[0]][0]
                out_res = src_res if op.up else src_res // op.stride
                for p in op.parts:
                    cin = sum(s[2] for s in p.segs)
                    flops += 2 * cin * p.cout * out_res * out_res * (1 if p.transposed else op.k * op.k) * self.B
            rows.append((op.label, float(t), flops))
        for label, t in zip(("box_decode", "select_nms", "gather", "flame_decode"), ms[len(self.plan.ops):]):
            rows.append((label, float(t), 0))
        return rows

    @property
    def launch_count(self):
        return _lib.lib().vgh_detector_launch_count(self._h)
