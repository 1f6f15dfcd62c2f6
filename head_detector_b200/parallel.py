"""Batch-sharded multi-GPU inference: one process per GPU, images sharded by rank, no collective on the data path
except the final gather of predictions to rank 0 (BASELINE.json north_star; batched semantics of
yolo_heads_post_prediction_callback.py:55-97: every image independently).  torch.distributed is plumbing only.

Every step's predictions travel as ONE packed record (layout: include/vggheads_b200.h, "multi-GPU gather") that a
single kernel writes from the live result buffers (`Engine.arm_push`).  Two transports:

* `PeerGather` (default on NVLink boxes): rank 0 owns a receive ring that every rank maps through CUDA IPC; the pack
  kernel stores the record straight into it over NVLink and raises a flag; a one-block consumer kernel on rank 0 waits
  for the flags of a step and acknowledges the slots.  No size exchange, no host synchronisation, no NCCL call per step.
* `RecordGather` (NCCL, or gloo in the CPU tests): the record is packed locally; the head counts are all-gathered
  on a side stream and read by the host `lag` steps later (the copy has long finished, so nothing stalls), then one
  exactly-sized send/recv per rank moves the record.
"""
import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

NUM_PARAMS, VERT_WORDS, HEADER = 413, 5023 * 3, 16


def shard_range(total: int, rank: int, world: int):
    """Contiguous slice of the global batch owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ------------------------------------------------------------------------------------------ packed records
def _pad4(n: int) -> int:
    return (n + 3) & ~3


def record_layout(B: int, K: int) -> dict:
    """Python twin of csrc/gather.cu:record_layout (checked against the library in tests/test_gpu_parallel.py)."""
    fixed = HEADER + _pad4(B) + 4 * B * K + _pad4(B * K)
    return {"B": B, "K": K, "fixed_words": fixed, "param_words": NUM_PARAMS, "vert_words": VERT_WORDS,
            "capacity_words": fixed + _pad4(B * K * NUM_PARAMS) + B * K * VERT_WORDS}


def record_words(layout: dict, n_heads: int) -> int:
    """Words of a record that holds n_heads heads (what has to travel)."""
    return layout["fixed_words"] + _pad4(n_heads * NUM_PARAMS) + n_heads * VERT_WORDS


def pack_record(layout: dict, keep_cnt, boxes, scores, params, verts, seq: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Host/torch twin of the pack kernel (tests, gloo path): tensors -> float32 record."""
    B, K = layout["B"], layout["K"]
    n = int(params.shape[0])
    rec = out if out is not None else torch.zeros(record_words(layout, n), dtype=torch.float32, device=params.device)
    rec[:HEADER].view(torch.int32).copy_(torch.tensor([n, B, K, seq] + [0] * (HEADER - 4), dtype=torch.int32))
    at = HEADER
    rec[at:at + B].view(torch.int32).copy_(keep_cnt.to(torch.int32))
    at += _pad4(B)
    rec[at:at + 4 * B * K].copy_(boxes.reshape(-1))
    at += 4 * B * K
    rec[at:at + B * K].copy_(scores.reshape(-1))
    f = layout["fixed_words"]
    rec[f:f + n * NUM_PARAMS].copy_(params.reshape(-1))
    v0 = f + _pad4(n * NUM_PARAMS)
    rec[v0:v0 + n * VERT_WORDS].copy_(verts.reshape(-1))
    return rec


def unpack_record(layout: dict, rec: torch.Tensor, n_heads: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Zero-copy views of one record: keep_cnt [B] int32, boxes [B,K,4], scores [B,K], params [n,413], verts [n,5023,3].
    Reads the head count from the record (one device->host sync) unless `n_heads` is given."""
    B, K = layout["B"], layout["K"]
    n = int(rec[:1].view(torch.int32)[0]) if n_heads is None else n_heads
    at = HEADER
    cnt = rec[at:at + B].view(torch.int32)
    at += _pad4(B)
    boxes = rec[at:at + 4 * B * K].view(B, K, 4)
    at += 4 * B * K
    scores = rec[at:at + B * K].view(B, K)
    f = layout["fixed_words"]
    v0 = f + _pad4(n * NUM_PARAMS)
    return {"n_heads": n, "seq": None if n_heads is not None else int(rec[3:4].view(torch.int32)[0]), "keep_cnt": cnt, "boxes": boxes, "scores": scores,
            "params": rec[f:f + n * NUM_PARAMS].view(n, NUM_PARAMS), "verts": rec[v0:v0 + n * VERT_WORDS].view(n, 5023, 3)}


# ------------------------------------------------------------------------------------------ NCCL / gloo transport
class RecordGather:
    """Gather of packed records to rank `dst` with torch.distributed send/recv, sized exactly, without a host
    synchronisation on the launch path: the head counts of step i are all-gathered on the side stream right away but
    only READ by the host when step i+lag is submitted."""

    def __init__(self, layout: dict, lag: int = 2, dst: int = 0, group=None, device: Optional[torch.device] = None, ring: Optional[int] = None):
        self.layout, self.lag, self.dst, self.group = layout, max(0, lag), dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        n_ring = ring if ring is not None else self.lag + 2
        cap = layout["capacity_words"]
        self.ring = [torch.empty(cap, dtype=torch.float32, device=self.device) for _ in range(n_ring)]
        self.free_ev = [None] * n_ring           # event after which ring[i] may be overwritten (its send has been queued and waited)
        self.counts_dev = [torch.zeros(self.world, dtype=torch.int32, device=self.device) for _ in range(n_ring)]
        self.counts_host = [torch.zeros(self.world, dtype=torch.int32, pin_memory=self.cuda) for _ in range(n_ring)]
        self.recv = [torch.empty(cap, dtype=torch.float32, device=self.device) if r != dst else None for r in range(self.world)] if self.rank == dst else None
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.pending: List[tuple] = []
        self.next = 0
        self.last: Optional[List[torch.Tensor]] = None   # dst: records of the most recently completed gather (views, valid until the next one)
        self.last_counts: Optional[List[int]] = None
        self._done = 0                                   # completed gathers

    def acquire(self):
        """-> (slot id, record tensor to pack into).  The caller's stream must wait for `wait_event(slot)` first."""
        i = self.next
        self.next = (self.next + 1) % len(self.ring)
        return i, self.ring[i]

    def wait_event(self, slot: int):
        return self.free_ev[slot]

    def submit(self, slot: int, ready_event=None):
        """Record `slot` has been queued for packing (ready_event on the packing stream).  Starts the count exchange and
        completes the gather of the step `lag` submissions ago."""
        rec = self.ring[slot]
        if self.cuda:
            with torch.cuda.stream(self.side):
                if ready_event is not None:
                    self.side.wait_event(ready_event)
                dist.all_gather_into_tensor(self.counts_dev[slot], rec[:1].view(torch.int32), group=self.group)
                self.counts_host[slot].copy_(self.counts_dev[slot], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.side)
        else:
            parts = [torch.zeros(1, dtype=torch.int32) for _ in range(self.world)]
            dist.all_gather(parts, rec[:1].view(torch.int32).clone(), group=self.group)
            self.counts_host[slot].copy_(torch.cat(parts))
            ev = None
        self.pending.append((slot, ev))
        while len(self.pending) > self.lag:
            self._finish(*self.pending.pop(0))

    def _finish(self, slot: int, ev):
        if ev is not None:
            ev.synchronize()   # completed `lag` steps ago unless the host runs far ahead of the device
        counts = self.counts_host[slot].tolist()
        rec = self.ring[slot]
        ops = []
        if self.rank == self.dst:
            for r in range(self.world):
                if r != self.dst:
                    ops.append(dist.P2POp(dist.irecv, self.recv[r][:record_words(self.layout, counts[r])], r, group=self.group))
        else:
            ops.append(dist.P2POp(dist.isend, rec[:record_words(self.layout, counts[self.rank])], self.dst, group=self.group))
        if self.cuda:
            with torch.cuda.stream(self.side):
                for req in (dist.batch_isend_irecv(ops) if ops else []):
                    req.wait()
                fe = torch.cuda.Event()
                fe.record(self.side)
                self.free_ev[slot] = fe
        else:
            for req in (dist.batch_isend_irecv(ops) if ops else []):
                req.wait()
        if self.rank == self.dst:
            self.last = [rec if r == self.dst else self.recv[r] for r in range(self.world)]
            self.last_counts = counts
        self._done += 1

    def flush(self, stream=None):
        """Complete every outstanding gather; `stream` (cuda) then waits for the side stream."""
        while self.pending:
            self._finish(*self.pending.pop(0))
        if self.cuda:
            (stream or torch.cuda.current_stream()).wait_stream(self.side)


# ------------------------------------------------------------------------------------------ NVLink peer-memory transport
class _DevView:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerUnavailable(RuntimeError):
    """CUDA IPC peer mapping could not be set up on this box (raised on every rank together)."""


class PeerGather:
    """Records are stored by the pack kernel straight into rank 0's receive ring over NVLink (CUDA IPC mapping);
    flow control by device-side flags.  Ring: `depth` slots per rank; step t of a rank uses slot t % depth."""

    def __init__(self, layout: dict, depth: int = 4, group=None, timeout_ms: int = 4000):
        from . import _lib

        self.lib, self.layout, self.depth, self.group, self.timeout_ms = _lib.lib(), layout, depth, group, timeout_ms
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.words = (layout["capacity_words"] + 63) & ~63
        self._opened: List[int] = []
        self._owned: List[int] = []
        # every rank runs every collective of the set-up even when a local step failed, so that a box without working
        # CUDA IPC makes ALL ranks raise PeerUnavailable together (the caller then falls back to NCCL send/recv)
        err = None
        h_ring = h_ready = h_ack = None
        try:
            if self.rank == 0:
                self.ring_ptr, h_ring = self._alloc(self.world * depth * self.words * 4)
                self.ready_ptr, h_ready = self._alloc(self.world * depth * 8)
            self.ack_ptr, h_ack = self._alloc(max(depth * 8, 256))
        except RuntimeError as ex:
            err = str(ex)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (h_ring, h_ready, h_ack, err), group=group)
        err = next((f"rank {r}: {g[3]}" for r, g in enumerate(gathered) if g[3]), None)
        if err is None:
            try:
                if self.rank != 0:
                    self.ring_ptr = self._open(gathered[0][0])
                    self.ready_ptr = self._open(gathered[0][1])
                else:
                    ack_ptrs = [self.ack_ptr if r == 0 else self._open(gathered[r][2]) for r in range(self.world)]
                    table = torch.tensor([[ack_ptrs[r] + 8 * s for r in range(self.world)] for s in range(depth)], dtype=torch.int64)
                    self.ack_table = table.cuda()                      # [depth][world] device pointers
                    self.totals = torch.zeros(depth, dtype=torch.int32, device="cuda")
                    self.status = torch.zeros(1, dtype=torch.int32, device="cuda")
                    self.side = torch.cuda.Stream(priority=-1)
            except RuntimeError as ex:
                err = str(ex)
        errs = [None] * self.world
        dist.all_gather_object(errs, err, group=group)
        err = next((e if str(e).startswith("rank ") else f"rank {r}: {e}" for r, e in enumerate(errs) if e), None)
        if err is not None:
            self._release()
            raise PeerUnavailable(err)

    def _alloc(self, nbytes: int):
        ptr, handle = C.c_void_p(), (C.c_uint8 * 64)()
        from . import _lib

        _lib.check(self.lib.vgh_peer_alloc(nbytes, C.byref(ptr), handle), "vgh_peer_alloc")
        self._owned.append(ptr.value)
        return ptr.value, bytes(handle)

    def _open(self, handle: bytes) -> int:
        from . import _lib

        ptr = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _lib.check(self.lib.vgh_peer_open(buf, C.byref(ptr)), "vgh_peer_open")
        self._opened.append(ptr.value)
        return ptr.value

    def arm(self, eng, t: int):
        """Arm engine `eng` so that its next submit pushes step t (this rank's step counter) to rank 0."""
        s = t % self.depth
        at = self.rank * self.depth + s
        eng.arm_push(self.ring_ptr + at * self.words * 4, self.ack_ptr + 8 * s, t // self.depth, self.ready_ptr + 8 * at, t + 1)

    def consume(self, t: int, ack: bool = True):
        """Rank 0: queue the consumer of step t on the side stream: a one-block kernel waits for every rank's flag and sums
        the head counts into `totals[t % depth]`; with `ack` it then hands the slots back to the producers.  A consumer
        that reads the records (copy-out, D2H ...) passes ack=False, queues its reads on `side`, then calls `ack(t)`.
        No host synchronisation."""
        from . import _lib

        s = t % self.depth
        _lib.check(self.lib.vgh_gather_wait(C.c_void_p(self.ready_ptr + 8 * s), self.world, self.depth, t + 1,
                                            C.c_void_p(self.ring_ptr + s * self.words * 4), self.depth * self.words,
                                            C.c_void_p(self.ack_table[s].data_ptr()) if ack else None, t // self.depth + 1,
                                            C.c_void_p(self.totals[s:].data_ptr()), C.c_void_p(self.status.data_ptr()), self.timeout_ms,
                                            C.c_void_p(self.side.cuda_stream)), "vgh_gather_wait")

    def ack(self, t: int):
        """Rank 0: hand the slots of step t back (after the reads queued on `side`)."""
        self.consume(t, ack=True)

    def record(self, rank: int, t: int) -> torch.Tensor:
        """Rank 0: the record rank `rank` pushed for step t (valid until the slot is reused `depth` steps later)."""
        at = rank * self.depth + t % self.depth
        return torch.as_tensor(_DevView(self.ring_ptr + at * self.words * 4, (self.layout["capacity_words"],), "<f4"), device="cuda")

    def _release(self):
        for p in self._opened:
            self.lib.vgh_peer_close(C.c_void_p(p))
        self._opened = []
        dist.barrier(group=self.group)   # nobody frees memory a peer still has mapped
        for p in self._owned:
            self.lib.vgh_peer_free(C.c_void_p(p))
        self._owned = []

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._release()


# ------------------------------------------------------------------------------------------ dict-level gather (generic)
FIXED_KEYS = ("keep_cnt", "boxes", "scores")  # per-image, fixed-capacity tensors: gathered as they are


def gather_predictions(local: Dict[str, torch.Tensor], dst: int = 0, group=None, n_heads: Optional[int] = None) -> Optional[Dict[str, torch.Tensor]]:
    """Ragged gather of one step's predictions (a dict of tensors) to rank `dst` in two collectives - the generic,
    synchronous form for callers outside the streaming pipeline.  Keys in FIXED_KEYS have the same shape on every rank;
    all other tensors have the number of local heads as first dim (image-major).  Returns the concatenated dict on
    `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = local["keep_cnt"]
    dev = cnt.device
    if n_heads is None:
        n_heads = int(cnt.sum())
    mine = torch.tensor([n_heads], device=dev, dtype=torch.int64)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    heads = torch.cat(sizes).tolist()
    keys = list(local.keys())
    fixed = [k for k in keys if k in FIXED_KEYS]
    ragged = [k for k in keys if k not in FIXED_KEYS]
    row_w = {k: math.prod(local[k].shape[1:]) for k in ragged}
    fixed_len = sum(local[k].numel() for k in fixed)
    max_len = fixed_len + max(heads + [0]) * sum(row_w.values())
    parts = [local[k].reshape(-1).to(torch.float32) for k in fixed] + [local[k][:n_heads].reshape(-1).to(torch.float32) for k in ragged]
    used = fixed_len + n_heads * sum(row_w.values())
    buf = torch.empty(max(max_len, 1), dtype=torch.float32, device=dev)
    if used:
        torch.cat(parts, out=buf[:used])
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out: Dict[str, torch.Tensor] = {}
    pieces: Dict[str, list] = {k: [] for k in keys}
    for r, rb in enumerate(bufs):
        at = 0
        for k in fixed:
            n = local[k].numel()
            pieces[k].append(rb[at:at + n].reshape(local[k].shape))
            at += n
        for k in ragged:
            n = heads[r] * row_w[k]
            pieces[k].append(rb[at:at + n].reshape((heads[r],) + tuple(local[k].shape[1:])))
            at += n
    for k in keys:
        t = torch.cat(pieces[k])
        out[k] = t if local[k].dtype == torch.float32 else t.round().to(local[k].dtype)
    return out
