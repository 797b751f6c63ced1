"""Host-side orchestration of the CUDA forward path: buffer reuse, graph preparation cache and the
launch sequences for GCNRig / TemporalAttn / SkinNet_inner.  All arithmetic happens in
`libmorig_b200.so`; this file only allocates torch tensors and fills descriptors.

Data layout in HBM (fp32, row-major; R = n_frames * N rows, key-frame f owns rows [f*N, (f+1)*N)):

  feat [R, ldf]         = [ x1 (64) | x2 (256) | x3 (512) | feature (F) | pos (3) | 0-pad to 32 ]
                          written in place by the GCU layers, read (strided) by the next ones, so
                          the reference's torch.cat calls (models/rignet.py:62,65) never happen
  ec_k [R, 2(H+Dp)]     = [ tpl: x-branch (H) | pos-branch (Dp) | geo: x-branch | pos-branch ]
                          EdgeConv outputs of GCU k in the order GCUMotion concatenates them
                          (models/basic_modules.py:215-217)
  pq_k [R, 4H]          = [ P_tpl | Q_tpl | P_geo | Q_geo ]   factorised first edge Linear
  pqpos [N, 3*4*Dp]     = same for the pos branches of the three GCUs (key-frame independent)
  xg / gb [n_frames*B, 1024]  per-graph max pool and the per-graph bias derived from it
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .packing import KIND_F16, AttnPack, DenseLayer, EdgeBranch, GCNRigPack, GCUPack, SkinPack

NEG_INF = float("-inf")


class Workspace:
    """Named device buffers reused across calls (stable pointers, no allocator traffic)."""

    def __init__(self):
        self._bufs: Dict[Tuple, torch.Tensor] = {}
        self.tracker = AmaxTracker()

    def get(self, name: str, shape, device, dtype=torch.float32, zero: bool = False) -> torch.Tensor:
        key = (name, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            # drop stale shapes of the same name so the cache does not grow with varying batch sizes
            for k in [k for k in self._bufs if k[0] == name and k[3] == str(device)]:
                del self._bufs[k]
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def clear(self):
        self._bufs.clear()


class AmaxTracker:
    """Range bookkeeping for the fp16-split tensor-core layers (csrc/gemm_tc.cuh): one device float per buffer
    holding max |value| of everything written to it during the current forward.  Kernels of this library raise
    it while they store (`c_amax` / `out_amax` / `dst_amax` of the C-ABI); for tensors that came from outside
    (or were last written by a kernel that does not track) `morig_absmax_f32` computes it on demand.  All slots
    are zeroed by one fill at the start of every forward (`begin`), so results do not depend on history."""

    SLOTS = 256

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None
        self.slots: Dict[int, int] = {}
        self.valid: set = set()

    def begin(self, device) -> None:
        if self.buf is None or self.buf.device != torch.device(device):
            self.buf = torch.zeros(self.SLOTS, dtype=torch.float32, device=device)
            self.slots = {}
        elif len(self.slots) >= self.SLOTS - 8:
            self.slots = {}
        fill(self.buf, 0.0)
        self.valid = set()

    def _slot(self, t: torch.Tensor) -> int:
        key = t.untyped_storage().data_ptr()
        idx = self.slots.get(key)
        if idx is None:
            idx = len(self.slots)
            if idx >= self.SLOTS - 1:
                raise RuntimeError("morig_b200: out of amax slots")
            self.slots[key] = idx
        return idx

    def out_ptr(self, t: torch.Tensor) -> int:
        """slot address for a kernel that is about to write (part of) `t` and tracks what it stores"""
        idx = self._slot(t)
        self.valid.add(idx)
        return self.buf.data_ptr() + 4 * idx

    def invalidate(self, t: torch.Tensor) -> None:
        """`t` was written by a kernel that does not track: its next fp16 consumer recomputes the range"""
        idx = self.slots.get(t.untyped_storage().data_ptr())
        if idx is not None and idx in self.valid:
            self.valid.discard(idx)
            _absmax_reset(self.buf, idx)

    def in_ptr(self, t: torch.Tensor, off: int, ld: int, rows: int, cols: int) -> int:
        """slot address holding max|t| for an fp16 consumer of the [rows, cols] block at element offset `off`"""
        idx = self._slot(t)
        if idx not in self.valid:
            _absmax(t, off, ld, rows, cols, self.buf, idx)
            self.valid.add(idx)
        return self.buf.data_ptr() + 4 * idx


_tracker: Optional[AmaxTracker] = None       # set for the duration of one forward (`forward_scope`)
_scratch: Dict[str, torch.Tensor] = {}       # per-device one-float scratch for calls outside a forward scope


class forward_scope:
    """`with forward_scope(ws, device):` around one top-level forward: activates the workspace's AmaxTracker.
    Nested scopes (a module's forward called from another module's forward) are no-ops."""

    def __init__(self, ws: "Workspace", device):
        self.ws, self.device, self.outer = ws, device, False
        self.guard = None

    def __enter__(self):
        global _tracker
        if _tracker is None:
            self.outer = True
            # every launch helper enqueues on torch's current stream of the CURRENT device: make that the device
            # the tensors live on (PyTorch ops guard implicitly; raw launches have to do it themselves)
            if torch.device(self.device).type == "cuda":
                self.guard = torch.cuda.device(self.device)
                self.guard.__enter__()
            _tracker = self.ws.tracker
            _tracker.begin(self.device)
        return self

    def __exit__(self, *exc):
        global _tracker
        if self.outer:
            _tracker = None
            if self.guard is not None:
                self.guard.__exit__(*exc)
                self.guard = None
        return False


def _absmax_reset(buf: torch.Tensor, idx: int) -> None:
    _lib.check(_lib.load().morig_fill_f32(buf.data_ptr() + 4 * idx, 1, 0.0, _lib.stream_ptr()), "morig_fill_f32")
    if _counter is not None:
        _counter.count += 1


def _absmax(t: torch.Tensor, off: int, ld: int, rows: int, cols: int, buf: torch.Tensor, idx: int) -> None:
    tok = _begin("absmax", 1, 0.0, 4.0 * rows * cols) if (_counter is not None or _timer is not None) else None
    _lib.check(_lib.load().morig_absmax_f32(t.data_ptr() + 4 * off, ld, rows, cols, buf.data_ptr() + 4 * idx,
                                            _lib.stream_ptr()), "morig_absmax_f32")
    _end(tok)


def _amax_in(t: torch.Tensor, off: int, ld: int, rows: int, cols: int) -> int:
    if _tracker is not None:
        return _tracker.in_ptr(t, off, ld, rows, cols)
    key = str(t.device)                              # direct helper call (tests, micro-benchmarks): always recompute
    buf = _scratch.get(key)
    if buf is None:
        buf = _scratch[key] = torch.zeros(1, dtype=torch.float32, device=t.device)
    _absmax_reset(buf, 0)
    _absmax(t, off, ld, rows, cols, buf, 0)
    return buf.data_ptr()


def _amax_out(t: Optional[torch.Tensor]) -> int:
    return _tracker.out_ptr(t) if (_tracker is not None and t is not None) else 0


def _untracked(t: Optional[torch.Tensor]) -> None:
    if _tracker is not None and t is not None:
        _tracker.invalidate(t)


@dataclass
class Graph:
    rowptr: torch.Tensor
    col: torch.Tensor
    tgt: torch.Tensor
    n: int
    e_max: int
    ready: Optional[torch.cuda.Event] = None     # set while the CSR is still being built on a side stream
    scratch: Optional[tuple] = None              # keeps the builder's workspace (and edge-list copy) alive until then

    def join(self) -> None:
        """make the current stream wait for a CSR that is being built on a side stream (no-op afterwards)"""
        if self.ready is not None:
            torch.cuda.current_stream(self.rowptr.device).wait_event(self.ready)
            self.ready = None
            self.scratch = None


_side: Dict[str, list] = {}                      # per device: [streams, next index]


def _side_stream(device) -> torch.cuda.Stream:
    ent = _side.get(str(device))
    if ent is None:
        ent = _side[str(device)] = [[torch.cuda.Stream(device), torch.cuda.Stream(device)], 0]
    ent[1] ^= 1
    return ent[0][ent[1]]


def graph_prep(edge_index: torch.Tensor, n: int, overlap: bool = False) -> Graph:
    """`morig_graph_prep` on a [2, E] int64 CUDA tensor.  With `overlap` the five small kernels run on a side stream
    (the two edge sets of a forward and the first vertex layers are independent); `Graph.join()` -- called by the
    first EdgeConv that needs the CSR -- orders the consumer after them.  Inside a CUDA-graph capture the fork and
    the join become graph edges."""
    lib = _lib.load()
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    ei = _lib.require_cuda(edge_index, "edge_index", torch.int64)
    e = ei.shape[1]
    dev = ei.device
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    col = torch.empty(e + n, dtype=torch.int32, device=dev)
    tgt = torch.empty(e + n, dtype=torch.int32, device=dev)
    ws_bytes = lib.morig_graph_prep_workspace(e, n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    g = Graph(rowptr=rowptr, col=col, tgt=tgt, n=n, e_max=e + n)
    if overlap and not hooks_active():
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)                    # the edge list (possibly a copy queued on `cur`) is ready
        _lib.check(lib.morig_graph_prep(ei.data_ptr(), e, n, rowptr.data_ptr(), col.data_ptr(), tgt.data_ptr(),
                                        ws.data_ptr(), ws_bytes, side.cuda_stream), "morig_graph_prep")
        g.ready = torch.cuda.Event()
        g.ready.record(side)
        g.scratch = (ws, ei)     # `ei` may be a contiguous temporary: it must outlive the side-stream kernels too
        return g
    tok = _begin(f"graph_prep E={e} N={n}", 5, 0.0, 16.0 * e + 8.0 * (e + n)) if (_counter is not None or _timer is not None) else None
    _lib.check(lib.morig_graph_prep(ei.data_ptr(), e, n, rowptr.data_ptr(), col.data_ptr(), tgt.data_ptr(),
                                    ws.data_ptr(), ws_bytes, _lib.stream_ptr()), "morig_graph_prep")
    _end(tok)
    return g


class GraphCache:
    """Keeps the CSR of the most recent edge-index tensors.  Entries are matched by tensor identity
    and in-place version and hold a strong reference, so a recycled allocation can never alias."""

    def __init__(self, slots: int = 4):
        self._slots = slots
        self._entries: list = []   # (tensor, version, n, Graph)

    def get(self, edge_index: torch.Tensor, n: int) -> Graph:
        for ent in self._entries:
            if ent[0] is edge_index and ent[1] == edge_index._version and ent[2] == n:
                return ent[3]
        g = graph_prep(edge_index, n, overlap=True)
        self._entries.insert(0, (edge_index, edge_index._version, n, g))
        del self._entries[self._slots:]
        return g


@dataclass
class BatchInfo:
    batch32: torch.Tensor
    n_graphs: int


class BatchCache:
    def __init__(self):
        self._ent = None

    def get(self, batch: torch.Tensor, data=None) -> BatchInfo:
        ent = self._ent
        if ent is not None and ent[0] is batch and ent[1] == batch._version:
            return ent[2]
        b = _lib.require_cuda(batch, "batch", torch.int64)
        ng = getattr(data, "num_graphs", None) if data is not None else None
        if ng is None and data is not None and torch.is_tensor(getattr(data, "ptr", None)):
            ng = data.ptr.numel() - 1          # PyG `Batch.ptr` [B + 1]
        if ng is None:
            # plain `Data` with a hand-made batch vector (evaluate/joint2rig.py:261-263) and no `num_graphs`: the one
            # place that has to read 8 bytes back, once per batch tensor -- exactly what the reference's
            # `scatter_max(x_4, batch)` does for dim_size (models/rignet.py:63).  Set `data.num_graphs` to avoid it.
            ng = int(b[-1].item()) + 1         # `batch` is sorted (PyG collation): last id = B - 1
        info = BatchInfo(batch32=b.to(torch.int32), n_graphs=int(ng))
        self._ent = (batch, batch._version, info)
        return info


# ---- instrumentation (bench.py): launch counting and per-kernel CUDA-event timing ------------------------

class LaunchCounter:
    """counts kernels launched by this package (a graph_prep call launches 5)"""

    def __init__(self):
        self.count = 0

    def reset(self):
        self.count = 0


class KernelTimer:
    """CUDA events on the launch stream around every helper call; summarised per kernel label with
    the algorithmic FLOPs / bytes of one launch (DESIGN.md 'Measurement')."""

    def __init__(self):
        self.records = []

    def reset(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for label, flops, nbytes, s, e in self.records:
            a = agg.setdefault(label, dict(launches=0, ms=0.0, flops=flops, bytes=nbytes))
            a["launches"] += 1
            a["ms"] += s.elapsed_time(e)
        out = []
        for label, a in agg.items():
            avg = a["ms"] / a["launches"]
            out.append(dict(kernel=label, launches=a["launches"], total_ms=round(a["ms"], 4), avg_ms=round(avg, 5),
                            alg_gflop_per_launch=round(a["flops"] / 1e9, 4), alg_mb_per_launch=round(a["bytes"] / 1e6, 4),
                            tflops=round(a["flops"] / (avg * 1e-3) / 1e12, 3) if avg > 0 else None,
                            gbs=round(a["bytes"] / (avg * 1e-3) / 1e9, 1) if avg > 0 else None))
        out.sort(key=lambda r: -r["total_ms"])
        return out


_counter: Optional[LaunchCounter] = None
_timer: Optional[KernelTimer] = None


def set_hooks(counter: Optional[LaunchCounter], timer: Optional[KernelTimer]) -> None:
    global _counter, _timer
    _counter, _timer = counter, timer


def hooks_active() -> bool:
    return _counter is not None or _timer is not None


def _begin(label: str, kernels: int, flops: float, nbytes: float):
    if _counter is not None:
        _counter.count += kernels
    if _timer is None:
        return None
    s = torch.cuda.Event(enable_timing=True)
    s.record(torch.cuda.current_stream())
    return (label, flops, nbytes, s)


def _end(tok) -> None:
    if tok is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        _timer.records.append(tok + (e,))


FP32_FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # nominal CUDA-core peak at max clock (74.4)


def roofline_report(kstats, peaks, ms_per_step: float, alg_flops_step: float, traffic_bytes=None) -> dict:
    """`roofline` object of the bench line for the dominant kernel (largest share of the step).

    The dominant kernels run on the tcgen05 tensor pipe with split-operand error compensation: every algorithmic
    fp32 FLOP costs three MMA FLOPs (hi*hi + hi*lo + lo*hi).  `kind::f16` (default) issues them at the bf16/fp16
    rate, `kind::tf32` (`MORIG_TC_KIND=tf32`) at half of it.  `peak` is the measured bf16 dense figure of
    MEASURED_PEAKS.json (the only compute peak the driver measures); `fp32_equiv_peak` = peak / 3 (f16) or peak / 6
    (tf32) is what this arithmetic could reach at best, and `frac_of_fp32_equiv_peak` is the utilisation figure."""
    from .packing import tc_kind
    f16 = tc_kind() == KIND_F16
    div = 3 if f16 else 6
    if not kstats:
        return {"bound": "tensor", "achieved": None, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": None, "traffic": None}
    top = kstats[0]
    per_step_ms = sum(k["total_ms"] for k in kstats)
    achieved = top["tflops"]
    peak = peaks["bf16_tflops_sustained"]
    pipe = ("tcgen05 kind::f16 on fp16 hi/lo operand splits (power-of-two scaled), 3 MMAs per algorithmic FLOP, "
            "fp32 accumulate in TMEM") if f16 else \
           "tcgen05 kind::tf32, 3 MMAs per algorithmic FLOP (3xTF32 compensation, fp32 accumulate in TMEM)"
    return {"bound": "tensor", "kernel": top["kernel"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": (achieved / peak if achieved else None), "traffic": traffic_bytes,
            "peak_source": peaks["source"] + " bf16 dense, sustained (kernel timed inside a long step)",
            "avg_launch_ms": top["avg_ms"], "alg_gflop_per_launch": top["alg_gflop_per_launch"],
            "alg_mb_per_launch": top["alg_mb_per_launch"],
            "share_of_kernel_time": round(top["total_ms"] / per_step_ms, 4) if per_step_ms else None,
            "pipe": pipe,
            "tensor_tflops_issued": round(3 * achieved, 1) if achieved else None,
            "frac_issued_of_peak": (round(3 * achieved / peak, 4) if achieved else None),
            # the sustained figure is a 4 s cuBLAS loop that the board power limit holds at ~1330 MHz; a sub-millisecond
            # kernel inside the step runs at the full 1965 MHz, so the burst figure is the other bracket
            "burst_peak": peaks["bf16_tflops"],
            "frac_issued_of_burst_peak": (round(3 * achieved / peaks["bf16_tflops"], 4) if achieved else None),
            "fp32_equiv_peak": round(peak / div, 1),
            "frac_of_fp32_equiv_peak": (round(achieved / (peak / div), 4) if achieved else None),
            "fp32_ffma_nominal_peak": round(FP32_FFMA_PEAK_TFLOPS, 1),
            "hbm_gbs_same_kernel": top["gbs"], "hbm_frac_same_kernel": (round(top["gbs"] / peaks["hbm_gbs"], 4)
                                                                       if top["gbs"] else None)}


# ---- thin launch helpers ---------------------------------------------------------------------------

def dense(layer: DenseLayer, A: torch.Tensor, a_off: int, lda: int, M: int, *, K: Optional[int] = None,
          C: Optional[torch.Tensor] = None, c_off: int = 0, ldc: int = 0,
          pool: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None,
          binfo: Optional[BatchInfo] = None, n_vtx: int = 0) -> None:
    lib = _lib.load()
    kk = layer.K if K is None else K
    if kk > layer.W.shape[0] or lda < kk:
        raise ValueError(f"dense: K={kk} exceeds the packed weight rows ({layer.W.shape[0]}) or the row stride ({lda})")
    d = _lib.DenseDesc()
    d.A = A.data_ptr() + 4 * a_off
    d.lda = lda
    d.W, d.ldw = layer.W.data_ptr(), layer.ldw
    d.bias, d.scale, d.shift = _lib.ptr(layer.bias), _lib.ptr(layer.scale), _lib.ptr(layer.shift)
    d.rowbias, d.ldrb = _lib.ptr(rowbias), (rowbias.shape[1] if rowbias is not None else 0)
    if binfo is not None:
        d.batch, d.n_vtx, d.n_graphs = binfo.batch32.data_ptr(), n_vtx, binfo.n_graphs
    d.C, d.ldc = (C.data_ptr() + 4 * c_off if C is not None else 0), ldc
    d.pool, d.ldpool = _lib.ptr(pool), (pool.shape[1] if pool is not None else 0)
    d.M, d.N, d.K = M, layer.N, (layer.K if K is None else K)
    d.relu = 1 if layer.relu else 0
    if layer.Wtc is not None and d.K == layer.K:
        d.Wtc, d.tc_bn, d.tc_kind, d.tc_w_inv = layer.Wtc.data_ptr(), layer.tc_bn, layer.tc_kind, layer.tc_w_inv
        if layer.tc_kind == KIND_F16:
            d.a_amax = _amax_in(A, a_off, lda, M, d.K)
    d.c_amax = _amax_out(C)
    tok = None
    if _counter is not None or _timer is not None:
        kk = d.K
        tok = _begin(f"dense M={M} N={layer.N} K={kk}" + (" +pool" if pool is not None else ""), 1,
                     2.0 * M * layer.N * kk, 4.0 * (M * kk + kk * layer.N + (M * layer.N if C is not None else 0)))
    _lib.check(lib.morig_dense_fwd(ctypes.byref(d), _lib.stream_ptr()), "morig_dense_fwd")
    _end(tok)


def _edge_desc(d, br: EdgeBranch, pq: torch.Tensor, ldpq: int, p_off: int, q_off: int, g: Graph, n_frames: int,
               out: torch.Tensor, ldo: int, out_off: int, out_repeat: int) -> None:
    d.PQ, d.ldpq, d.p_off, d.q_off = pq.data_ptr(), ldpq, p_off, q_off
    d.rowptr, d.col, d.tgt = g.rowptr.data_ptr(), g.col.data_ptr(), g.tgt.data_ptr()
    d.N, d.E_max, d.n_frames, d.out_repeat = g.n, g.e_max, n_frames, out_repeat
    d.W1, d.ldw = br.W1.data_ptr(), br.W1.shape[1]
    d.b1, d.scale, d.shift = br.b1.data_ptr(), br.scale.data_ptr(), br.shift.data_ptr()
    d.out, d.ldo, d.out_off = out.data_ptr(), ldo, out_off
    d.H = br.H
    d.W1tc, d.tc_kind, d.tc_w_inv = _lib.ptr(br.W1tc), br.tc_kind, br.tc_w_inv
    if (br.W1tc is not None and br.tc_kind == KIND_F16) or br.H <= 32:       # split-fp16 kernels need the operand range
        d.pq_amax = _amax_in(pq, 0, ldpq, g.n * (n_frames if out_repeat == 1 else 1), ldpq)
    d.out_amax = _amax_out(out)


def _edge_label(br: EdgeBranch, g: Graph, n_frames: int, out_repeat: int, count: int = 1):
    H, e = br.H, g.e_max
    return (f"edgeconv H={H} E={e} N={g.n} frames={n_frames}" + (f" x{count}" if count > 1 else ""), 1,
            count * 2.0 * e * H * H * n_frames,
            count * (n_frames * 4.0 * g.n * 2 * H + 8.0 * e + 4.0 * g.n * H * n_frames * out_repeat + 4.0 * (H * H + 3 * H)))


def edgeconv(br: EdgeBranch, pq: torch.Tensor, ldpq: int, p_off: int, q_off: int, g: Graph, n_frames: int,
             out: torch.Tensor, ldo: int, out_off: int, out_repeat: int = 1) -> None:
    lib = _lib.load()
    g.join()
    d = _lib.EdgeDesc()
    _edge_desc(d, br, pq, ldpq, p_off, q_off, g, n_frames, out, ldo, out_off, out_repeat)
    tok = _begin(*_edge_label(br, g, n_frames, out_repeat)) if (_counter is not None or _timer is not None) else None
    _lib.check(lib.morig_edgeconv_fwd(ctypes.byref(d), _lib.stream_ptr()), "morig_edgeconv_fwd")
    _end(tok)


def edgeconv_batch(items, g: Graph, n_frames: int, out_repeat: int = 1) -> None:
    """Up to four narrow (H <= 32, equal width) EdgeConv branches on the same graph in one launch.
    items: (br, pq, ldpq, p_off, q_off, out, ldo, out_off) per branch."""
    lib = _lib.load()
    g.join()
    descs = (_lib.EdgeDesc * len(items))()
    for d, (br, pq, ldpq, p_off, q_off, out, ldo, out_off) in zip(descs, items):
        _edge_desc(d, br, pq, ldpq, p_off, q_off, g, n_frames, out, ldo, out_off, out_repeat)
    tok = None
    if _counter is not None or _timer is not None:
        tok = _begin(*_edge_label(items[0][0], g, n_frames, out_repeat, len(items)))
    _lib.check(lib.morig_edgeconv_fwd_batch(descs, len(items), _lib.stream_ptr()), "morig_edgeconv_fwd_batch")
    _end(tok)


def fill(t: torch.Tensor, value: float) -> None:
    tok = _begin("fill", 1, 0.0, 4.0 * t.numel()) if (_counter is not None or _timer is not None) else None
    _lib.check(_lib.load().morig_fill_f32(t.data_ptr(), t.numel(), value, _lib.stream_ptr()), "morig_fill_f32")
    _end(tok)


def fill_many(tensors, value: float) -> None:
    """one launch for up to 8 buffers"""
    tensors = list(tensors)
    while tensors:
        chunk, tensors = tensors[:8], tensors[8:]
        ptrs = (ctypes.c_void_p * len(chunk))(*[t.data_ptr() for t in chunk])
        sizes = (ctypes.c_int64 * len(chunk))(*[t.numel() for t in chunk])
        tok = _begin("fill", 1, 0.0, 4.0 * sum(t.numel() for t in chunk)) if (_counter is not None or _timer is not None) else None
        _lib.check(_lib.load().morig_fill_many_f32(ptrs, sizes, len(chunk), value, _lib.stream_ptr()), "morig_fill_many_f32")
        _end(tok)


def fill_cut(entries, n: int, n_frames: int, value: float) -> None:
    """`morig_fill_cut_f32`: start values for the rows the edge kernels merge atomically.  entries = (Graph, out, ld, col0,
    ncols); one launch for up to 8 of them"""
    entries = list(entries)
    while entries:
        chunk, entries = entries[:8], entries[8:]
        k = len(chunk)
        rp = (ctypes.c_void_p * k)(*[g.rowptr.data_ptr() for g, *_ in chunk])
        op = (ctypes.c_void_p * k)(*[o.data_ptr() for _, o, *_ in chunk])
        ld = (ctypes.c_int32 * k)(*[e[2] for e in chunk])
        c0 = (ctypes.c_int32 * k)(*[e[3] for e in chunk])
        nc = (ctypes.c_int32 * k)(*[e[4] for e in chunk])
        tok = (_begin("fill", 1, 0.0, 4.0 * n * k) if (_counter is not None or _timer is not None) else None)
        _lib.check(_lib.load().morig_fill_cut_f32(rp, op, ld, c0, nc, k, n, n_frames, value, _lib.stream_ptr()),
                   "morig_fill_cut_f32")
        _end(tok)


def gather_cols(src: torch.Tensor, lds: int, src_off: int, frame_stride: int, cols: Optional[torch.Tensor], c: int,
                n: int, n_frames: int, dst: torch.Tensor, ldd: int, dst_off: int) -> None:
    tok = _begin("gather_cols", 1, 0.0, 8.0 * n * n_frames * c) if (_counter is not None or _timer is not None) else None
    _lib.check(_lib.load().morig_gather_cols(src.data_ptr(), lds, src_off, frame_stride, _lib.ptr(cols), c, n,
                                             n_frames, dst.data_ptr(), ldd, dst_off, _amax_out(dst),
                                             _lib.stream_ptr()), "morig_gather_cols")
    _end(tok)


def row_normalize(x: torch.Tensor, ldx: int, rows: int, c: int, dst2: Optional[torch.Tensor] = None, n: int = 0,
                  n_frames: int = 0) -> None:
    tok = _begin("row_normalize", 1, 0.0, 8.0 * rows * c) if (_counter is not None or _timer is not None) else None
    _lib.check(_lib.load().morig_row_normalize(x.data_ptr(), ldx, rows, c, _lib.ptr(dst2), n, n_frames,
                                               _lib.stream_ptr()), "morig_row_normalize")
    _end(tok)
    _untracked(x)
    _untracked(dst2)


# ---- layer sequences -------------------------------------------------------------------------------

def _gcu_buffers(ws: Workspace, tag: str, gp: GCUPack, R: int, dev):
    H, Dp = gp.H, gp.Dp
    return ws.get(tag + ".pq", (R, 4 * H), dev), ws.get(tag + ".ec", (R, 2 * (H + Dp)), dev)


def run_pos_branches(ws: Workspace, tags, gps, pqpos: torch.Tensor, gt: Graph, gg: Graph, n: int, n_frames: int,
                     also_fill=()) -> None:
    """The pos branches of several GCUs (models/basic_modules.py:194, one per edge set and GCU) only read `pos`, so
    they do not take part in the GCU chain: the EdgeConv outputs of all GCUs are initialised and, per edge set, the
    pos branches of all GCUs run as ONE launch (narrow branches) before the chain starts."""
    dev = pqpos.device
    R = n * n_frames
    ldpp = pqpos.shape[1]
    ecs = [_gcu_buffers(ws, tag, gp, R, dev)[1] for tag, gp in zip(tags, gps)]
    # edge tiles merge the segments they cut with an (exact, ordered-int) atomic max: -inf start values for exactly those
    # vertices (per edge set: the CSR decides), one launch; the pooled feature gets a plain fill
    gt.join()
    gg.join()
    fill_cut([(g, ec, 2 * (gp.H + gp.Dp), s * (gp.H + gp.Dp), gp.H + gp.Dp)
              for gp, ec in zip(gps, ecs) for s, g in enumerate((gt, gg))], n, n_frames, NEG_INF)
    if also_fill:
        fill_many(list(also_fill), NEG_INF)
    for s, g in enumerate((gt, gg)):
        items = []
        for gp, ec in zip(gps, ecs):
            H, Dp = gp.H, gp.Dp
            wec = 2 * (H + Dp)
            bp = gp.pos_geo if s else gp.pos_tpl
            pc = gp.pos_col + 2 * s * Dp
            items.append((bp, pqpos, ldpp, pc, pc + Dp, ec, wec, s * (H + Dp) + H))
        narrow = all(it[0].H == items[0][0].H and it[0].H <= 32 for it in items)
        if narrow and 1 < len(items) <= 4:
            edgeconv_batch(items, g, 1, out_repeat=n_frames)
        else:
            for bp, pq_, ld_, po, qo, ec, wec, off in items:
                if bp.H >= 64:      # wide pos branch (skinning net): key-frame count is 1 there
                    edgeconv(bp, pq_, ld_, po, qo, g, 1, ec, wec, off)
                else:
                    edgeconv(bp, pq_, ld_, po, qo, g, 1, ec, wec, off, out_repeat=n_frames)


def run_gcu(ws: Workspace, tag: str, gp: GCUPack, x: torch.Tensor, x_off: int, ldx: int, k_x: int,
            gt: Graph, gg: Graph, n: int, n_frames: int, out: torch.Tensor, out_off: int, ldo: int) -> None:
    """One GCUMotion (models/basic_modules.py:205-219) after `run_pos_branches`: the x branch of both
    EdgeConvMotion layers, then the vertex mlp, written to out[:, out_off : out_off + gp.out]."""
    R = n * n_frames
    H, Dp = gp.H, gp.Dp
    wec = 2 * (H + Dp)
    pq, ec = _gcu_buffers(ws, tag, gp, R, x.device)
    dense(gp.pq_x, x, x_off, ldx, R, K=k_x, C=pq, ldc=4 * H)
    for s, (g, bx) in enumerate(((gt, gp.x_tpl), (gg, gp.x_geo))):
        edgeconv(bx, pq, 4 * H, 2 * s * H, 2 * s * H + H, g, n_frames, ec, wec, s * (H + Dp))
    dense(gp.mlp, ec, 0, wec, R, C=out, c_off=out_off, ldc=ldo)


def run_gcn_rig(ws: Workspace, tag: str, pk: GCNRigPack, pos: torch.Tensor, feature: torch.Tensor,
                feat_lds: int, feat_frame_stride: int, gt: Graph, gg: Graph, binfo: BatchInfo,
                n_frames: int) -> torch.Tensor:
    """GCNRig.forward (models/rignet.py:58-67) for `n_frames` key-frames at once (shared weights and
    graph, models/rignet.py:85-86).  `feature` is [N, feat_lds]; key-frame f reads columns
    [f*feat_frame_stride, f*feat_frame_stride + F).  Returns [n_frames*N, O] (workspace buffer)."""
    dev = pos.device
    n = pos.shape[0]
    R = n * n_frames
    G = n_frames * binfo.n_graphs
    F, ldf = pk.F, pk.ldf
    feat = ws.get(tag + ".feat", (R, ldf), dev, zero=True)       # pad columns stay zero forever
    gather_cols(feature, feat_lds, 0, feat_frame_stride, None, F, n, n_frames, feat, ldf, pk.feat_off)
    gather_cols(pos, 3, 0, 0, None, 3, n, n_frames, feat, ldf, pk.pos_off)
    pqpos = ws.get(tag + ".pqpos", (n, pk.pq_pos.N), dev)
    dense(pk.pq_pos, pos, 0, 3, n, C=pqpos, ldc=pk.pq_pos.N)
    # GCU chain: input of gcu_1 is the feature block, of gcu_2 / gcu_3 the previous output block
    srcs = [(pk.feat_off, pk.gcus[0].pq_x.K), (pk.x_off[0], pk.gcus[1].pq_x.K), (pk.x_off[1], pk.gcus[2].pq_x.K)]
    tags = [f"{tag}.gcu{k}" for k in range(len(pk.gcus))]
    xg = ws.get(tag + ".xg", (G, pk.glb.N), dev)
    run_pos_branches(ws, tags, pk.gcus, pqpos, gt, gg, n, n_frames, also_fill=(xg,))
    for k, gp in enumerate(pk.gcus):
        run_gcu(ws, tags[k], gp, feat, srcs[k][0], ldf, srcs[k][1], gt, gg, n, n_frames, feat, pk.x_off[k], ldf)
    dense(pk.glb, feat, 0, ldf, R, pool=xg, binfo=binfo, n_vtx=n)                     # x_4 is never stored
    gb = ws.get(tag + ".gb", (G, pk.t0_global.N), dev)
    dense(pk.t0_global, xg, 0, pk.glb.N, G, C=gb, ldc=pk.t0_global.N)
    h1 = ws.get(tag + ".h1", (R, pk.t0.N), dev)
    dense(pk.t0, feat, 0, ldf, R, C=h1, ldc=pk.t0.N, rowbias=gb, binfo=binfo, n_vtx=n)
    h2 = ws.get(tag + ".h2", (R, pk.t1.N), dev)
    dense(pk.t1, h1, 0, pk.t0.N, R, C=h2, ldc=pk.t1.N)
    out = ws.get(tag + ".out", (R, pk.O), dev)
    dense(pk.head, h2, 0, pk.t1.N, R, C=out, ldc=pk.O)
    return out


def temporal_attn(pk: AttnPack, x: torch.Tensor, out: torch.Tensor) -> None:
    n, t, c = x.shape
    tok = _begin("temporal_attn", 1, 0.0, 4.0 * n * (t * c + pk.D)) if (_counter is not None or _timer is not None) else None
    _lib.check(_lib.load().morig_temporal_attn_fwd(x.data_ptr(), n, t, c, pk.heads, pk.D, pk.u.data_ptr(),
                                                   pk.l0.data_ptr(), pk.Mv.data_ptr(), pk.c0.data_ptr(),
                                                   out.data_ptr(), out.shape[1], _amax_out(out), _lib.stream_ptr()),
               "morig_temporal_attn_fwd")
    _end(tok)


def frame_reduce(x: torch.Tensor, mode: str, out: torch.Tensor) -> None:
    n, t, c = x.shape
    _lib.check(_lib.load().morig_frame_reduce(x.data_ptr(), n, t, c, 0 if mode == "mean" else 1, out.data_ptr(),
                                              out.shape[1], _lib.stream_ptr()), "morig_frame_reduce")
    _untracked(out)


def run_temporal_attn(ws: Workspace, tag: str, pk: AttnPack, x: torch.Tensor, out: torch.Tensor) -> None:
    """TemporalAttn.forward (models/rignet.py:36-46): x [N, T, C] -> out [N, ff1.N] (not normalised)."""
    n = x.shape[0]
    dev = x.device
    a = ws.get(tag + ".attn", (n, pk.D), dev)
    temporal_attn(pk, x, a)
    h = ws.get(tag + ".ff", (n, pk.ff0.N), dev)
    dense(pk.ff0, a, 0, pk.D, n, C=h, ldc=pk.ff0.N)
    dense(pk.ff1, h, 0, pk.ff0.N, n, C=out, ldc=out.shape[1])


def run_skin(ws: Workspace, tag: str, pk: SkinPack, pos: torch.Tensor, skin_input: torch.Tensor,
             motion: torch.Tensor, gt: Graph, gg: Graph, binfo: BatchInfo) -> torch.Tensor:
    """SkinNet_inner.forward (models/rignet.py:158-182). Returns [N, nearest_bone] logits."""
    dev = pos.device
    n = pos.shape[0]
    B = binfo.n_graphs
    raw = ws.get(tag + ".raw", (n, pk.k_pos), dev, zero=True)                  # [pos | samples | 0-pad]
    gather_cols(pos, 3, 0, 0, None, 3, n, 1, raw, pk.k_pos, 0)
    gather_cols(skin_input, skin_input.shape[1], 0, 0, pk.skin_cols, pk.in_pos - 3, n, 1, raw, pk.k_pos, 3)
    pqpos = ws.get(tag + ".pqpos", (n, pk.pq_pos.N), dev)
    dense(pk.pq_pos, raw, 0, pk.k_pos, n, C=pqpos, ldc=pk.pq_pos.N)
    c = pk.gcus[0].out
    xs = [ws.get(f"{tag}.x{k}", (n, c), dev) for k in range(3)]
    tags = [f"{tag}.gcu{k}" for k in range(3)]
    xg = ws.get(tag + ".xg", (B, pk.g1.N), dev)
    run_pos_branches(ws, tags, pk.gcus, pqpos, gt, gg, n, 1, also_fill=(xg,))
    run_gcu(ws, tags[0], pk.gcus[0], motion, 0, motion.shape[1], pk.gcus[0].pq_x.K, gt, gg, n, 1, xs[0], 0, c)
    g0 = ws.get(tag + ".g0", (n, pk.g0.N), dev)
    dense(pk.g0, xs[0], 0, c, n, C=g0, ldc=pk.g0.N)
    dense(pk.g1, g0, 0, pk.g0.N, n, pool=xg, binfo=binfo, n_vtx=n)
    run_gcu(ws, tags[1], pk.gcus[1], xs[0], 0, c, c, gt, gg, n, 1, xs[1], 0, c)
    run_gcu(ws, tags[2], pk.gcus[2], xs[1], 0, c, c, gt, gg, n, 1, xs[2], 0, c)
    gb = ws.get(tag + ".gb", (B, pk.c0_global.N), dev)
    dense(pk.c0_global, xg, 0, pk.g1.N, B, C=gb, ldc=pk.c0_global.N)
    h1 = ws.get(tag + ".h1", (n, pk.c0.N), dev)
    dense(pk.c0, xs[2], 0, c, n, C=h1, ldc=pk.c0.N, rowbias=gb, binfo=binfo, n_vtx=n)
    h2 = ws.get(tag + ".h2", (n, pk.c1.N), dev)
    dense(pk.c1, h1, 0, pk.c0.N, n, C=h2, ldc=pk.c1.N)
    out = ws.get(tag + ".out", (n, pk.head.N), dev)
    dense(pk.head, h2, 0, pk.c1.N, n, C=out, ldc=pk.head.N)
    return out
