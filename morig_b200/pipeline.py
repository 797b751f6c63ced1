"""Host-to-host streaming of batches through one of the rigging networks.

The reference's loops (`training/train_rig.py:206-258`, `evaluate/joint2rig.py:447`) do, per batch,
`data.to(device)` -> `model(data, flow)` -> `.to("cpu")`, strictly one after the other.  `HostPipeline`
keeps that per-batch contract (every batch is copied in from host memory, every result is copied
back out) but runs the three phases of consecutive batches concurrently on three CUDA streams:

    copy-in stream   pinned host batch -> per-slot device staging tensors
    compute stream   model(data, flow)   (the caller's current stream; CUDA-graph replay once warm)
    copy-out stream  the three outputs -> per-slot pinned host tensors

    pipe = HostPipeline(model, depth=2)
    for motion_all, motion_aggr, pred in pipe.run(host_batches):      # results in submission order
        ...                                                            # pinned CPU tensors, valid until
                                                                       # `depth` more batches were submitted
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Tuple

import torch

_FIELDS = ("pos", "tpl_edge_index", "geo_edge_index", "batch", "skin_input")


class _Slot:
    def __init__(self):
        self.dev_in = {}              # field -> device staging tensor
        self.dev_flow: Optional[torch.Tensor] = None
        self.host_out: Optional[List[torch.Tensor]] = None
        self.dev_out = None           # keeps the output tensors alive until their copy-out finished
        self.compute_done: Optional[torch.cuda.Event] = None
        self.out_done: Optional[torch.cuda.Event] = None
        self.pending = False


class _DeviceBatch:
    """attribute bag handed to the model (only attribute access is used, models/rignet.py:83-86)"""


def _staging(cache: dict, name: str, src: torch.Tensor, dev) -> torch.Tensor:
    t = cache.get(name)
    if t is None or t.shape != src.shape or t.dtype != src.dtype:
        t = cache[name] = torch.empty(src.shape, dtype=src.dtype, device=dev)
    return t


class HostPipeline:
    def __init__(self, model: torch.nn.Module, depth: int = 2, device=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.model = model
        self.dev = torch.device(device) if device is not None else next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("morig_b200.HostPipeline needs the model on a CUDA device (there is no CPU path)")
        self.copy_in = torch.cuda.Stream(self.dev)
        self.copy_out = torch.cuda.Stream(self.dev)
        self.slots = [_Slot() for _ in range(depth)]
        self.n_submitted = 0
        self.n_fetched = 0

    # -- one batch in ------------------------------------------------------------------------------
    def submit(self, data, input_flow: torch.Tensor) -> None:
        """Queue one HOST batch (`data` attributes and `input_flow` are CPU tensors, ideally pinned)."""
        slot = self.slots[self.n_submitted % len(self.slots)]
        if slot.pending:
            raise RuntimeError("HostPipeline: all slots in flight; call result() before the next submit()")
        compute = torch.cuda.current_stream(self.dev)
        # the staging tensors of this slot are free once the forward that last read them has finished
        if slot.compute_done is not None:
            self.copy_in.wait_event(slot.compute_done)
        else:
            self.copy_in.wait_stream(compute)
        d = _DeviceBatch()
        with torch.cuda.stream(self.copy_in):
            for f in _FIELDS:
                src = getattr(data, f, None)
                if torch.is_tensor(src):
                    if src.is_cuda:
                        raise TypeError(f"HostPipeline.submit: data.{f} must be a host tensor")
                    dst = _staging(slot.dev_in, f, src, self.dev)
                    dst.copy_(src, non_blocking=True)
                    setattr(d, f, dst)
            slot.dev_flow = _staging(slot.dev_in, "__flow__", input_flow, self.dev)
            slot.dev_flow.copy_(input_flow, non_blocking=True)
        ng = getattr(data, "num_graphs", None)
        if ng is not None:
            d.num_graphs = ng
        compute.wait_stream(self.copy_in)
        with torch.no_grad():
            outs = self.model(d, slot.dev_flow)
        slot.compute_done = torch.cuda.Event()
        slot.compute_done.record(compute)
        if slot.host_out is None or any(h.shape != o.shape for h, o in zip(slot.host_out, outs)):
            slot.host_out = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
        self.copy_out.wait_event(slot.compute_done)
        with torch.cuda.stream(self.copy_out):
            for h, o in zip(slot.host_out, outs):
                h.copy_(o, non_blocking=True)
        slot.out_done = torch.cuda.Event()
        slot.out_done.record(self.copy_out)
        slot.dev_out = outs
        slot.pending = True
        self.n_submitted += 1

    # -- one result out ----------------------------------------------------------------------------
    def result(self) -> Tuple[torch.Tensor, ...]:
        """Outputs of the oldest batch in flight as pinned host tensors (blocks until they have landed)."""
        if self.n_fetched >= self.n_submitted:
            raise RuntimeError("HostPipeline.result: nothing in flight")
        slot = self.slots[self.n_fetched % len(self.slots)]
        slot.out_done.synchronize()
        slot.dev_out = None
        slot.pending = False
        self.n_fetched += 1
        return tuple(slot.host_out)

    @property
    def in_flight(self) -> int:
        return self.n_submitted - self.n_fetched

    def run(self, batches: Iterable, flow_attr: str = "pred_flow") -> Iterator[Tuple[torch.Tensor, ...]]:
        """Yield the outputs of every batch of `batches` in order.  Items are `data` objects carrying the
        flow as `data.<flow_attr>` or `(data, input_flow)` pairs."""
        for item in batches:
            data, flow = item if isinstance(item, tuple) else (item, getattr(item, flow_attr))
            if self.in_flight == len(self.slots):
                yield self.result()
            self.submit(data, flow)
        while self.in_flight:
            yield self.result()

    def join(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """make `stream` (default: current) wait for everything queued on the copy streams"""
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        s.wait_stream(self.copy_in)
        s.wait_stream(self.copy_out)
