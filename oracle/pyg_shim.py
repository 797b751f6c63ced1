"""TEST INFRASTRUCTURE — not part of the product path.

Stand-ins for the two third-party packages the reference's hot path imports but this image
does not have: `torch_geometric` 2.0.4 and `torch_scatter` 2.0.9 (`environment.yml:91,103`).
Installing them into `sys.modules` lets the reference's own, unmodified
`models/rignet.py` + `models/basic_modules.py` import and run on CPU, which is how the
restatement in `oracle/rignet_port.py` and the fixtures in `tests/golden/` are pinned.

Semantics restated from the published behaviour of those releases (their sources are not in
/root/reference, so this part is "restated, unpinned" — see DESIGN.md):

* `remove_self_loops(ei)`   keep columns with ei[0] != ei[1], order preserved
                             (call site `models/basic_modules.py:188`)
* `add_self_loops(ei, num_nodes=N)`  append [[0..N-1],[0..N-1]] (call site `:189`)
* `MessagePassing(aggr='max').propagate(ei, **kw)`  flow source_to_target: `*_j = t[ei[0]]`,
  `*_i = t[ei[1]]`, `message(...)`, then scatter-max over `ei[1]` with `dim_size=N`,
  empty segments -> 0, then `update(...)`  (call site `:190`)
* `scatter_max(src, index, dim=0)` -> (max per index, argmax); first maximal element wins,
  untouched rows -> 0 / argmax = src.size(0)  (call sites `models/rignet.py:63,176`); its gradient goes to
  that first maximal element only (torch_scatter's backward gathers grad_out by argmax)
"""
from __future__ import annotations

import inspect
import sys
import types

import torch


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else int(num_nodes)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), edge_attr


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and src.dim() == 2 and index.dim() == 1
    n = (int(index.max()) + 1 if index.numel() else 0) if dim_size is None else int(dim_size)
    lowest = torch.finfo(src.dtype).min
    idx2 = index.unsqueeze(1).expand_as(src)
    with torch.no_grad():
        res = torch.full((n, src.shape[1]), lowest, dtype=src.dtype, device=src.device)
        res = res.scatter_reduce(0, idx2, src, reduce="amax", include_self=True)
        # argmax: first row (in src order) attaining the max of its segment
        hit = src == res[index]
        rows = torch.arange(src.shape[0], device=src.device).unsqueeze(1).expand_as(src)
        cand = torch.where(hit, rows, torch.full_like(rows, src.shape[0]))
        arg = torch.full((n, src.shape[1]), src.shape[0], dtype=torch.long, device=src.device)
        arg = arg.scatter_reduce(0, idx2, cand, reduce="amin", include_self=True)
        touched = torch.zeros(n, dtype=torch.bool, device=src.device)
        touched[index] = True
        res = torch.where(touched.unsqueeze(1), res, torch.zeros_like(res))
    if torch.is_grad_enabled() and src.requires_grad and src.shape[0] > 0:
        # torch_scatter's backward sends the gradient of out[s, c] to src[arg[s, c], c] only (the first maximal
        # element) -- unlike torch's scatter_reduce('amax'), which splits it evenly among ties.  Same values,
        # re-expressed as a gather so that autograd follows that rule.
        picked = src.gather(0, arg.clamp(max=src.shape[0] - 1))
        res = torch.where(arg < src.shape[0], picked, torch.zeros_like(picked))
    return res, arg


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else int(dim_size)
    res = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        if flow != "source_to_target":
            raise NotImplementedError(flow)
        self.aggr = aggr
        self._msg_args = list(inspect.signature(self.message).parameters)

    def propagate(self, edge_index, size=None, **kwargs):
        if self.aggr != "max":
            raise NotImplementedError(self.aggr)
        src, dst = edge_index[0], edge_index[1]
        n = next(iter(kwargs.values())).size(0)
        feed = {}
        for name in self._msg_args:
            base, side = name[:-2], name[-2:]
            feed[name] = kwargs[base].index_select(0, dst if side == "_i" else src)
        msg = self.message(**feed)
        pooled, _ = scatter_max(msg, dst, dim=0, dim_size=n)
        return self.update(pooled)

    def message(self, x_j):  # pragma: no cover - overridden by the reference classes
        return x_j

    def update(self, aggr_out):
        return aggr_out


class Data:
    """minimal stand-in for `torch_geometric.data.Data` / a collated `Batch`: an attribute bag with `.to(device)`"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device, *a, **k):
        return Data(**{key: (v.to(device) if torch.is_tensor(v) else v) for key, v in self.__dict__.items()})


class InMemoryDataset:
    def __init__(self, *a, **k):
        raise NotImplementedError("datasets are external downloads; the oracle shim feeds synthetic batches")


class DataLoader(list):
    """`torch_geometric.loader.DataLoader` stand-in: a list of already collated batches"""

    def __init__(self, batches=(), *a, **k):
        super().__init__(batches)


def _unavailable(*_a, **_k):
    raise NotImplementedError("not on the rignet.py forward path; not provided by the oracle shim")


def install() -> None:
    """Register the stand-in modules (idempotent)."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "torch_geometric" in sys.modules and not getattr(sys.modules["torch_geometric"], "_morig_shim", False):
        return  # a real PyG is importable: use it
    tg = mod("torch_geometric", _morig_shim=True)
    tg.nn = mod("torch_geometric.nn", MessagePassing=MessagePassing, knn=_unavailable,
                knn_interpolate=_unavailable, fps=_unavailable, radius=_unavailable,
                global_max_pool=_unavailable, PointConv=_unavailable)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=MessagePassing)
    tg.utils = mod("torch_geometric.utils", remove_self_loops=remove_self_loops,
                   add_self_loops=add_self_loops, softmax=_unavailable)
    tg.loader = mod("torch_geometric.loader", DataLoader=DataLoader)
    tg.data = mod("torch_geometric.data", Data=Data, InMemoryDataset=InMemoryDataset)
    mod("torch_scatter", scatter_max=scatter_max, scatter_add=scatter_add)
    mod("torch_cluster", fps=_unavailable, knn=_unavailable, radius=_unavailable)


REFERENCE_ROOT = "/root/reference"


def import_reference_models():
    """Import the reference's `models` package unmodified (build container only: the GPU box has no
    /root/reference). Returns the module; raises FileNotFoundError where the reference is absent."""
    import os
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "models")):
        raise FileNotFoundError(REFERENCE_ROOT)
    install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models  # noqa: WPS433  (the reference package)
    return models
