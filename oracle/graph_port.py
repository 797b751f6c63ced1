"""TEST INFRASTRUCTURE — numpy restatement of the graph normalisation the reference redoes inside
every `EdgeConvMotion.forward` (models/basic_modules.py:188-189) plus the target-sorted CSR the CUDA
path derives from it. Integer work: the CUDA `morig_graph_prep` must match this bit for bit.
"""
from __future__ import annotations

import numpy as np


def normalized_edges(edge_index: np.ndarray, n: int) -> np.ndarray:
    """remove_self_loops (order preserved) then add_self_loops(num_nodes=n) appended at the end —
    models/basic_modules.py:188-189 (PyG 2.0.4 semantics, see oracle/pyg_shim.py)."""
    ei = np.asarray(edge_index, dtype=np.int64)
    keep = ei[0] != ei[1]
    loops = np.arange(n, dtype=np.int64)
    return np.concatenate([ei[:, keep], np.stack([loops, loops])], axis=1)


def csr_by_target(edge_index: np.ndarray, n: int):
    """CSR over targets (row 1) of the normalised edge list. Within a target the sources keep the
    order of the normalised list (stable sort), i.e. real edges in input order, self loop last.
    Returns rowptr int32 [n+1], col int32 [E'] (sources)."""
    ei = normalized_edges(edge_index, n)
    order = np.argsort(ei[1], kind="stable")
    col = ei[0][order].astype(np.int32)
    counts = np.bincount(ei[1], minlength=n)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    rowptr[1:] = np.cumsum(counts)
    return rowptr, col


def brute_force_edgeconv_max(msg_fn, pos, x, edge_index, n):
    """Independent O(E) python loop used to check the scatter-max shim on tiny graphs:
    out[i] = max over normalised edges (j -> i) of msg_fn(pos[i], pos[j], x[i], x[j])."""
    ei = normalized_edges(edge_index, n)
    out = None
    seen = np.zeros(n, dtype=bool)
    for e in range(ei.shape[1]):
        j, i = int(ei[0, e]), int(ei[1, e])
        m = msg_fn(pos[i], pos[j], x[i], x[j])
        if out is None:
            out = np.zeros((n, m.shape[0]), dtype=m.dtype)
        out[i] = m if not seen[i] else np.maximum(out[i], m)
        seen[i] = True
    return out
