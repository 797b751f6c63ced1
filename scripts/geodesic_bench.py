"""Measurement of the geodesic graph-build row: GPU time (CUDA events) of
`morig_surface_geodesic` + `morig_geo_ball_edges` at the reference's size (4000 samples, 4096-vertex mesh), next to
the reference's numpy + scipy path (oracle port, bit-identical to data_proc/common_ops.py:182-226) on the host
cores at a bounded size.  One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from morig_b200 import graph_build  # noqa: E402
from oracle import gen_golden_geodesic as gg  # noqa: E402
from oracle import geodesic_port  # noqa: E402

dev = torch.device("cuda:0")


def gpu_ms(s, v, iters=5):
    pts, nrm, verts = gg.make_inputs(s, v, 6)
    p, n, vv = (torch.from_numpy(x).to(dev) for x in (pts, nrm, verts))
    g = graph_build.surface_geodesic(p, n, vv)
    graph_build.geo_ball_edges(g, 0.06, 15)
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_geo = t_ball = 0.0
    for _ in range(iters):
        a.record()
        g = graph_build.surface_geodesic(p, n, vv)
        b.record()
        e = graph_build.geo_ball_edges(g, 0.06, 15)
        c.record()
        torch.cuda.synchronize()
        t_geo += a.elapsed_time(b)
        t_ball += b.elapsed_time(c)
    return t_geo / iters, t_ball / iters, int(e.shape[0])


geo_ms, ball_ms, n_edges = gpu_ms(4000, 4096)
small_geo_ms, small_ball_ms, _ = gpu_ms(1500, 1500)
pts, nrm, verts = gg.make_inputs(1500, 1500, 6)
t0 = time.perf_counter()
ref = geodesic_port.surface_geodesic_from_samples(pts, nrm, verts)
t1 = time.perf_counter()
geodesic_port.geo_ball_edges(ref, 0.06, 15)
t2 = time.perf_counter()
print(json.dumps({"row": "calc_surface_geodesic + get_geo_edges (data_proc/common_ops.py:182-226)",
                  "gpu_ms_geodesic_4000_samples_4096_vertices": geo_ms, "gpu_ms_ball_edges_4096_vertices": ball_ms,
                  "geo_edges": n_edges, "gpu_ms_geodesic_1500x1500": small_geo_ms, "gpu_ms_ball_edges_1500": small_ball_ms,
                  "cpu_ref_s_geodesic_1500x1500": t1 - t0, "cpu_ref_s_ball_edges_1500": t2 - t1,
                  "cpu_cores": os.cpu_count(),
                  "note": "reference cost grows ~ S^2 log S (Dijkstra from every sample) + the python loop over samples"}))
