"""The product package must not reach into the oracle, the reference, or any CPU / torch fallback."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "morig_b200")


def _py_files():
    for d, _, fs in os.walk(PKG):
        for f in fs:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_product_never_imports_oracle_or_reference():
    for path in _py_files():
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                assert not m.split(".")[0] in ("oracle", "models", "torch_geometric", "torch_scatter", "triton"), (path, m)
        src = open(path).read()
        assert "/root/reference" not in src, path


def test_gpu_side_files_do_not_read_the_reference():
    for rel in ("bench.py", "__graft_entry__.py"):
        assert "/root/reference" not in open(os.path.join(ROOT, rel)).read()


def test_forward_modules_have_no_torch_compute_fallback():
    """the hot modules may allocate tensors but must not call torch compute ops"""
    banned = ("F.linear", "torch.matmul", "torch.relu", ".scatter_reduce", "torch.bmm", "F.batch_norm", "torch.softmax",
              "index_select", "torch.compile")
    for name in ("rignet.py", "basic_modules.py", "engine.py"):
        src = open(os.path.join(PKG, name)).read()
        for b in banned:
            assert b not in src, (name, b)


def test_host_pipeline_refuses_cpu_models():
    """no CPU path: the streaming front-end needs the model on a CUDA device"""
    import pytest
    import morig_b200
    from morig_b200 import synth
    model = morig_b200.jointnet_motion(**synth.ARCH_KWARGS["jointnet_motion"]).eval()
    with pytest.raises(RuntimeError):
        morig_b200.HostPipeline(model)


def test_post_process_and_graph_build_have_no_cpu_path():
    """numpy-in / numpy-out helpers run on the GPU; without one they raise instead of falling back"""
    import numpy as np
    import pytest
    import torch
    from morig_b200 import cluster_utils, graph_build
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    pts = np.zeros((4, 3))
    with pytest.raises(RuntimeError):
        cluster_utils.meanshift_cluster(pts, 0.05)
    with pytest.raises(RuntimeError):
        graph_build.surface_geodesic(pts, pts, pts)
    with pytest.raises(RuntimeError):
        graph_build.geo_ball_edges(np.zeros((4, 4)))
    with pytest.raises(RuntimeError):
        graph_build.tpl_edges(pts, np.zeros((2, 3), dtype=np.int64))
    with pytest.raises(RuntimeError):
        cluster_utils.meanshift_cluster(torch.zeros(4, 3, dtype=torch.float64), 0.05)
