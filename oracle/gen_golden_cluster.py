"""Generates tests/golden/meanshift_n300.npz by running the UNMODIFIED reference function
`utils.cluster_utils.meanshift_cluster` (/root/reference, build container only) on seeded inputs shaped like
evaluate/eval_rigging.py:80-91: shifted points clustered around a few joints, reflected, attention weights [N,1]."""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from utils.cluster_utils import meanshift_cluster  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_inputs(n_half, seed):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-0.4, 0.4, size=(6, 3))
    pts = centres[rng.integers(0, 6, n_half)] + rng.normal(0, 0.02, size=(n_half, 3))
    pts = np.concatenate([pts, pts * np.array([[-1, 1, 1]])], axis=0)          # eval_rigging.py:86-87
    attn = rng.uniform(0.05, 1.0, size=(n_half, 1)).astype(np.float32)
    attn = np.tile(attn, (2, 1))                                                # :88
    return pts, attn


if __name__ == "__main__":
    pts, attn = make_inputs(150, 7)
    out = meanshift_cluster(pts.copy(), 0.045, attn, max_iter=30)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "meanshift_n300.npz"), pts=pts, attn=attn,
                        bandwidth=0.045, max_iter=30, out=out)
    print("wrote meanshift_n300.npz", out.shape)
