"""Generates tests/golden/geodesic_s400_v150.npz by running the UNMODIFIED reference functions
`data_proc.common_ops.calc_surface_geodesic` / `get_geo_edges` (/root/reference, build container only; `open3d` is
absent, so a stand-in module and a stand-in mesh object hand the reference the pre-drawn samples)."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load_reference():
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import data_proc.common_ops as co
    return co


class FakeSamples:
    def __init__(self, pts, normals):
        self.points, self.normals = pts, normals

    def estimate_normals(self):
        pass


class FakeMesh:
    """the two attributes / methods calc_surface_geodesic touches (data_proc/common_ops.py:178-181, 204)"""

    def __init__(self, verts, pts, normals):
        self.vertices, self._s = verts, FakeSamples(pts, normals)

    def sample_points_poisson_disk(self, number_of_points):
        return self._s


def make_inputs(n_samples, n_verts, seed, two_parts=False):
    """samples on a torus (plus, optionally, a far-away second torus: unreachable pairs), outward normals with a few
    flipped ones (normal filter), mesh vertices = jittered subset of the surface"""
    rng = np.random.default_rng(seed)

    def torus(n, shift):
        u, v = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, 2 * np.pi, n)
        R, r = 0.35, 0.12
        p = np.stack([(R + r * np.cos(v)) * np.cos(u), r * np.sin(v), (R + r * np.cos(v)) * np.sin(u)], 1) + shift
        nrm = np.stack([np.cos(v) * np.cos(u), np.sin(v), np.cos(v) * np.sin(u)], 1)
        return p, nrm
    if two_parts:
        p1, n1 = torus(n_samples // 2, np.zeros(3))
        p2, n2 = torus(n_samples - n_samples // 2, np.array([2.0, 0.0, 0.0]))
        pts, nrm = np.concatenate([p1, p2]), np.concatenate([n1, n2])
    else:
        pts, nrm = torus(n_samples, np.zeros(3))
    flip = rng.uniform(size=len(pts)) < 0.03
    nrm[flip] *= -1.0
    verts = pts[rng.choice(len(pts), n_verts, replace=n_verts > len(pts))] + rng.normal(0, 0.004, size=(n_verts, 3))
    return pts, nrm, verts


if __name__ == "__main__":
    co = load_reference()
    pts, nrm, verts = make_inputs(400, 150, 3, two_parts=True)
    geo = co.calc_surface_geodesic(FakeMesh(verts, pts, nrm))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "geodesic_s400_v150.npz"), pts=pts, normals=nrm, verts=verts,
                        surface_geodesic=geo)
    print("wrote geodesic_s400_v150.npz", geo.shape, float(geo.max()))
