"""TEST INFRASTRUCTURE (oracle): numpy / scipy restatement of the reference's surface-geodesic graph build,
`data_proc/common_ops.py:176-226` (`calc_surface_geodesic`, `get_geo_edges`), from the point where open3d has
delivered the Poisson-disk samples and their normals (the sampling itself is open3d's and is not restated).

Pinned in the build container (tests/test_oracle_pinning.py) by running the UNMODIFIED reference functions on a
stand-in mesh object that hands out the same samples, and through the committed fixture
tests/golden/geodesic_s400_v150.npz generated that way (oracle/gen_golden_geodesic.py)."""
import numpy as np
from scipy.sparse import lil_matrix
from scipy.sparse.csgraph import dijkstra


def surface_geodesic_from_samples(pts, pts_normal, verts):
    pts = np.asarray(pts, dtype=np.float64)
    pts_normal = np.asarray(pts_normal, dtype=np.float64)
    verts = np.asarray(verts, dtype=np.float64)
    n = len(pts)                                                                              # :183
    verts_dist = np.sqrt(np.sum((pts[np.newaxis, ...] - pts[:, np.newaxis, :]) ** 2, axis=2))   # :184
    verts_nn = np.argsort(verts_dist, axis=1)                                                 # :185
    conn_matrix = lil_matrix((n, n), dtype=np.float32)                                        # :186
    for p in range(n):                                                                        # :188-194
        nn_p = verts_nn[p, 1:6]
        norm_nn_p = np.linalg.norm(pts_normal[nn_p], axis=1)
        norm_p = np.linalg.norm(pts_normal[p])
        cos_similar = np.dot(pts_normal[nn_p], pts_normal[p]) / (norm_nn_p * norm_p + 1e-10)
        nn_p = nn_p[cos_similar > -0.5]
        conn_matrix[p, nn_p] = verts_dist[p, nn_p]
    dist = dijkstra(conn_matrix, directed=False, indices=range(n), return_predecessors=False, unweighted=False)   # :195
    inf_pos = np.argwhere(np.isinf(dist))                                                     # :200-203
    if len(inf_pos) > 0:
        euc = np.sqrt(np.sum((pts[np.newaxis, ...] - pts[:, np.newaxis, :]) ** 2, axis=2))
        dist[inf_pos[:, 0], inf_pos[:, 1]] = 8.0 + euc[inf_pos[:, 0], inf_pos[:, 1]]
    vert_pts_distance = np.sqrt(np.sum((verts[np.newaxis, ...] - pts[:, np.newaxis, :]) ** 2, axis=2))   # :206
    vert_pts_nn = np.argmin(vert_pts_distance, axis=0)                                        # :207
    return dist[vert_pts_nn, :][:, vert_pts_nn]                                               # :208


def geo_ball_edges(surface_geodesic, radius=0.06, max_nn=15):
    """`get_geo_edges` (:214-226) after the geodesic matrix.  Vertices whose ball holds at most `max_nn` neighbours are
    the reference's deterministic case (ascending index order).  For larger balls the reference draws a random subset
    (`np.random.choice`, :221); the deterministic rule used instead -- here and in the CUDA path -- keeps the `max_nn`
    nearest by (distance, index)."""
    g = np.array(surface_geodesic, dtype=np.float64)
    g += 10.0 * np.eye(len(g))                                                                # :218
    edge_index = []
    for i in range(len(g)):
        ball = np.argwhere(g[i, :] <= radius).squeeze(1)                                      # :220
        if len(ball) > max_nn:
            order = np.lexsort((ball, g[i, ball]))
            ball = ball[order[:max_nn]]
        edge_index.append(np.concatenate((np.repeat(i, len(ball))[:, np.newaxis], ball[:, np.newaxis]), axis=1))   # :223
    return np.concatenate(edge_index, axis=0)


def tpl_edges(obj_v, obj_f):
    """`get_tpl_edges` (data_proc/common_ops.py:15-32): for every vertex, the distinct other corners of the faces it
    belongs to, as rows (v, n); neighbours of a vertex in python-set iteration order, vertices without faces skipped"""
    rows = []
    for v in range(len(obj_v)):
        faces_of_v = np.argwhere(obj_f == v)[:, 0]                                            # :18
        others = [obj_f[f, c] for f in faces_of_v for c in range(3) if obj_f[f, c] != v]      # :20-23
        uniq = list(set(others))                                                              # :24 (set order kept)
        if uniq:
            rows.append(np.array([[v, n] for n in uniq]))                                     # :25-28
    return np.concatenate(rows, axis=0)                                                       # :31


def sorted_rows(e):
    """rows in (v, n) order: the canonical form in which edge lists are compared (a vertex's neighbours come out of
    the reference in python-set iteration order)"""
    e = np.asarray(e, dtype=np.int64)
    return e[np.lexsort((e[:, 1], e[:, 0]))]


def geo_edges_random_subset(surface_geodesic, radius=0.06, max_nn=15):
    """`get_geo_edges` (data_proc/common_ops.py:214-226) after the geodesic matrix, WITH the reference's subset rule:
    `np.random.choice(members, max_nn, replace=False)` from numpy's global generator, in vertex order"""
    g = np.array(surface_geodesic, dtype=np.float64)
    g += 10.0 * np.eye(len(g))
    edge_index = []
    for i in range(len(g)):
        ball = np.argwhere(g[i, :] <= radius).squeeze(1)
        if len(ball) > max_nn:
            ball = np.random.choice(ball, max_nn, replace=False)
        edge_index.append(np.concatenate((np.repeat(i, len(ball))[:, np.newaxis], ball[:, np.newaxis]), axis=1))
    return np.concatenate(edge_index, axis=0)
