// Surface-geodesic graph build: the reference computes, per mesh, all-pairs geodesic
// distances between ~4000 surface samples with numpy + scipy's Dijkstra and keeps, per vertex, the neighbours inside a
// geodesic ball (data_proc/common_ops.py:176-226; run offline and at inference, evaluate/joint2rig.py:505) -- minutes
// per mesh.  Here, given the surface samples and their normals (the Poisson-disk sampling itself is open3d's and is not
// rebuilt):
//
//   geo_knn_kernel        5 nearest samples of every sample by fp64 Euclidean distance, ordered by (distance, index),
//                         filtered by the normal test cos(n_p, n_q) > -0.5                       (common_ops.py:184-194)
//   geo_count / geo_fill  CSR of the undirected sample graph, float32-rounded edge lengths       (:187, :194)
//   geo_apsp_kernel       all-pairs shortest paths: one CTA per source, the distance row lives in shared memory and is
//                         relaxed until nothing changes.  Edge weights are non-negative and fp64 addition is monotone,
//                         so the fixed point d[v] = min_u fl(d[u] + w_uv) is unique: the result equals scipy's
//                         Dijkstra (:195) bit for bit, whatever the relaxation order.
//   geo_vert_nn_kernel    nearest sample of every mesh vertex (first minimum)                    (:204-205)
//   geo_gather_kernel     surface_geodesic[a][b] = D[nn[a]][nn[b]], unreachable pairs -> 8 + Euclidean (:199-206)
//   geo_ball_kernel       per vertex the neighbours with geodesic distance <= radius (self excluded), at most max_nn:
//                         all of them in ascending index order when they fit (the reference's deterministic case),
//                         otherwise the max_nn nearest (the reference draws a random subset there, :221)   (:214-226)
#include <math.h>
#include "common.cuh"

namespace morig {

constexpr int GEO_K = 5;

// numpy evaluates (dx^2 + dy^2) + dz^2 with separately rounded products; __dmul_rn / __dadd_rn are never contracted
// into FMAs, so the distances -- and everything derived from them -- are bit-identical to the reference's
__device__ __forceinline__ double sumsq3(double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}
__device__ __forceinline__ double dist3(const double *a, const double *b) {
    return sqrt(sumsq3(a[0] - b[0], a[1] - b[1], a[2] - b[2]));
}

// one warp per sample: every lane keeps its own sorted top-(K+1) over a strided share, then a 32-way merge
__global__ void __launch_bounds__(256) geo_knn_kernel(const double *__restrict__ pts, const double *__restrict__ nrm,
                                                      int S, int32_t *__restrict__ nbr, float *__restrict__ wgt) {
    constexpr int KK = GEO_K + 1;                               // position 0 of the sorted row is skipped (:188)
    const int lane = threadIdx.x & 31;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= S) return;
    double bd[KK];
    int bi[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) { bd[k] = INFINITY; bi[k] = 0x7fffffff; }
    const double pp[3] = {pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]};
    for (int q = lane; q < S; q += 32) {
        const double d = dist3(pp, pts + 3 * q);
        if (d < bd[KK - 1] || (d == bd[KK - 1] && q < bi[KK - 1])) {
            bd[KK - 1] = d; bi[KK - 1] = q;
#pragma unroll
            for (int k = KK - 1; k > 0; --k) {
                if (bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1])) {
                    const double td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                    const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                }
            }
        }
    }
    // merge: KK rounds of warp arg-min over the lanes' list heads
    int head = 0;
    for (int r = 0; r < KK; ++r) {
        double d = INFINITY; int i = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < KK; ++k) if (k == head) { d = bd[k]; i = bi[k]; }
        double md = d; int mi = i;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, md, off);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
        }
        if (d == md && i == mi && mi != 0x7fffffff) ++head;      // exactly one lane owns the winner (indices are unique)
        if (r >= 1 && lane == 0) {
            int q = (mi == 0x7fffffff) ? -1 : mi;
            float w = 0.f;
            if (q >= 0) {
                const double *a = nrm + 3 * q, *b = nrm + 3 * p;
                const double dot = __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
                const double na = sqrt(sumsq3(a[0], a[1], a[2])), nb = sqrt(sumsq3(b[0], b[1], b[2]));
                const double cs = dot / __dadd_rn(__dmul_rn(na, nb), 1e-10);
                w = (float)md;                                   // the reference stores the lengths in a float32 matrix
                if (!(cs > -0.5) || w == 0.f) q = -1;            // a zero entry is "no edge" for the sparse matrix
            }
            nbr[p * GEO_K + r - 1] = q;
            wgt[p * GEO_K + r - 1] = w;
        }
    }
}

__global__ void geo_count_kernel(const int32_t *__restrict__ nbr, int S, int32_t *cnt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * GEO_K) return;
    const int q = nbr[e];
    if (q < 0) return;
    atomicAdd(&cnt[e / GEO_K], 1);
    atomicAdd(&cnt[q], 1);
}

__global__ void geo_scan_kernel(const int32_t *__restrict__ cnt, int n, int32_t *rowptr, int32_t *cursor) {
    __shared__ int32_t part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    int32_t s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {
        const int32_t v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int32_t run = part[t] - s;
    for (int i = lo; i < hi; ++i) { rowptr[i] = run; cursor[i] = run; run += cnt[i]; }
    if (t == nt - 1) rowptr[n] = part[t];
}

__global__ void geo_fill_kernel(const int32_t *__restrict__ nbr, const float *__restrict__ wgt, int S, int32_t *cursor,
                                int32_t *adj, float *adjw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * GEO_K) return;
    const int q = nbr[e];
    if (q < 0) return;
    const int p = e / GEO_K;
    const float w = wgt[e];
    int s = atomicAdd(&cursor[p], 1); adj[s] = q; adjw[s] = w;
    s = atomicAdd(&cursor[q], 1); adj[s] = p; adjw[s] = w;
}

// one CTA per source; d[] in dynamic shared memory (S doubles)
__global__ void __launch_bounds__(512) geo_apsp_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ adj,
                                                       const float *__restrict__ adjw, int S, double *__restrict__ D) {
    extern __shared__ double d[];
    const int src = blockIdx.x;
    for (int v = threadIdx.x; v < S; v += blockDim.x) d[v] = (v == src) ? 0.0 : INFINITY;
    __syncthreads();
    for (;;) {
        int changed = 0;
        for (int v = threadIdx.x; v < S; v += blockDim.x) {
            double best = d[v];
            const int lo = rowptr[v], hi = rowptr[v + 1];
            for (int e = lo; e < hi; ++e) {
                const double c = d[adj[e]] + (double)adjw[e];
                if (c < best) best = c;
            }
            if (best < d[v]) { d[v] = best; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int v = threadIdx.x; v < S; v += blockDim.x) D[(size_t)src * S + v] = d[v];
}

__global__ void __launch_bounds__(256) geo_vert_nn_kernel(const double *__restrict__ verts, int V,
                                                          const double *__restrict__ pts, int S, int32_t *nn) {
    const int lane = threadIdx.x & 31;
    const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (v >= V) return;
    const double vv[3] = {verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]};
    double md = INFINITY; int mi = 0x7fffffff;
    for (int q = lane; q < S; q += 32) {
        const double dd = dist3(pts + 3 * q, vv);
        if (dd < md) { md = dd; mi = q; }                        // strided scan: the lowest index of a lane's minimum
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, md, off);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
        if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }   // np.argmin: first minimum
    }
    if (lane == 0) nn[v] = mi;
}

__global__ void geo_gather_kernel(const double *__restrict__ D, const double *__restrict__ pts, int S,
                                  const int32_t *__restrict__ nn, int V, double *__restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)V * V) return;
    const int a = (int)(idx / V), b = (int)(idx % V);
    const int sa = nn[a], sb = nn[b];
    double g = D[(size_t)sa * S + sb];
    if (isinf(g)) g = 8.0 + dist3(pts + 3 * sb, pts + 3 * sa);    // :199-202
    out[idx] = g;
}

// one warp per vertex
__global__ void __launch_bounds__(256) geo_ball_kernel(const double *__restrict__ G, int V, double radius, int max_nn,
                                                       int64_t *__restrict__ edges, int32_t *__restrict__ deg) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= V) return;
    const double *row = G + (size_t)i * V;
    int64_t *out = edges + (size_t)i * max_nn * 2;
    // pass 1: size of the ball
    int cnt = 0;
    for (int j = lane; j < V; j += 32) cnt += (j != i && row[j] <= radius) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (cnt <= max_nn) {                                         // ascending index order, like np.argwhere (:219)
        int base = 0;
        for (int j0 = 0; j0 < V; j0 += 32) {
            const int j = j0 + lane;
            const bool in = j < V && j != i && row[j] <= radius;
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                out[2 * pos] = i; out[2 * pos + 1] = j;
            }
            base += __popc(m);
        }
        if (lane == 0) deg[i] = cnt;
        return;
    }
    // more than max_nn: the max_nn nearest by (distance, index), emitted in that order
    double last_d = -1.0; int last_j = -1;
    for (int r = 0; r < max_nn; ++r) {
        double md = INFINITY; int mj = 0x7fffffff;
        for (int j = lane; j < V; j += 32) {
            const double g = row[j];
            if (j == i || !(g <= radius)) continue;
            if (g < last_d || (g == last_d && j <= last_j)) continue;       // already emitted
            if (g < md || (g == md && j < mj)) { md = g; mj = j; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, md, off);
            const int oj = __shfl_xor_sync(0xffffffffu, mj, off);
            if (od < md || (od == md && oj < mj)) { md = od; mj = oj; }
        }
        if (lane == 0) { out[2 * r] = i; out[2 * r + 1] = mj; }
        last_d = md; last_j = mj;
    }
    if (lane == 0) deg[i] = max_nn;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace morig

using namespace morig;

// workspace layout: D [S*S] f64 | nbr [S*K] i32 | wgt [S*K] f32 | cnt, rowptr(+1), cursor [S] i32 | adj, adjw [2*S*K] | nn [V]
extern "C" MORIG_API size_t morig_surface_geodesic_workspace(int32_t S, int32_t V) {
    if (S <= 0 || V <= 0) return 0;
    size_t b = align256((size_t)S * S * 8);
    b += 2 * align256((size_t)S * GEO_K * 4);
    b += 3 * align256((size_t)(S + 1) * 4);
    b += 2 * align256((size_t)2 * S * GEO_K * 4);
    b += align256((size_t)V * 4);
    return b;
}

extern "C" MORIG_API int morig_surface_geodesic(const double *pts, const double *normals, int32_t S, const double *verts,
                                                int32_t V, double *out, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(pts && normals && verts && out && ws && S > 1 && V > 0, "surface_geodesic: bad argument");
    MORIG_CHECK_ARG(ws_bytes >= morig_surface_geodesic_workspace(S, V), "surface_geodesic: workspace too small");
    MORIG_CHECK_ARG((size_t)S * 8 <= 200 * 1024, "surface_geodesic: S=%d samples exceed the shared-memory distance row", S);
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    double *D = reinterpret_cast<double *>(w); w += align256((size_t)S * S * 8);
    int32_t *nbr = reinterpret_cast<int32_t *>(w); w += align256((size_t)S * GEO_K * 4);
    float *wgt = reinterpret_cast<float *>(w); w += align256((size_t)S * GEO_K * 4);
    int32_t *cnt = reinterpret_cast<int32_t *>(w); w += align256((size_t)(S + 1) * 4);
    int32_t *rowptr = reinterpret_cast<int32_t *>(w); w += align256((size_t)(S + 1) * 4);
    int32_t *cursor = reinterpret_cast<int32_t *>(w); w += align256((size_t)(S + 1) * 4);
    int32_t *adj = reinterpret_cast<int32_t *>(w); w += align256((size_t)2 * S * GEO_K * 4);
    float *adjw = reinterpret_cast<float *>(w); w += align256((size_t)2 * S * GEO_K * 4);
    int32_t *nn = reinterpret_cast<int32_t *>(w);
    const int T = 256;
    geo_knn_kernel<<<ceil_div(S * 32, T), T, 0, stream>>>(pts, normals, S, nbr, wgt);
    MORIG_LAUNCH_CHECK("geo_knn_kernel");
    MORIG_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(S + 1) * 4, stream));
    geo_count_kernel<<<ceil_div(S * GEO_K, T), T, 0, stream>>>(nbr, S, cnt);
    geo_scan_kernel<<<1, 1024, 0, stream>>>(cnt, S, rowptr, cursor);
    geo_fill_kernel<<<ceil_div(S * GEO_K, T), T, 0, stream>>>(nbr, wgt, S, cursor, adj, adjw);
    MORIG_LAUNCH_CHECK("geo_fill_kernel");
    const size_t smem = (size_t)S * 8;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(geo_apsp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured_dev = dev;
    }
    geo_apsp_kernel<<<S, 512, smem, stream>>>(rowptr, adj, adjw, S, D);
    MORIG_LAUNCH_CHECK("geo_apsp_kernel");
    geo_vert_nn_kernel<<<ceil_div(V * 32, T), T, 0, stream>>>(verts, V, pts, S, nn);
    const size_t total = (size_t)V * V;
    geo_gather_kernel<<<(unsigned)ceil_div64((int64_t)total, T), T, 0, stream>>>(D, pts, S, nn, V, out);
    MORIG_LAUNCH_CHECK("geo_gather_kernel");
    return 0;
}

extern "C" MORIG_API int morig_geo_ball_edges(const double *geodesic, int32_t V, double radius, int32_t max_nn,
                                              int64_t *edges, int32_t *degree, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(geodesic && edges && degree && V > 0 && max_nn > 0, "geo_ball_edges: bad argument");
    geo_ball_kernel<<<ceil_div(V * 32, 256), 256, 0, stream>>>(geodesic, V, radius, max_nn, edges, degree);
    MORIG_LAUNCH_CHECK("geo_ball_kernel");
    return 0;
}

// ---- topological edges from the triangle list (data_proc/common_ops.py:15-32 `get_tpl_edges`) -------------------------
// The reference scans the whole face array once per vertex (O(V F) in python).  Here every face emits its six directed
// half-edges as 64-bit keys (v << 32 | n); a radix sort and a unique pass leave, per vertex, its distinct neighbours in
// ascending order.  (The reference lists a vertex's neighbours in python-set iteration order; the edge SET is the same,
// and the networks' max-aggregation does not depend on edge order.)
#include <cub/cub.cuh>

namespace morig {

__global__ void tpl_emit_kernel(const int64_t *__restrict__ faces, int64_t F, uint64_t *keys) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const uint64_t a = (uint64_t)faces[3 * f], b = (uint64_t)faces[3 * f + 1], c = (uint64_t)faces[3 * f + 2];
    const uint64_t SKIP = ~0ull;                                 // degenerate corner pairs (n == v) are not neighbours
    uint64_t *k = keys + 6 * f;
    k[0] = a != b ? (a << 32 | b) : SKIP; k[1] = a != c ? (a << 32 | c) : SKIP;
    k[2] = b != a ? (b << 32 | a) : SKIP; k[3] = b != c ? (b << 32 | c) : SKIP;
    k[4] = c != a ? (c << 32 | a) : SKIP; k[5] = c != b ? (c << 32 | b) : SKIP;
}

__global__ void tpl_unpack_kernel(const uint64_t *__restrict__ keys, const int32_t *__restrict__ n_unique, int64_t cap,
                                  int64_t *edges, int64_t *count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n = *n_unique;
    if (n > 0 && keys[n - 1] == ~0ull) --n;                      // the SKIP key sorts last
    if (i == 0) *count = n;
    if (i >= n || i >= cap) return;
    edges[2 * i] = (int64_t)(keys[i] >> 32);
    edges[2 * i + 1] = (int64_t)(keys[i] & 0xffffffffull);
}

}  // namespace morig

extern "C" MORIG_API size_t morig_tpl_edges_workspace(int64_t F) {
    if (F <= 0) return 0;
    const int64_t n = 6 * F;
    size_t sort_bytes = 0, uniq_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int)n);
    cub::DeviceSelect::Unique(nullptr, uniq_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr, (int)n);
    const size_t tmp = sort_bytes > uniq_bytes ? sort_bytes : uniq_bytes;
    return 3 * align256((size_t)n * 8) + align256(tmp) + 256;
}

// edges [6F, 2] int64 (capacity), count: device int64 receiving the number of rows written
extern "C" MORIG_API int morig_tpl_edges(const int64_t *faces, int64_t F, int64_t *edges, int64_t *count, void *ws,
                                         size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(faces && edges && count && ws && F > 0 && 6 * F < (1ll << 31), "tpl_edges: bad argument");
    MORIG_CHECK_ARG(ws_bytes >= morig_tpl_edges_workspace(F), "tpl_edges: workspace too small");
    const int64_t n = 6 * F;
    uint8_t *w = reinterpret_cast<uint8_t *>(ws);
    uint64_t *k0 = reinterpret_cast<uint64_t *>(w); w += align256((size_t)n * 8);
    uint64_t *k1 = reinterpret_cast<uint64_t *>(w); w += align256((size_t)n * 8);
    uint64_t *k2 = reinterpret_cast<uint64_t *>(w); w += align256((size_t)n * 8);
    int32_t *n_unique = reinterpret_cast<int32_t *>(w); w += 256;
    size_t sort_bytes = 0, uniq_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, k0, k1, (int)n);
    cub::DeviceSelect::Unique(nullptr, uniq_bytes, k1, k2, n_unique, (int)n);
    tpl_emit_kernel<<<(unsigned)ceil_div64(F, 256), 256, 0, stream>>>(faces, F, k0);
    MORIG_LAUNCH_CHECK("tpl_emit_kernel");
    MORIG_CUDA(cub::DeviceRadixSort::SortKeys(w, sort_bytes, k0, k1, (int)n, 0, 64, stream));
    MORIG_CUDA(cub::DeviceSelect::Unique(w, uniq_bytes, k1, k2, n_unique, (int)n, stream));
    tpl_unpack_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, stream>>>(k2, n_unique, n, edges, count);
    MORIG_LAUNCH_CHECK("tpl_unpack_kernel");
    return 0;
}
