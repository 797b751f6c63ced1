// Graph preparation: int64 [2,E] edge list -> self-loop-normalised, target-sorted CSR (int32).
// Replaces remove_self_loops/add_self_loops of models/basic_modules.py:188-189 (run 36x per forward
// by the reference) with one pass per edge set; also a brute-force kNN graph builder used for the
// synthetic geodesic stand-in.  Integer work: results are bit-exact against oracle/graph_port.py.
#include <cub/cub.cuh>
#include "common.cuh"

namespace morig {

// Stable target-sorted CSR = a STABLE sort of the normalised edge list by target.  Three steps, all parallel over edges and
// independent of the in-degree distribution (a hub with 64 K in-edges costs the same as 64 K ordinary edges):
//   gp_keys_kernel    slot e < E:  key = target (or the sentinel N for dropped edges: self loops, out-of-range ids),
//                     value = source;  slot E + v: the appended self loop (v, v) -- after all real edges, so the stable
//                     sort leaves it last inside its target, exactly like add_self_loops' append
//   cub radix sort    keys -> tgt, values -> col   (LSD radix sort is stable; only ceil(log2(N + 1)) key bits are sorted)
//   gp_rowptr_kernel  rowptr[v] = first slot whose key >= v (binary search); rowptr[N] = E' = first sentinel
__global__ void gp_keys_kernel(const int64_t *__restrict__ ei, int64_t E, int32_t n, int32_t *__restrict__ keys,
                               int32_t *__restrict__ vals) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (e >= E + n) return;
    if (e < E) {
        const int64_t j = ei[e], i = ei[E + e];
        const bool drop = (j == i || i < 0 || i >= n || j < 0 || j >= n);
        keys[e] = drop ? n : (int32_t)i;
        vals[e] = drop ? 0 : (int32_t)j;
    } else {
        keys[e] = vals[e] = (int32_t)(e - E);
    }
}

__global__ void gp_rowptr_kernel(const int32_t *__restrict__ tgt, int32_t total, int32_t n, int32_t *__restrict__ rowptr) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (v > n) return;
    int lo = 0, hi = total;                          // first slot with tgt[slot] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tgt[mid] < v) lo = mid + 1; else hi = mid;
    }
    rowptr[v] = lo;
}

static int key_bits(int32_t n) {                     // bits needed for keys 0..n (n = sentinel)
    int b = 1;
    while (((int64_t)1 << b) <= (int64_t)n) ++b;
    return b;
}

static size_t sort_temp_bytes(int64_t total, int32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t *)nullptr, (int32_t *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)total, 0, key_bits(n));
    return bytes;
}

// ---- brute-force kNN inside each graph: one warp per query vertex --------------------------------
// Each lane scans a strided share of the graph's vertices and keeps its own sorted top-k in
// registers/local memory; the warp then merges the 32 lists by repeated arg-min.
// Ordering key = (distance, index): ties go to the lower index, exactly like a stable sort.
template <int KMAX>
__global__ void knn_kernel(const float *__restrict__ pos, const int32_t *__restrict__ gptr, int32_t n_graphs,
                           int32_t n, int32_t k, int64_t *__restrict__ ei) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int q = warp;
    int g = 0;  // graph of q (B is small: linear search)
    while (g + 1 < n_graphs && gptr[g + 1] <= q) ++g;
    const int lo = gptr[g], hi = gptr[g + 1];
    const float qx = pos[3 * q], qy = pos[3 * q + 1], qz = pos[3 * q + 2];
    float bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) { bd[s] = __int_as_float(0x7f800000); bi[s] = 0x7fffffff; }
    for (int v = lo + lane; v < hi; v += 32) {
        if (v == q) continue;
        // same expression and evaluation order as the host generator: ((a-b)**2).sum(-1)
        float dx = qx - pos[3 * v], dy = qy - pos[3 * v + 1], dz = qz - pos[3 * v + 2];
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d < bd[KMAX - 1] || (d == bd[KMAX - 1] && v < bi[KMAX - 1])) {
            bd[KMAX - 1] = d; bi[KMAX - 1] = v;
#pragma unroll
            for (int s = KMAX - 1; s > 0; --s) {
                bool sw = bd[s] < bd[s - 1] || (bd[s] == bd[s - 1] && bi[s] < bi[s - 1]);
                if (sw) {
                    float td = bd[s]; bd[s] = bd[s - 1]; bd[s - 1] = td;
                    int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
                }
            }
        }
    }
    // merge: k rounds of warp arg-min over each lane's current head
    int head = 0;
    for (int r = 0; r < k; ++r) {
        float d = __int_as_float(0x7f800000);
        int i = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < KMAX; ++s)
            if (s == head) { d = bd[s]; i = bi[s]; }
        float md = d; int mi = i;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float od = __shfl_xor_sync(0xffffffffu, md, off);
            int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (od < md || (od == md && oi < mi)) { md = od; mi = oi; }
        }
        if (mi == i && md == d && i != 0x7fffffff) ++head;  // the winning lane pops its head
        if (lane == 0) {
            ei[(int64_t)q * k + r] = q;
            ei[(int64_t)n * k + (int64_t)q * k + r] = (mi == 0x7fffffff) ? q : mi;
        }
    }
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API size_t morig_graph_prep_workspace(int64_t E, int32_t N) {
    const int64_t total = E + N;
    return sizeof(int32_t) * (size_t)(2 * total) + 512 + sort_temp_bytes(total, N);
}

extern "C" MORIG_API int morig_graph_prep(const int64_t *edge_index, int64_t E, int32_t N, int32_t *rowptr, int32_t *col,
                                int32_t *tgt, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(N > 0 && E >= 0, "graph_prep: N=%d E=%lld", N, (long long)E);
    MORIG_CHECK_ARG(E + (int64_t)N < (int64_t)1 << 31, "graph_prep: E+N exceeds int32 range");
    MORIG_CHECK_ARG(rowptr && col && tgt && ws && (E == 0 || edge_index), "graph_prep: null pointer");
    if (ws_bytes < morig_graph_prep_workspace(E, N)) {
        set_error("graph_prep: workspace %zu < %zu", ws_bytes, morig_graph_prep_workspace(E, N));
        return MORIG_E_WORKSPACE;
    }
    const int64_t total = E + N;
    int32_t *keys = (int32_t *)ws, *vals = keys + total;
    void *temp = (void *)(((uintptr_t)(vals + total) + 255) & ~(uintptr_t)255);
    size_t temp_bytes = sort_temp_bytes(total, N);
    const int T = 256;
    MORIG_CUDA(launch_pdl(gp_keys_kernel, dim3((unsigned)ceil_div64(total, T)), dim3(T), 0, stream, edge_index, E, N, keys, vals));
    MORIG_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const int32_t *)keys, tgt, (const int32_t *)vals, col, (int)total, 0,
                                               key_bits(N), stream));
    MORIG_CUDA(launch_pdl(gp_rowptr_kernel, dim3(ceil_div(N + 1, T)), dim3(T), 0, stream, (const int32_t *)tgt, (int32_t)total, N, rowptr));
    return 0;
}

extern "C" MORIG_API int morig_knn_graph(const float *pos, const int32_t *gptr, int32_t n_graphs, int32_t N, int32_t k,
                               int64_t *edge_index, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(pos && gptr && edge_index && N > 0 && n_graphs > 0, "knn_graph: bad argument");
    MORIG_CHECK_ARG(k >= 1 && k <= 16, "knn_graph: k=%d unsupported (1..16)", k);
    const int T = 256;
    knn_kernel<16><<<ceil_div(N * 32, T), T, 0, stream>>>(pos, gptr, n_graphs, N, k, edge_index);
    MORIG_LAUNCH_CHECK("knn_kernel");
    return 0;
}
