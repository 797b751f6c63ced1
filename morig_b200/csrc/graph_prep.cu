// Graph preparation: int64 [2,E] edge list -> self-loop-normalised, target-sorted CSR (int32).
// Replaces remove_self_loops/add_self_loops of models/basic_modules.py:188-189 (run 36x per forward
// by the reference) with one pass per edge set; also a brute-force kNN graph builder used for the
// synthetic geodesic stand-in.  Integer work: results are bit-exact against oracle/graph_port.py.
#include "common.cuh"

namespace morig {

__global__ void gp_init_kernel(int32_t *cnt, int32_t n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = 1;  // the appended self loop
}

__global__ void gp_count_kernel(const int64_t *__restrict__ ei, int64_t E, int32_t n, int32_t *cnt) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t j = ei[e], i = ei[E + e];
    if (j == i || i < 0 || i >= n || j < 0 || j >= n) return;
    atomicAdd(&cnt[(int)i], 1);
}

// single-block exclusive scan; n+1 outputs. cursor[i] = rowptr[i].
__global__ void gp_scan_kernel(const int32_t *__restrict__ cnt, int32_t n, int32_t *rowptr, int32_t *cursor) {
    __shared__ int32_t part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    int32_t s = 0;
    for (int i = lo; i < hi; ++i) s += cnt[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < nt; off <<= 1) {  // Hillis-Steele inclusive scan of the partials
        int32_t v = (t >= off) ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int32_t run = part[t] - s;
    for (int i = lo; i < hi; ++i) {
        rowptr[i] = run;
        cursor[i] = run;
        run += cnt[i];
    }
    if (t == nt - 1) rowptr[n] = part[t];
}

__global__ void gp_fill_kernel(const int64_t *__restrict__ ei, int64_t E, int32_t n, int32_t *cursor,
                               int32_t *col, int32_t *eid) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t j = ei[e], i = ei[E + e];
    if (j == i || i < 0 || i >= n || j < 0 || j >= n) return;
    int slot = atomicAdd(&cursor[(int)i], 1);
    col[slot] = (int32_t)j;
    eid[slot] = (int32_t)e;
}

// one thread per target: restore input order inside the segment (stable CSR), put the self loop
// last, expand the target id per slot.
__global__ void gp_finalize_kernel(const int32_t *__restrict__ rowptr, int32_t n, int32_t *col, int32_t *eid,
                                   int32_t *tgt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int lo = rowptr[i], hi = rowptr[i + 1] - 1;  // [lo, hi) real edges, slot hi = self loop
    for (int a = lo + 1; a < hi; ++a) {                // insertion sort by original edge id
        int32_t ke = eid[a], kc = col[a];
        int b = a - 1;
        while (b >= lo && eid[b] > ke) {
            eid[b + 1] = eid[b];
            col[b + 1] = col[b];
            --b;
        }
        eid[b + 1] = ke;
        col[b + 1] = kc;
    }
    col[hi] = i;
    for (int a = lo; a <= hi; ++a) tgt[a] = i;
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
    return sizeof(int32_t) * (size_t)(2 * (int64_t)N + E + N) + 256;
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
    int32_t *cnt = (int32_t *)ws, *cursor = cnt + N, *eid = cursor + N;
    const int T = 256;
    gp_init_kernel<<<ceil_div(N, T), T, 0, stream>>>(cnt, N);
    if (E > 0) gp_count_kernel<<<(unsigned)ceil_div64(E, T), T, 0, stream>>>(edge_index, E, N, cnt);
    gp_scan_kernel<<<1, 1024, 0, stream>>>(cnt, N, rowptr, cursor);
    if (E > 0) gp_fill_kernel<<<(unsigned)ceil_div64(E, T), T, 0, stream>>>(edge_index, E, N, cursor, col, eid);
    gp_finalize_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(rowptr, N, col, eid, tgt);
    MORIG_LAUNCH_CHECK("graph_prep");
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
