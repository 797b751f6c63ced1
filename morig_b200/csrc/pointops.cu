// Point-cloud primitives of the upstream flow producer (models/corrnet.py, models/deformnet.py,
// PointNet++ modules models/basic_modules.py:66-138) and of the surface-sampling front-end of the geodesic graph build
// (section 8(f) #2, data_proc/common_ops.py:175-181).  The reference reaches these through torch_cluster / PyG
// (fps, radius, knn, knn_interpolate); semantics restated in oracle/pointops_port.py.
//
//   fps            farthest point sampling inside each batch segment: one CTA per segment, min-distance array in
//                  registers, one block-wide arg-max per sample (ties -> lower index)
//   ball_query     for every centre the first K points (index order) of its segment with |x - y|^2 < r^2
//   knn_topk       k <= 8 nearest points of the same segment by squared Euclidean distance or by cosine similarity
//                  (ties -> lower index): warp per query, per-lane top-k in registers, warp merge
//   knn_interpolate  inverse-squared-distance interpolation of features from the k neighbours
//   sample_surface   area-weighted uniform samples on a triangle mesh (counter-based hash RNG: reproducible) + face normals
//   edge_mlp_layer   one Linear -> ReLU -> affine layer on E edge rows of a bipartite neighbourhood graph: A rows plain or
//                    gathered as relu(P[tgt] + Q[col]); output stored per edge or max-reduced per target (PointConv)
#include "gemm_simt.cuh"

namespace morig {

constexpr int FPS_THREADS = 1024, FPS_PER_THREAD = 16;        // segments of up to 16384 points

__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float *__restrict__ pos, const int32_t *__restrict__ ptr,
                                                          const int32_t *__restrict__ out_ptr, const int32_t *__restrict__ start,
                                                          int32_t *__restrict__ out) {
    __shared__ float s_val[32];
    __shared__ int s_idx[32];
    __shared__ float s_cur[3];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int lo = ptr[b], n = ptr[b + 1] - lo;
    const int m = out_ptr[b + 1] - out_ptr[b];
    int32_t *dst = out + out_ptr[b];
    float px[FPS_PER_THREAD], py[FPS_PER_THREAD], pz[FPS_PER_THREAD], dmin[FPS_PER_THREAD];
#pragma unroll
    for (int q = 0; q < FPS_PER_THREAD; ++q) {
        const int i = t + q * FPS_THREADS;
        const bool ok = i < n;
        px[q] = ok ? pos[3 * (size_t)(lo + i)] : 0.f;
        py[q] = ok ? pos[3 * (size_t)(lo + i) + 1] : 0.f;
        pz[q] = ok ? pos[3 * (size_t)(lo + i) + 2] : 0.f;
        dmin[q] = ok ? __int_as_float(0x7f800000) : -1.f;       // padding never wins
    }
    int cur = start ? start[b] : 0;
    for (int s = 0; s < m; ++s) {
        if (t == 0) {
            dst[s] = lo + cur;
            s_cur[0] = pos[3 * (size_t)(lo + cur)]; s_cur[1] = pos[3 * (size_t)(lo + cur) + 1]; s_cur[2] = pos[3 * (size_t)(lo + cur) + 2];
        }
        __syncthreads();
        const float cx = s_cur[0], cy = s_cur[1], cz = s_cur[2];
        float best = -2.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int q = 0; q < FPS_PER_THREAD; ++q) {
            const float dx = px[q] - cx, dy = py[q] - cy, dz = pz[q] - cz;
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            dmin[q] = fminf(dmin[q], d);
            const int i = t + q * FPS_THREADS;
            if (dmin[q] > best) { best = dmin[q]; bi = i; }      // ascending i inside a thread: strict > keeps the lowest
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane]; bi = s_idx[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) s_idx[0] = bi;
        }
        __syncthreads();
        cur = s_idx[0];
        __syncthreads();
    }
}

// warp per centre: scan the segment in index order, keep the first K hits (ballot-ordered compaction)
__global__ void __launch_bounds__(256) ball_query_kernel(const float *__restrict__ x, const int32_t *__restrict__ x_ptr,
                                                         const float *__restrict__ y, const int32_t *__restrict__ y_batch, int M,
                                                         float r2, int K, int32_t *__restrict__ nbr, int32_t *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= M) return;
    const int b = y_batch[i];
    const int lo = x_ptr[b], hi = x_ptr[b + 1];
    const float cx = y[3 * (size_t)i], cy = y[3 * (size_t)i + 1], cz = y[3 * (size_t)i + 2];
    int found = 0;
    for (int j0 = lo; j0 < hi && found < K; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        if (j < hi) {
            const float dx = x[3 * (size_t)j] - cx, dy = x[3 * (size_t)j + 1] - cy, dz = x[3 * (size_t)j + 2] - cz;
            hit = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < r2;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        const int pos = found + __popc(mask & ((1u << lane) - 1u));
        if (hit && pos < K) nbr[(size_t)i * K + pos] = j;
        found += __popc(mask);
    }
    if (lane == 0) count[i] = found < K ? found : K;
}

// k nearest of every query among the points of its segment.  metric 0: squared Euclidean distance (smaller is nearer);
// metric 1: cosine similarity x.y / (|x| |y|) (larger is nearer).  Keys are ordered (score, index): ties -> lower index.
constexpr int KNN_K = 8;

__global__ void __launch_bounds__(256) knn_topk_kernel(const float *__restrict__ x, int ldx, const int32_t *__restrict__ x_ptr,
                                                       const float *__restrict__ y, int ldy, const int32_t *__restrict__ y_batch,
                                                       int M, int D, int k, int metric, int32_t *__restrict__ nbr,
                                                       float *__restrict__ score) {
    extern __shared__ float s_q[];                              // [warps][D] query rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= M) return;
    float *q = s_q + (size_t)warp * D;
    float qn = 0.f;
    for (int d = lane; d < D; d += 32) { const float v = y[(size_t)i * ldy + d]; q[d] = v; qn = fmaf(v, v, qn); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) qn += __shfl_xor_sync(0xffffffffu, qn, off);
    __syncwarp();
    const float q_inv = metric ? 1.f / fmaxf(sqrtf(qn), 1e-30f) : 1.f;
    const int b = y_batch[i];
    const int lo = x_ptr[b], hi = x_ptr[b + 1];
    float bs[KNN_K];
    int bi[KNN_K];
#pragma unroll
    for (int s = 0; s < KNN_K; ++s) { bs[s] = __int_as_float(0x7f800000); bi[s] = 0x7fffffff; }
    for (int j = lo + lane; j < hi; j += 32) {
        const float *p = x + (size_t)j * ldx;
        float sc;
        if (metric == 0) {
            float acc = 0.f;
            for (int d = 0; d < D; ++d) { const float df = p[d] - q[d]; acc = __fadd_rn(acc, __fmul_rn(df, df)); }
            sc = acc;
        } else {
            float dot = 0.f, nn = 0.f;
            for (int d = 0; d < D; ++d) { dot = fmaf(p[d], q[d], dot); nn = fmaf(p[d], p[d], nn); }
            sc = -(dot * q_inv / fmaxf(sqrtf(nn), 1e-30f));     // negated: smaller is nearer, like the distance
        }
        if (sc < bs[KNN_K - 1] || (sc == bs[KNN_K - 1] && j < bi[KNN_K - 1])) {
            bs[KNN_K - 1] = sc; bi[KNN_K - 1] = j;
#pragma unroll
            for (int s = KNN_K - 1; s > 0; --s) {
                const bool sw = bs[s] < bs[s - 1] || (bs[s] == bs[s - 1] && bi[s] < bi[s - 1]);
                if (sw) {
                    const float ts = bs[s]; bs[s] = bs[s - 1]; bs[s - 1] = ts;
                    const int ti = bi[s]; bi[s] = bi[s - 1]; bi[s - 1] = ti;
                }
            }
        }
    }
    int head = 0;
    for (int r = 0; r < k; ++r) {
        float sc = __int_as_float(0x7f800000);
        int id = 0x7fffffff;
#pragma unroll
        for (int s = 0; s < KNN_K; ++s)
            if (s == head) { sc = bs[s]; id = bi[s]; }
        float ms = sc;
        int mi = id;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, ms, off);
            const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
            if (os < ms || (os == ms && oi < mi)) { ms = os; mi = oi; }
        }
        if (mi == id && ms == sc && id != 0x7fffffff) ++head;
        if (lane == 0) {
            nbr[(size_t)i * k + r] = mi == 0x7fffffff ? -1 : mi;
            if (score) score[(size_t)i * k + r] = metric ? -ms : ms;
        }
    }
}

// PyG knn_interpolate: w = 1 / max(|pos_x[n] - pos_y[i]|^2, 1e-16);  out[i] = sum_k w_k f[n_k] / sum_k w_k
__global__ void __launch_bounds__(256) knn_interpolate_kernel(const float *__restrict__ f, int ldf, const float *__restrict__ pos_x,
                                                              const float *__restrict__ pos_y, const int32_t *__restrict__ nbr,
                                                              int M, int k, int C, float *__restrict__ out, int ldo) {
    const int64_t total = (int64_t)M * C;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int i = (int)(idx / C);
        float num = 0.f, den = 0.f;
        for (int r = 0; r < k; ++r) {
            const int j = nbr[(size_t)i * k + r];
            if (j < 0) continue;
            const float dx = pos_x[3 * (size_t)j] - pos_y[3 * (size_t)i], dy = pos_x[3 * (size_t)j + 1] - pos_y[3 * (size_t)i + 1],
                        dz = pos_x[3 * (size_t)j + 2] - pos_y[3 * (size_t)i + 2];
            const float w = 1.f / fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)), 1e-16f);
            num = fmaf(w, f[(size_t)j * ldf + c], num);
            den += w;
        }
        out[(size_t)i * ldo + c] = num / den;
    }
}

// ---- area-weighted surface samples (fp64 like the reference's numpy / open3d pipeline) ----------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t bits) { return (double)(bits >> 11) * (1.0 / 9007199254740992.0); }

// cdf [F] = inclusive prefix sums of the triangle areas (computed by the caller); sample s picks the triangle by binary
// search on u * total and a uniform point inside it (square-root parametrisation); normal = unit face normal
__global__ void __launch_bounds__(256) sample_surface_kernel(const double *__restrict__ verts, const int64_t *__restrict__ faces,
                                                             const double *__restrict__ cdf, int F, int S, uint64_t seed,
                                                             double *__restrict__ pts, double *__restrict__ normals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const uint64_t h0 = splitmix64(seed ^ (0x1000003ull * (uint64_t)(s + 1)));
    const uint64_t h1 = splitmix64(h0), h2 = splitmix64(h1);
    const double target = u01(h0) * cdf[F - 1];
    int lo = 0, hi = F - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= target) lo = mid + 1; else hi = mid;
    }
    const int64_t *t = faces + 3 * (size_t)lo;
    const double *a = verts + 3 * t[0], *b = verts + 3 * t[1], *c = verts + 3 * t[2];
    const double r1 = sqrt(u01(h1)), r2 = u01(h2);
    const double wa = 1.0 - r1, wb = r1 * (1.0 - r2), wc = r1 * r2;
    for (int d = 0; d < 3; ++d) pts[3 * (size_t)s + d] = wa * a[d] + wb * b[d] + wc * c[d];
    const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    const double nn = sqrt(nx * nx + ny * ny + nz * nz);
    const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
    normals[3 * (size_t)s] = nx * inv; normals[3 * (size_t)s + 1] = ny * inv; normals[3 * (size_t)s + 2] = nz * inv;
}

__global__ void tri_area_kernel(const double *__restrict__ verts, const int64_t *__restrict__ faces, int F, double *__restrict__ area) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int64_t *t = faces + 3 * (size_t)f;
    const double *a = verts + 3 * t[0], *b = verts + 3 * t[1], *c = verts + 3 * t[2];
    const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    area[f] = 0.5 * sqrt(nx * nx + ny * ny + nz * nz);
}

// single-CTA inclusive scan in fp64 (a mesh has tens of thousands of faces; runs once per mesh)
__global__ void __launch_bounds__(1024) scan_f64_kernel(const double *__restrict__ x, int n, double *__restrict__ out) {
    __shared__ double part[1024];
    const int t = threadIdx.x, per = (n + 1023) / 1024;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    double s = 0.0;
    for (int i = lo; i < hi; ++i) s += x[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) { double run = 0.0; for (int i = 0; i < 1024; ++i) { const double v = part[i]; part[i] = run; run += v; } }
    __syncthreads();
    double run = part[t];
    for (int i = lo; i < hi; ++i) { run += x[i]; out[i] = run; }
}

template <int BN, int AMODE, int EPI>
static int launch_edge_layer(const GemmP &p, int rows, int frames, cudaStream_t stream) {
    auto kern = gemm_simt_kernel<128, BN, AMODE, EPI>;
    constexpr size_t smem = gemm_smem_bytes<128, BN, EPI>();
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev = dev;
    }
    kern<<<dim3(ceil_div(rows, 128), ceil_div(p.N, BN), frames), GEMM_THREADS, smem, stream>>>(p);
    MORIG_LAUNCH_CHECK("edge_mlp_layer");
    return 0;
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API int morig_fps(const float *pos, const int32_t *ptr, const int32_t *out_ptr, const int32_t *start, int32_t B,
                                   int32_t max_segment, int32_t *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(pos && ptr && out_ptr && out && B > 0, "fps: bad argument");
    MORIG_CHECK_ARG(max_segment <= FPS_THREADS * FPS_PER_THREAD, "fps: segment of %d points exceeds %d", max_segment,
                    FPS_THREADS * FPS_PER_THREAD);
    fps_kernel<<<B, FPS_THREADS, 0, stream>>>(pos, ptr, out_ptr, start, out);
    MORIG_LAUNCH_CHECK("fps_kernel");
    return 0;
}

extern "C" MORIG_API int morig_ball_query(const float *x, const int32_t *x_ptr, const float *y, const int32_t *y_batch, int32_t M,
                                          float radius, int32_t K, int32_t *nbr, int32_t *count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && x_ptr && y && y_batch && nbr && count && M > 0 && K > 0 && radius > 0.f, "ball_query: bad argument");
    ball_query_kernel<<<ceil_div(M * 32, 256), 256, 0, stream>>>(x, x_ptr, y, y_batch, M, radius * radius, K, nbr, count);
    MORIG_LAUNCH_CHECK("ball_query_kernel");
    return 0;
}

extern "C" MORIG_API int morig_knn_topk(const float *x, int32_t ldx, const int32_t *x_ptr, const float *y, int32_t ldy,
                                        const int32_t *y_batch, int32_t M, int32_t D, int32_t k, int32_t metric, int32_t *nbr,
                                        float *score, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && x_ptr && y && y_batch && nbr && M > 0 && D > 0 && D <= 1024, "knn_topk: bad argument");
    MORIG_CHECK_ARG(k >= 1 && k <= KNN_K && (metric == 0 || metric == 1), "knn_topk: k=%d (1..%d), metric=%d", k, KNN_K, metric);
    knn_topk_kernel<<<ceil_div(M * 32, 256), 256, (size_t)8 * D * sizeof(float), stream>>>(x, ldx, x_ptr, y, ldy, y_batch, M, D, k, metric,
                                                                                         nbr, score);
    MORIG_LAUNCH_CHECK("knn_topk_kernel");
    return 0;
}

extern "C" MORIG_API int morig_knn_interpolate(const float *f, int32_t ldf, const float *pos_x, const float *pos_y, const int32_t *nbr,
                                               int32_t M, int32_t k, int32_t C, float *out, int32_t ldo, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(f && pos_x && pos_y && nbr && out && M > 0 && k > 0 && C > 0, "knn_interpolate: bad argument");
    const int64_t blocks = ceil_div64((int64_t)M * C, 256), cap = (int64_t)sm_count() * 8;
    knn_interpolate_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(f, ldf, pos_x, pos_y, nbr, M, k, C, out, ldo);
    MORIG_LAUNCH_CHECK("knn_interpolate_kernel");
    return 0;
}

extern "C" MORIG_API int morig_sample_surface(const double *verts, const int64_t *faces, int32_t F, int32_t S, uint64_t seed,
                                              double *pts, double *normals, double *ws /* 2 F doubles */, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(verts && faces && pts && normals && ws && F > 0 && S > 0, "sample_surface: bad argument");
    tri_area_kernel<<<ceil_div(F, 256), 256, 0, stream>>>(verts, faces, F, ws);
    scan_f64_kernel<<<1, 1024, 0, stream>>>(ws, F, ws + F);
    sample_surface_kernel<<<ceil_div(S, 256), 256, 0, stream>>>(verts, faces, ws + F, F, S, seed, pts, normals);
    MORIG_LAUNCH_CHECK("sample_surface_kernel");
    return 0;
}

extern "C" MORIG_API int morig_edge_mlp_layer(const morig_edge_layer_desc *d, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(d && d->W && d->E > 0 && d->N > 0 && d->K > 0, "edge_mlp_layer: bad argument");
    MORIG_CHECK_ARG(d->ldw % 4 == 0 && d->ldw >= d->N, "edge_mlp_layer: ldw");
    MORIG_CHECK_ARG((d->A != nullptr) != (d->P != nullptr), "edge_mlp_layer: exactly one of A and (P, Q)");
    MORIG_CHECK_ARG((d->C != nullptr) != (d->out != nullptr), "edge_mlp_layer: exactly one of C and out");
    MORIG_CHECK_ARG(d->rowptr && d->tgt && (d->A || d->col), "edge_mlp_layer: graph arrays");
    GemmP p{};
    p.A = d->A; p.lda = d->lda;
    p.a_vec = (d->A && d->lda % 4 == 0 && d->K % 4 == 0 && (reinterpret_cast<uintptr_t>(d->A) & 15u) == 0) ? 1 : 0;
    p.P = d->P; p.Q = d->Q; p.ldpq = d->ldpq;
    p.rowptr = d->rowptr; p.col = d->col; p.tgt = d->tgt; p.n_vtx_frame = d->n_targets;
    p.W = d->W; p.ldw = d->ldw; p.bias = d->bias; p.scale = d->scale; p.shift = d->shift;
    p.M = d->E; p.N = d->N; p.K = d->K; p.relu = 1;
    p.n_vtx = d->E; p.n_graphs = 1;
    if (d->P) MORIG_CHECK_ARG(d->ldpq % 4 == 0 && d->K % 4 == 0 && (reinterpret_cast<uintptr_t>(d->P) & 15u) == 0 &&
                              (reinterpret_cast<uintptr_t>(d->Q) & 15u) == 0, "edge_mlp_layer: P / Q alignment");
    if (d->C) { p.C = d->C; p.ldc = d->ldc; p.c_vec = (d->ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(d->C) & 15u) == 0) ? 1 : 0; }
    else { p.C = d->out; p.ldc = d->ldo; }
    const bool wide = d->N > 64;
    if (d->P && d->C) return wide ? launch_edge_layer<128, AMODE_GATHER, EPI_STORE>(p, d->E, 1, stream)
                                  : launch_edge_layer<64, AMODE_GATHER, EPI_STORE>(p, d->E, 1, stream);
    if (d->P) return wide ? launch_edge_layer<128, AMODE_GATHER, EPI_SEGMAX>(p, d->E, 1, stream)
                          : launch_edge_layer<64, AMODE_GATHER, EPI_SEGMAX>(p, d->E, 1, stream);
    if (d->C) return wide ? launch_edge_layer<128, AMODE_PLAIN, EPI_STORE>(p, d->E, 1, stream)
                          : launch_edge_layer<64, AMODE_PLAIN, EPI_STORE>(p, d->E, 1, stream);
    return wide ? launch_edge_layer<128, AMODE_PLAIN, EPI_SEGMAX>(p, d->E, 1, stream)
                : launch_edge_layer<64, AMODE_PLAIN, EPI_SEGMAX>(p, d->E, 1, stream);
}
