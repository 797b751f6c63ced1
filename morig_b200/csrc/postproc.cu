// Rest of the joint-extraction post-process and the two losses next to it:
//
//   nms_meanshift   utils/cluster_utils.py:38-63 -- non-maximum suppression over the mean-shift modes: points are visited by
//                   decreasing neighbour count; a visited live point clears its bandwidth-ball and survives iff its ball
//                   holds an attention > thrd_attn or more than thrd_density * N points.
//                   The ball statistics do not depend on the visiting order, so they (and a bit matrix of the balls) are
//                   computed in parallel; only the greedy sweep is sequential -- one CTA, one 64-bit word of the live set
//                   per thread.  fp64 distances, evaluated like numpy (sqrt(((dx^2 + dy^2) + dz^2)) <= bandwidth).
//                   Tie rule (the reference's np.argsort()[::-1] is an unstable sort, i.e. unspecified): equal counts are
//                   visited from the higher index down, = np.argsort(kind="stable")[::-1] (oracle/cluster_port.py).
//   chamfer         models/customized_losses.py:231-251 (`chamfer_distance_with_average`, fp32, with gradient) and
//                   utils/mst_utils.py:316-321 (`chamfer_dist`, fp64 numpy): nearest-neighbour distances both ways
//   info_nce        row-wise cross entropy of anchor . key^T / tau against a label column (models/customized_losses.py:107-135
//                   `infoNCE` per direction; :137-158 `multi_pos_infoNCE` after its sampling): streaming log-sum-exp, no
//                   [R, M] logit matrix in memory
#include "common.cuh"

namespace morig {

// ---- NMS ------------------------------------------------------------------------------------------------------------------
// block = 32 centres x 8 slices of the other points; per centre: neighbour count, max attention in the ball, ball bit row
__global__ void __launch_bounds__(256) nms_ball_kernel(const double *__restrict__ pts, const double *__restrict__ attn,
                                                       double bw, int N, int words, int32_t *__restrict__ count,
                                                       double *__restrict__ amax, unsigned long long *__restrict__ bits) {
    __shared__ int s_cnt[8][32];
    __shared__ double s_max[8][32];
    const int il = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + il;
    const bool ok = i < N;
    const double px = ok ? pts[3 * i] : 0.0, py = ok ? pts[3 * i + 1] : 0.0, pz = ok ? pts[3 * i + 2] : 0.0;
    int cnt = 0;
    double mx = -1.0e300;
    // slice sl owns the words sl, sl + 8, ...: every (centre, word) is written by exactly one thread
    for (int w = sl; w < words; w += 8) {
        unsigned long long word = 0ull;
        for (int b = 0; b < 64; ++b) {
            const int j = w * 64 + b;
            if (j >= N) break;
            const double dx = __dsub_rn(pts[3 * j], px), dy = __dsub_rn(pts[3 * j + 1], py), dz = __dsub_rn(pts[3 * j + 2], pz);
            const double y = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            if (sqrt(y) <= bw) {
                word |= 1ull << b;
                ++cnt;
                mx = fmax(mx, attn[j]);
            }
        }
        if (ok) bits[(size_t)i * words + w] = word;
    }
    s_cnt[sl][il] = cnt; s_max[sl][il] = mx;
    __syncthreads();
    if (sl == 0 && ok) {
        for (int s = 1; s < 8; ++s) { cnt += s_cnt[s][il]; mx = fmax(mx, s_max[s][il]); }
        count[i] = cnt;
        amax[i] = mx;
    }
}

// visiting order: decreasing count, equal counts from the higher index down; rank by counting (N is a few thousand)
__global__ void __launch_bounds__(256) nms_rank_kernel(const int32_t *__restrict__ count, int N, int32_t *__restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int ci = count[i];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
        const int cj = count[j];
        rank += (cj > ci) || (cj == ci && j > i);
    }
    order[rank] = i;
}

// the greedy sweep: thread t owns word t (+ blockDim.x, ...) of the live set
__global__ void __launch_bounds__(1024) nms_sweep_kernel(const int32_t *__restrict__ order, const int32_t *__restrict__ count,
                                                         const double *__restrict__ amax,
                                                         const unsigned long long *__restrict__ bits, int N, int words,
                                                         double thrd_density, double thrd_attn, uint8_t *__restrict__ keep) {
    extern __shared__ unsigned long long live[];                // [words]
    __shared__ int s_alive;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const int left = N - w * 64;
        live[w] = left >= 64 ? ~0ull : (left > 0 ? ((1ull << left) - 1ull) : 0ull);
    }
    __syncthreads();
    for (int k = 0; k < N; ++k) {
        const int i = order[k];
        if (threadIdx.x == 0) s_alive = (int)((live[i >> 6] >> (i & 63)) & 1ull);
        __syncthreads();
        if (s_alive) {                                          // uniform branch
            for (int w = threadIdx.x; w < words; w += blockDim.x) live[w] &= ~bits[(size_t)i * words + w];
            __syncthreads();
            const bool survive = amax[i] > thrd_attn || (double)count[i] / (double)N > thrd_density;
            if (threadIdx.x == 0 && survive) live[i >> 6] |= 1ull << (i & 63);
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) keep[j] = (uint8_t)((live[j >> 6] >> (j & 63)) & 1ull);
}

// ---- chamfer: nearest-neighbour distance of every row of A [N, D] among the rows of B [M, D] -------------------------------
// block = 32 rows of A x 8 slices of B; ties -> lower index of B (torch.min / np.min return the first minimum)
template <typename T>
__global__ void __launch_bounds__(256) nn_dist_kernel(const T *__restrict__ A, int N, const T *__restrict__ B, int M, int D,
                                                      T *__restrict__ dist, int32_t *__restrict__ arg) {
    __shared__ T s_d[8][32];
    __shared__ int s_i[8][32];
    const int il = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + il;
    T a[4] = {0, 0, 0, 0};
    if (i < N)
        for (int d = 0; d < D; ++d) a[d] = A[(size_t)i * D + d];
    T best = (T)1e30;
    int bi = -1;
    for (int j = sl; j < M; j += 8) {
        T s = 0;
        for (int d = 0; d < D; ++d) { const T df = a[d] - B[(size_t)j * D + d]; s += df * df; }
        if (bi < 0 || s < best) { best = s; bi = j; }
    }
    s_d[sl][il] = best; s_i[sl][il] = bi;
    __syncthreads();
    if (sl == 0 && i < N) {
        for (int s = 1; s < 8; ++s) {
            const T v = s_d[s][il];
            const int j = s_i[s][il];
            if (j >= 0 && (bi < 0 || v < best || (v == best && j < bi))) { best = v; bi = j; }
        }
        dist[i] = sqrt(best);
        if (arg) arg[i] = bi;
    }
}

// gradient of  L = g1 * sum_i |a_i - b_{n1(i)}|  +  g2 * sum_j |b_j - a_{n2(j)}|  with respect to A (dA zeroed by the caller):
// the first term is local to row i; the second scatters onto the nearest rows (fp32 atomics)
__global__ void __launch_bounds__(256) chamfer_bwd_kernel(const float *__restrict__ A, int N, const float *__restrict__ B, int M,
                                                          int D, const float *__restrict__ d1, const int32_t *__restrict__ n1,
                                                          const float *__restrict__ d2, const int32_t *__restrict__ n2,
                                                          float g1, float g2, float *__restrict__ dA) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < N) {
        const int j = n1[t];
        const float inv = d1[t] > 0.f ? g1 / d1[t] : 0.f;
        for (int d = 0; d < D; ++d) atomicAdd(dA + (size_t)t * D + d, (A[(size_t)t * D + d] - B[(size_t)j * D + d]) * inv);
    } else if (t < N + M) {
        const int j = t - N, i = n2[j];
        const float inv = d2[j] > 0.f ? g2 / d2[j] : 0.f;
        for (int d = 0; d < D; ++d) atomicAdd(dA + (size_t)i * D + d, (A[(size_t)i * D + d] - B[(size_t)j * D + d]) * inv);
    }
}

// ---- infoNCE rows: loss[r] = logsumexp_m(a_r . k_m / tau) - a_r . k_{label[r]} / tau ------------------------------------------
// warp per anchor row; lanes stride over the keys with an online (max, sum) pair, merged by shuffles
// sel != NULL: row r only sees the S keys K[sel[r, s]] (its candidate list), and label[r] is a position in that list
__global__ void __launch_bounds__(256) info_nce_fwd_kernel(const float *__restrict__ A, int lda, const float *__restrict__ K, int ldk,
                                                           const int64_t *__restrict__ label, const int64_t *__restrict__ sel,
                                                           int S, int R, int M, int C, float inv_tau,
                                                           float *__restrict__ loss, float *__restrict__ lse) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const float *a = A + (size_t)r * lda;
    float mx = neg_inf(), sum = 0.f;
    if (sel) M = S;
    for (int m = lane; m < M; m += 32) {
        const float *k = K + (size_t)(sel ? sel[(size_t)r * S + m] : m) * ldk;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(a[c], k[c], s);
        s *= inv_tau;
        if (s > mx) { sum = sum * expf(mx - s) + 1.f; mx = s; }
        else sum += expf(s - mx);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, off), os = __shfl_xor_sync(0xffffffffu, sum, off);
        const float nm = fmaxf(mx, om);
        sum = (mx > neg_inf() ? sum * expf(mx - nm) : 0.f) + (om > neg_inf() ? os * expf(om - nm) : 0.f);
        mx = nm;
    }
    if (lane == 0) {
        const float *k = K + (size_t)(sel ? sel[(size_t)r * S + label[r]] : label[r]) * ldk;
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(a[c], k[c], s);
        const float l = mx + logf(sum);
        lse[r] = l;
        loss[r] = l - s * inv_tau;
    }
}

// dA[r] = g[r] / tau * ( sum_m p_rm k_m - k_label ),  dK[m] += g[r] / tau * (p_rm - [m = label]) a_r   (dK zeroed by the caller)
// warp per anchor row, lane = channel (C <= 128): the keys are walked in order, so dA has a fixed summation order; dK
// collects contributions of different rows with fp32 atomics
constexpr int NCE_CPL = 4;

__global__ void __launch_bounds__(256) info_nce_bwd_kernel(const float *__restrict__ A, int lda, const float *__restrict__ K, int ldk,
                                                           const int64_t *__restrict__ label, const int64_t *__restrict__ sel,
                                                           int S, const float *__restrict__ lse,
                                                           const float *__restrict__ g, int R, int M, int C, float inv_tau,
                                                           float *__restrict__ dA, int ldda, float *__restrict__ dK, int lddk) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    if (sel) M = S;
    float a[NCE_CPL], acc[NCE_CPL];
#pragma unroll
    for (int q = 0; q < NCE_CPL; ++q) {
        const int c = lane + 32 * q;
        a[q] = c < C ? A[(size_t)r * lda + c] : 0.f;
        acc[q] = 0.f;
    }
    const float gr = g[r] * inv_tau, l = lse[r];
    const int lab = (int)label[r];
    for (int m = 0; m < M; ++m) {
        float kv[NCE_CPL];
        float s = 0.f;
        const size_t row = sel ? (size_t)sel[(size_t)r * S + m] : (size_t)m;
#pragma unroll
        for (int q = 0; q < NCE_CPL; ++q) {
            const int c = lane + 32 * q;
            kv[q] = c < C ? K[row * ldk + c] : 0.f;
            s = fmaf(a[q], kv[q], s);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const float w = gr * (expf(s * inv_tau - l) - (m == lab ? 1.f : 0.f));
#pragma unroll
        for (int q = 0; q < NCE_CPL; ++q) {
            const int c = lane + 32 * q;
            acc[q] = fmaf(w, kv[q], acc[q]);
            if (dK && c < C && w != 0.f) atomicAdd(dK + row * lddk + c, w * a[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < NCE_CPL; ++q) {
        const int c = lane + 32 * q;
        if (c < C) dA[(size_t)r * ldda + c] = acc[q];
    }
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API size_t morig_nms_meanshift_workspace(int32_t N) {
    const size_t words = (size_t)(N + 63) / 64;
    return (size_t)N * words * 8 + (size_t)N * (4 + 8 + 4) + 256;
}

extern "C" MORIG_API int morig_nms_meanshift(const double *pts, const double *attn, int32_t N, double bandwidth, double thrd_density,
                                             double thrd_attn, uint8_t *keep, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(pts && attn && keep && N > 0, "nms_meanshift: bad argument");
    const int words = (N + 63) / 64;
    MORIG_CHECK_ARG((size_t)words * 8 <= 200 * 1024, "nms_meanshift: N=%d too large for the single-CTA sweep", N);
    if (!ws || ws_bytes < morig_nms_meanshift_workspace(N)) { set_error("nms_meanshift: workspace too small"); return MORIG_E_WORKSPACE; }
    unsigned long long *bits = reinterpret_cast<unsigned long long *>(ws);
    double *amax = reinterpret_cast<double *>(bits + (size_t)N * words);
    int32_t *count = reinterpret_cast<int32_t *>(amax + N);
    int32_t *order = count + N;
    nms_ball_kernel<<<ceil_div(N, 32), 256, 0, stream>>>(pts, attn, bandwidth, N, words, count, amax, bits);
    MORIG_LAUNCH_CHECK("nms_ball_kernel");
    nms_rank_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(count, N, order);
    MORIG_LAUNCH_CHECK("nms_rank_kernel");
    const size_t smem = (size_t)words * 8;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured_dev = dev;
    }
    const int threads = words < 32 ? 32 : (words > 1024 ? 1024 : ceil_div(words, 32) * 32);
    nms_sweep_kernel<<<1, threads, smem, stream>>>(order, count, amax, bits, N, words, thrd_density, thrd_attn, keep);
    MORIG_LAUNCH_CHECK("nms_sweep_kernel");
    return 0;
}

extern "C" MORIG_API int morig_nn_dist_f32(const float *A, int32_t N, const float *B, int32_t M, int32_t D, float *dist,
                                           int32_t *arg, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(A && B && dist && N > 0 && M > 0 && D >= 1 && D <= 4, "nn_dist_f32: bad argument (D <= 4)");
    nn_dist_kernel<float><<<ceil_div(N, 32), 256, 0, stream>>>(A, N, B, M, D, dist, arg);
    MORIG_LAUNCH_CHECK("nn_dist_kernel<float>");
    return 0;
}

extern "C" MORIG_API int morig_nn_dist_f64(const double *A, int32_t N, const double *B, int32_t M, int32_t D, double *dist,
                                           int32_t *arg, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(A && B && dist && N > 0 && M > 0 && D >= 1 && D <= 4, "nn_dist_f64: bad argument (D <= 4)");
    nn_dist_kernel<double><<<ceil_div(N, 32), 256, 0, stream>>>(A, N, B, M, D, dist, arg);
    MORIG_LAUNCH_CHECK("nn_dist_kernel<double>");
    return 0;
}

extern "C" MORIG_API int morig_chamfer_bwd_f32(const float *A, int32_t N, const float *B, int32_t M, int32_t D, const float *d1,
                                               const int32_t *n1, const float *d2, const int32_t *n2, float g1, float g2,
                                               float *dA, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(A && B && d1 && n1 && d2 && n2 && dA && N > 0 && M > 0 && D >= 1 && D <= 4, "chamfer_bwd: bad argument");
    MORIG_CUDA(cudaMemsetAsync(dA, 0, (size_t)N * D * sizeof(float), stream));
    chamfer_bwd_kernel<<<ceil_div(N + M, 256), 256, 0, stream>>>(A, N, B, M, D, d1, n1, d2, n2, g1, g2, dA);
    MORIG_LAUNCH_CHECK("chamfer_bwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_info_nce_fwd(const float *A, int32_t lda, const float *K, int32_t ldk, const int64_t *label,
                                            const int64_t *sel, int32_t S, int32_t R, int32_t M, int32_t C, float tau, float *loss,
                                            float *lse, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(A && K && label && loss && lse && R > 0 && M > 0 && C > 0 && tau > 0.f && (!sel || S > 0),
                    "info_nce_fwd: bad argument");
    info_nce_fwd_kernel<<<ceil_div(R * 32, 256), 256, 0, stream>>>(A, lda, K, ldk, label, sel, S, R, M, C, 1.f / tau, loss, lse);
    MORIG_LAUNCH_CHECK("info_nce_fwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_info_nce_bwd(const float *A, int32_t lda, const float *K, int32_t ldk, const int64_t *label,
                                            const int64_t *sel, int32_t S, const float *lse, const float *g, int32_t R, int32_t M,
                                            int32_t C, float tau, float *dA, int32_t ldda, float *dK, int32_t lddk, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(A && K && label && lse && g && dA && R > 0 && M > 0 && C > 0 && C <= 32 * NCE_CPL && tau > 0.f,
                    "info_nce_bwd: bad argument (C <= 128)");
    if (dK) MORIG_CUDA(cudaMemset2DAsync(dK, (size_t)lddk * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)M, stream));
    info_nce_bwd_kernel<<<ceil_div(R * 32, 256), 256, 0, stream>>>(A, lda, K, ldk, label, sel, S, lse, g, R, M, C, 1.f / tau, dA, ldda,
                                                                   dK, lddk);
    MORIG_LAUNCH_CHECK("info_nce_bwd_kernel");
    return 0;
}
