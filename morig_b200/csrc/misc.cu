// Small memory-bound kernels of the path: key-frame cls-token attention, row L2 normalise,
// key-frame mean/max, strided column gather (replaces torch.cat / slicing), fill.
#include "common.cuh"

namespace morig {

// ---- TemporalAttn (models/rignet.py:36-44), cls query row only -------------------------------
// One warp per vertex, lane = input channel (C <= 32).  Per head h:
//   logit_t = u_h . x_t  (t < T),  logit_cls = l0_h;  a = softmax over the T+1 logits
//   y_h = sum_t a_t x_t
//   out = sum_h ( Mv_h y_h + a_cls,h * c0_h )
// u, l0, Mv, c0 are parameter-only products prepared once by the host layer.
constexpr int ATT_MAX_T = 8;
constexpr int ATT_MAX_HEADS = 4;

__global__ void __launch_bounds__(256) temporal_attn_kernel(const float *__restrict__ x, int N, int T, int C, int heads,
                                                            int D, const float *__restrict__ u,
                                                            const float *__restrict__ l0, const float *__restrict__ Mv,
                                                            const float *__restrict__ c0, float *__restrict__ out,
                                                            int ldo, float *out_amax) {
    extern __shared__ float s_mv[];                 // [heads][C][D]  (transposed for lane-contiguous reads)
    pdl_trigger();
    pdl_wait();
    for (int idx = threadIdx.x; idx < heads * C * D; idx += blockDim.x) {
        const int h = idx / (C * D), rem = idx % (C * D), cc = rem / D, dd = rem % D;
        s_mv[idx] = Mv[((size_t)h * D + dd) * C + cc];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    float am = 0.f;
    // persistent: the 16 KB parameter image above is staged once per CTA, warps stride over the vertices
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
    float xt[ATT_MAX_T];
#pragma unroll
    for (int t = 0; t < ATT_MAX_T; ++t) xt[t] = (t < T && lane < C) ? x[((size_t)n * T + t) * C + lane] : 0.f;

    float y[ATT_MAX_HEADS], a_cls[ATT_MAX_HEADS];
#pragma unroll
    for (int h = 0; h < ATT_MAX_HEADS; ++h) {
        y[h] = 0.f; a_cls[h] = 0.f;
        if (h >= heads) continue;
        const float uh = lane < C ? u[h * C + lane] : 0.f;
        float lg[ATT_MAX_T];
        const float lc = l0[h];
        float mx = lc;
#pragma unroll
        for (int t = 0; t < ATT_MAX_T; ++t) {
            float v = uh * xt[t];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            lg[t] = v;
            if (t < T) mx = fmaxf(mx, v);
        }
        float den = expf(lc - mx);
        const float e_cls = den;
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < ATT_MAX_T; ++t) {
            if (t < T) {
                const float e = expf(lg[t] - mx);
                den += e;
                acc = fmaf(e, xt[t], acc);
            }
        }
        y[h] = acc / den;
        a_cls[h] = e_cls / den;
    }
    for (int d = lane; d < D; d += 32) {
        float o = 0.f;
#pragma unroll
        for (int h = 0; h < ATT_MAX_HEADS; ++h) {
            if (h >= heads) continue;
            o = fmaf(a_cls[h], c0[h * D + d], o);
            for (int cc = 0; cc < C; ++cc)
                o = fmaf(__shfl_sync(0xffffffffu, y[h], cc), s_mv[(h * C + cc) * D + d], o);
        }
        out[(size_t)n * ldo + d] = o;
        am = fmaxf(am, fabsf(o));
    }
    }
    amax_commit(out_amax, am);
}

// ---- F.normalize(dim=1): warp per row ------------------------------------------------------------
__global__ void __launch_bounds__(256) row_normalize_kernel(float *x, int ldx, int R, int C, float *dst2, int N,
                                                            int n_frames) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    pdl_trigger();
    pdl_wait();
    if (r >= R) return;
    float *row = x + (size_t)r * ldx;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = row[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    float *row2 = nullptr;
    if (dst2) { const int f = r / N, v = r % N; row2 = dst2 + ((size_t)v * n_frames + f) * C; }
    for (int c = lane; c < C; c += 32) {
        const float v = row[c] / denom;
        row[c] = v;
        if (row2) row2[c] = v;
    }
}

__global__ void frame_reduce_kernel(const float *__restrict__ x, int N, int T, int C, int mode, float *out, int ldo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    pdl_trigger();
    pdl_wait();
    if (idx >= (int64_t)N * C) return;
    const int n = (int)(idx / C), c = (int)(idx % C);
    const float *src = x + (size_t)n * T * C + c;
    float acc = src[0];
    for (int t = 1; t < T; ++t) acc = mode == 0 ? acc + src[(size_t)t * C] : fmaxf(acc, src[(size_t)t * C]);
    out[(size_t)n * ldo + c] = mode == 0 ? acc / (float)T : acc;
}

__global__ void gather_cols_kernel(const float *__restrict__ src, int lds, int src_off, int frame_stride,
                                   const int32_t *__restrict__ cols, int C, int N, int n_frames, float *dst, int ldd,
                                   int dst_off, float *dst_amax) {
    // grid-stride over a bounded grid: the operand-range atomic below is per warp, all on one address
    const int64_t total = (int64_t)N * n_frames * C;
    float am = 0.f;
    pdl_trigger();
    pdl_wait();
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % C);
        const int64_t r = idx / C;
        const int f = (int)(r / N), v = (int)(r % N);
        const int sc = src_off + f * frame_stride + (cols ? cols[c] : c);
        const float val = src[(size_t)v * lds + sc];
        dst[(size_t)r * ldd + dst_off + c] = val;
        am = fmaxf(am, fabsf(val));
    }
    amax_commit(dst_amax, am);
}

// grid-stride max |x| of a strided [R, C] block
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ x, int ldx, int R, int C, float *amax) {
    const int64_t total = (int64_t)R * C;
    float am = 0.f;
    pdl_trigger();
    pdl_wait();
    if (((C | ldx) & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0) {       // 16-byte loads
        const int C4 = C >> 2;
        const int64_t total4 = (int64_t)R * C4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = i / C4;
            const float4 v = *reinterpret_cast<const float4 *>(x + (size_t)r * ldx + 4 * (i - r * C4));
            am = fmaxf(fmaxf(am, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
            am = fmaxf(am, fabsf(x[(size_t)(i / C) * ldx + (i % C)]));
    }
    amax_commit_block(amax, am);
}

__global__ void fill_kernel(float *dst, int64_t n, float value) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    pdl_trigger();
    pdl_wait();
    for (; i < n; i += stride) dst[i] = value;
}

// several buffers in one launch (the -inf initialisation of every EdgeConv output and pooled feature of a GCNRig)
constexpr int FILL_MANY_MAX = 8;
struct FillMany { float *dst[FILL_MANY_MAX]; long long n[FILL_MANY_MAX]; int count; float value; };

__global__ void __launch_bounds__(256) fill_many_kernel(const FillMany f) {
    pdl_trigger();
    pdl_wait();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int k = 0; k < f.count; ++k) {
        float4 *d4 = reinterpret_cast<float4 *>(f.dst[k]);
        const int64_t n4 = f.n[k] >> 2;
        const float4 v4 = make_float4(f.value, f.value, f.value, f.value);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) d4[i] = v4;
        for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < f.n[k]; i += stride) f.dst[k][i] = f.value;
    }
}

// Selective start values for the fused EdgeConv outputs: only the vertices whose CSR segment straddles a multiple of 32
// slots are ever merged with the atomic max (every edge kernel cuts its tiles into row ranges that are multiples of 32:
// 32 slots per warp in edge_mma_kernel, 64 / 128 rows per epilogue warp in the tcgen05 kernels, 32 / 64 rows per column
// walker in the CUDA-core fallback); every other vertex is written by exactly one plain store.  So instead of
// streaming -inf over whole [R, 2 (H + Dp)] buffers (304 MB per GCNRig at 4 x 4096 vertices x 5 key-frames) one warp
// per (vertex, branch group) tests the segment and fills `ncols` columns of that vertex in every key-frame.
constexpr int FILL_CUT_MAX = 8;
struct FillCut {
    const int32_t *rowptr[FILL_CUT_MAX];
    float *out[FILL_CUT_MAX];
    int ld[FILL_CUT_MAX], col0[FILL_CUT_MAX], ncols[FILL_CUT_MAX];
    int count, N, frames;
    float value;
};

__global__ void __launch_bounds__(256) fill_cut_kernel(const FillCut f) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = (int64_t)f.N * f.count;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += warps) {
        const int k = (int)(w / f.N), v = (int)(w - (int64_t)k * f.N);
        const int a = f.rowptr[k][v], b = f.rowptr[k][v + 1];
        if (b <= a || (a >> 5) == ((b - 1) >> 5)) continue;              // empty, or inside one 32-slot range: plain store
        const int nc = f.ncols[k];
        for (int fr = 0; fr < f.frames; ++fr) {
            float *row = f.out[k] + ((size_t)fr * f.N + v) * (size_t)f.ld[k] + f.col0[k];
            if (((nc | f.ld[k] | f.col0[k]) & 3) == 0) {
                const float4 v4 = make_float4(f.value, f.value, f.value, f.value);
                for (int c = 4 * lane; c < nc; c += 128) *reinterpret_cast<float4 *>(row + c) = v4;
            } else {
                for (int c = lane; c < nc; c += 32) row[c] = f.value;
            }
        }
    }
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API int morig_fill_cut_f32(const int32_t *const *rowptr, float *const *out, const int32_t *ld, const int32_t *col0,
                                            const int32_t *ncols, int32_t count, int32_t N, int32_t frames, float value,
                                            void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(rowptr && out && ld && col0 && ncols && count >= 1 && count <= FILL_CUT_MAX && N > 0 && frames >= 1,
                    "fill_cut_f32: count=%d (1..%d), N=%d, frames=%d", count, FILL_CUT_MAX, N, frames);
    FillCut f;
    f.count = count; f.N = N; f.frames = frames; f.value = value;
    for (int i = 0; i < count; ++i) {
        MORIG_CHECK_ARG(rowptr[i] && out[i] && ncols[i] > 0 && col0[i] >= 0 && ld[i] >= col0[i] + ncols[i] &&
                        (reinterpret_cast<uintptr_t>(out[i]) & 15u) == 0, "fill_cut_f32: entry %d", i);
        f.rowptr[i] = rowptr[i]; f.out[i] = out[i]; f.ld[i] = ld[i]; f.col0[i] = col0[i]; f.ncols[i] = ncols[i];
    }
    const int64_t blocks = ceil_div64((int64_t)N * count, 8);             // 8 warps per block, one (vertex, entry) per warp
    MORIG_CUDA(launch_pdl(fill_cut_kernel, dim3((unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks)), dim3(256), 0, stream, f));
    return 0;
}

extern "C" MORIG_API int morig_fill_many_f32(float *const *dst, const int64_t *n, int32_t count, float value, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dst && n && count >= 1 && count <= FILL_MANY_MAX, "fill_many_f32: count=%d unsupported (1..%d)", count, FILL_MANY_MAX);
    FillMany f;
    f.count = count; f.value = value;
    int64_t total = 0;
    for (int i = 0; i < count; ++i) {
        MORIG_CHECK_ARG(dst[i] && n[i] >= 0 && (reinterpret_cast<uintptr_t>(dst[i]) & 15u) == 0, "fill_many_f32: buffer %d", i);
        f.dst[i] = dst[i]; f.n[i] = n[i];
        total += n[i];
    }
    const int64_t blocks = ceil_div64(total / 4 + 1, 256);
    MORIG_CUDA(launch_pdl(fill_many_kernel, dim3((unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks)), dim3(256), 0, stream, f));
    return 0;
}

extern "C" MORIG_API int morig_temporal_attn_fwd(const float *x, int32_t N, int32_t T, int32_t C, int32_t heads, int32_t D,
                                       const float *u, const float *l0, const float *Mv, const float *c0, float *out,
                                       int32_t ldo, float *out_amax, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && u && l0 && Mv && c0 && out && N > 0, "temporal_attn_fwd: null operand");
    MORIG_CHECK_ARG(T >= 1 && T <= ATT_MAX_T, "temporal_attn_fwd: T=%d unsupported (1..%d key-frames)", T, ATT_MAX_T);
    MORIG_CHECK_ARG(C >= 1 && C <= 32, "temporal_attn_fwd: C=%d unsupported (<=32)", C);
    MORIG_CHECK_ARG(heads >= 1 && heads <= ATT_MAX_HEADS, "temporal_attn_fwd: heads=%d unsupported", heads);
    MORIG_CHECK_ARG(D >= 32 && D % 32 == 0 && (size_t)heads * C * D * 4 <= 48 * 1024,
                    "temporal_attn_fwd: D=%d unsupported (multiple of 32)", D);
    const int T_ = 256;
    const int blocks = ceil_div(N * 32, T_);
    const int cap = sm_count() * 4;
    MORIG_CUDA(launch_pdl(temporal_attn_kernel, dim3(blocks < cap ? blocks : cap), dim3(T_), (size_t)heads * C * D * sizeof(float), stream,
                          x, N, T, C, heads, D, u, l0, Mv, c0, out, ldo, out_amax));
    MORIG_LAUNCH_CHECK("temporal_attn_kernel");
    return 0;
}

extern "C" MORIG_API int morig_row_normalize(float *x, int32_t ldx, int32_t R, int32_t C, float *dst2, int32_t N, int32_t n_frames,
                                   void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && R > 0 && C > 0 && ldx >= C, "row_normalize: bad argument");
    MORIG_CHECK_ARG(!dst2 || (N > 0 && n_frames > 0 && R == N * n_frames), "row_normalize: R != N * n_frames");
    MORIG_CUDA(launch_pdl(row_normalize_kernel, dim3((unsigned)ceil_div64((int64_t)R * 32, 256)), dim3(256), 0, stream, x, ldx, R, C, dst2, N,
                          n_frames));
    MORIG_LAUNCH_CHECK("row_normalize_kernel");
    return 0;
}

extern "C" MORIG_API int morig_frame_reduce(const float *x, int32_t N, int32_t T, int32_t C, int32_t mode, float *out, int32_t ldo,
                                  void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && out && N > 0 && T > 0 && C > 0 && (mode == 0 || mode == 1), "frame_reduce: bad argument");
    MORIG_CUDA(launch_pdl(frame_reduce_kernel, dim3((unsigned)ceil_div64((int64_t)N * C, 256)), dim3(256), 0, stream, x, N, T, C, mode, out, ldo));
    MORIG_LAUNCH_CHECK("frame_reduce_kernel");
    return 0;
}

extern "C" MORIG_API int morig_gather_cols(const float *src, int32_t lds, int32_t src_off, int32_t frame_stride,
                                 const int32_t *cols, int32_t C, int32_t N, int32_t n_frames, float *dst, int32_t ldd,
                                 int32_t dst_off, float *dst_amax, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(src && dst && C > 0 && N > 0 && n_frames > 0, "gather_cols: bad argument");
    const int64_t total = (int64_t)N * n_frames * C;
    const int64_t blocks = ceil_div64(total, 256), cap = (int64_t)sm_count() * 8;
    MORIG_CUDA(launch_pdl(gather_cols_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), 0, stream, src, lds, src_off,
                          frame_stride, cols, C, N, n_frames, dst, ldd, dst_off, dst_amax));
    MORIG_LAUNCH_CHECK("gather_cols_kernel");
    return 0;
}

extern "C" MORIG_API int morig_absmax_f32(const float *x, int32_t ldx, int32_t R, int32_t C, float *amax, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && amax && R > 0 && C > 0 && ldx >= C, "absmax_f32: bad argument");
    const int64_t blocks = ceil_div64((int64_t)R * C, 256 * 8);
    MORIG_CUDA(launch_pdl(absmax_kernel, dim3((unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks)), dim3(256), 0, stream, x, ldx, R, C, amax));
    MORIG_LAUNCH_CHECK("absmax_kernel");
    return 0;
}

extern "C" MORIG_API int morig_fill_f32(float *dst, int64_t n, float value, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dst && n >= 0, "fill_f32: bad argument");
    if (n == 0) return 0;
    const int64_t blocks = ceil_div64(n, 256);
    MORIG_CUDA(launch_pdl(fill_kernel, dim3((unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks)), dim3(256), 0, stream, dst, n, value));
    MORIG_LAUNCH_CHECK("fill_kernel");
    return 0;
}
