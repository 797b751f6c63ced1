// Training path of the same networks (training/train_rig.py:136-195): the kernels behind the
// autograd functions of morig_b200/autograd_ops.py.  Everything is fp32 with fp64 accumulation where a reduction runs
// over rows (BatchNorm statistics over E or N rows, bias / BatchNorm parameter gradients), so that gradients stay
// within 1e-4 of the reference's autograd, and every reduction has a FIXED order (per-chunk partials in a workspace,
// summed by a finalize kernel) except the scatter-add of the source-side gradient dQ, which uses fp32 atomics.
//
//   transpose_pad        W [N, K] -> W^T [K, ldw]                     (weights change every step: re-packed on the device)
//   wgrad                dW [N, K] = dY^T (X * xs + xt), db = colsum dY   (split over rows, fixed-order reduction)
//   colstats             per-column  sum x, sum x^2   or   sum dy, sum dy*x   over rows (fp64 partials)
//   bn_finalize_fwd      mean / invstd / running stats / (scale, shift) of a train-mode BatchNorm1d
//   bn_finalize_bwd      dgamma, dbeta and the per-column coefficients of the input gradient
//   bn_apply             y = x * scale + shift
//   bn_relu_bwd          dz = [x > 0] * g * invstd * (dy - mean(dy) - xhat * mean(dy * xhat))     (Linear->ReLU->BN block)
//   edge_gather_relu     h[e] = relu(P[tgt[e]] + Q[col[e]])            (+ backward: dP by segments, dQ by atomics)
//   segmax               out[v] = max over the CSR segment, arg = FIRST maximal slot (torch_scatter's tie rule)
//   segmax_bwd           dy[arg[v, c], c] = dout[v, c], zero elsewhere
//   row_gather / seg_sum repeat_interleave(x_global, bincount(batch)) and its gradient
//   normalize fwd / bwd  F.normalize(dim=1)
//   attn_cls fwd / bwd   cls-query attention over the key-frames (models/rignet.py:36-45)
#include <stdlib.h>
#include <initializer_list>
#include "common.cuh"
#include "wgrad_tc.cuh"

namespace morig {

static inline bool aligned16p(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool vec4_ok(int C, std::initializer_list<int> lds, std::initializer_list<const void *> ptrs) {
    if (C % 4) return false;
    for (int ld : lds) if (ld % 4) return false;
    for (const void *p : ptrs) if (!aligned16p(p)) return false;
    return true;
}

// ---- W [rows, cols] (row stride lds) -> dst [cols, ldd], dst[c][r] = src[r][c], columns r >= rows zero -------------------
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float *__restrict__ src, int rows, int cols, int lds,
                                                            float *__restrict__ dst, int ldd) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 row-lanes
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < ldd) dst[(size_t)c * ldd + r] = tile[tx][i];
    }
}

// ---- weight gradient: dW[n, k] = sum_m dY[m, n] * (X[m, k] * xs[k] + xt[k]),  db[n] = sum_m dY[m, n] ----------------------
// 128 x 128 tile of dW per CTA, 8 x 8 micro-tile per thread, 16 rows of both operands per step (row-major loads are
// already the layout the outer products need).  blockIdx.z splits the rows; partials go to the workspace and are added
// in a fixed order by wgrad_reduce_kernel.
constexpr int WG_M = 16, WG_THREADS = 256;

// T = 128: 8 x 8 micro-tile per thread; T = 64 (narrow layers: N, K <= 64, e.g. the 16- / 32-channel edge branches, where a
// 128 x 128 tile would waste 15/16 of its FMAs): 4 x 4 micro-tile
template <int T>
__global__ void __launch_bounds__(WG_THREADS, 2) wgrad_kernel(const float *__restrict__ dY, int lddy, const float *__restrict__ X,
                                                              int ldx, int M, int N, int K, const float *__restrict__ xs,
                                                              const float *__restrict__ xt, int rows_per_split,
                                                              float *__restrict__ part, float *__restrict__ part_b,
                                                              int vec_y, int vec_x) {
    constexpr int MT = T / 16;                      // micro-tile edge
    constexpr int C4 = T / 4;                       // float4 per tile row
    constexpr int RPP = WG_THREADS / C4;            // rows loaded per pass (8 or 16)
    constexpr int PASSES = WG_M / RPP;              // 2 or 1
    __shared__ __align__(16) float Ys[2][WG_M][T];
    __shared__ __align__(16) float Xs[2][WG_M][T];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * T, k0 = blockIdx.y * T;
    const int m_begin = blockIdx.z * rows_per_split, m_end = min(M, m_begin + rows_per_split);
    const int lrow = tid / C4, lcol = (tid % C4) * 4;
    float4 ry[PASSES], rx[PASSES];
    float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        float *s = reinterpret_cast<float *>(&s4), *t = reinterpret_cast<float *>(&t4);
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + lcol + q;
            if (k < K) { if (xs) s[q] = xs[k]; if (xt) t[q] = xt[k]; }
        }
    }
    auto load = [&](int m0) {
#pragma unroll
        for (int l = 0; l < PASSES; ++l) {
            const int m = m0 + lrow + RPP * l;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (m < m_end) {
                const int n = n0 + lcol, k = k0 + lcol;
                const float *py = dY + (size_t)m * lddy + n;
                const float *px = X + (size_t)m * ldx + k;
                if (vec_y && n + 3 < N) a = *reinterpret_cast<const float4 *>(py);
                else {
                    if (n + 0 < N) a.x = py[0];
                    if (n + 1 < N) a.y = py[1];
                    if (n + 2 < N) a.z = py[2];
                    if (n + 3 < N) a.w = py[3];
                }
                if (vec_x && k + 3 < K) {
                    b = *reinterpret_cast<const float4 *>(px);
                    b.x = fmaf(b.x, s4.x, t4.x); b.y = fmaf(b.y, s4.y, t4.y); b.z = fmaf(b.z, s4.z, t4.z); b.w = fmaf(b.w, s4.w, t4.w);
                } else {
                    if (k + 0 < K) b.x = fmaf(px[0], s4.x, t4.x);
                    if (k + 1 < K) b.y = fmaf(px[1], s4.y, t4.y);
                    if (k + 2 < K) b.z = fmaf(px[2], s4.z, t4.z);
                    if (k + 3 < K) b.w = fmaf(px[3], s4.w, t4.w);
                }
            }
            ry[l] = a; rx[l] = b;
        }
    };
    auto store = [&](int buf) {
#pragma unroll
        for (int l = 0; l < PASSES; ++l) {
            *reinterpret_cast<float4 *>(&Ys[buf][lrow + RPP * l][lcol]) = ry[l];
            *reinterpret_cast<float4 *>(&Xs[buf][lrow + RPP * l][lcol]) = rx[l];
        }
    };
    float acc[MT][MT];
    float bsum[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        bsum[i] = 0.f;
#pragma unroll
        for (int j = 0; j < MT; ++j) acc[i][j] = 0.f;
    }
    const int steps = (m_end > m_begin) ? (m_end - m_begin + WG_M - 1) / WG_M : 0;
    if (steps > 0) { load(m_begin); store(0); }
    __syncthreads();
    for (int st = 0; st < steps; ++st) {
        if (st + 1 < steps) load(m_begin + (st + 1) * WG_M);
        const int buf = st & 1;
#pragma unroll
        for (int m = 0; m < WG_M; ++m) {
            float a[MT], b[MT];
#pragma unroll
            for (int h = 0; h < MT / 4; ++h) {
                const float4 av = *reinterpret_cast<const float4 *>(&Ys[buf][m][h * 64 + ty * 4]);
                const float4 bv = *reinterpret_cast<const float4 *>(&Xs[buf][m][h * 64 + tx * 4]);
                a[4 * h] = av.x; a[4 * h + 1] = av.y; a[4 * h + 2] = av.z; a[4 * h + 3] = av.w;
                b[4 * h] = bv.x; b[4 * h + 1] = bv.y; b[4 * h + 2] = bv.z; b[4 * h + 3] = bv.w;
            }
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                bsum[i] += a[i];
                // packed fp32 pairs (FFMA2): one instruction per two accumulators of the row
                const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < MT; j += 2) {
                    uint64_t r;
                    const float2 b2 = make_float2(b[j], b[j + 1]), c2 = make_float2(acc[i][j], acc[i][j + 1]);
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t *>(&a2)),
                        "l"(*reinterpret_cast<const uint64_t *>(&b2)), "l"(*reinterpret_cast<const uint64_t *>(&c2)));
                    const float2 o = *reinterpret_cast<const float2 *>(&r);
                    acc[i][j] = o.x; acc[i][j + 1] = o.y;
                }
            }
        }
        if (st + 1 < steps) store((st + 1) & 1);
        __syncthreads();
    }
    float *out = part + (size_t)blockIdx.z * N * K;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int n = n0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            const int k = k0 + (j >> 2) * 64 + tx * 4 + (j & 3);
            if (k < K) out[(size_t)n * K + k] = acc[i][j];
        }
        if (part_b && blockIdx.y == 0 && tx == 0) part_b[(size_t)blockIdx.z * N + n] = bsum[i];
    }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ part, const float *__restrict__ part_b,
                                                           int splits, int N, int K, float *__restrict__ dW, int lddw,
                                                           float *__restrict__ db, int accumulate) {
    const int64_t total = (int64_t)N * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + N; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < total) {
            double s = 0.0;
            for (int sp = 0; sp < splits; ++sp) s += (double)part[(size_t)sp * total + i];
            const int n = (int)(i / K), k = (int)(i % K);
            float *dst = dW + (size_t)n * lddw + k;
            *dst = (float)(s + (accumulate ? (double)*dst : 0.0));
        } else if (db && part_b) {
            const int n = (int)(i - total);
            double s = 0.0;
            for (int sp = 0; sp < splits; ++sp) s += (double)part_b[(size_t)sp * N + n];
            db[n] = (float)(s + (accumulate ? (double)db[n] : 0.0));
        }
    }
}

// ---- per-column statistics over rows (fp64 partials per row chunk) -----------------------------------------------------
//   Y == NULL:  p0 = sum x,   p1 = sum x^2
//   Y != NULL:  p0 = sum y,   p1 = sum y * x          (y = incoming gradient, x = saved BatchNorm input)
// block (32 columns, 8 row-lanes); every warp reads 128 contiguous bytes per row
__global__ void __launch_bounds__(256) colstats_kernel(const float *__restrict__ X, int ldx, const float *__restrict__ Y, int ldy,
                                                       int R, int C, int rows_per_chunk, double *__restrict__ part) {
    __shared__ double s0[8][32], s1[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int r_begin = blockIdx.y * rows_per_chunk, r_end = min(R, r_begin + rows_per_chunk);
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
        // four rows in flight per thread (one 4-byte load per row and thread would leave the kernel latency bound at a
        // third of the HBM rate); the sums are still taken row by row in ascending order
        int r = r_begin + ry;
        for (; r + 24 < r_end; r += 32) {
            float xv[4], yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                xv[u] = X[(size_t)(r + 8 * u) * ldx + c];
                yv[u] = Y ? Y[(size_t)(r + 8 * u) * ldy + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (Y) { a0 += (double)yv[u]; a1 += (double)yv[u] * (double)xv[u]; }
                else { a0 += (double)xv[u]; a1 += (double)xv[u] * (double)xv[u]; }
            }
        }
        for (; r < r_end; r += 8) {
            const float x = X[(size_t)r * ldx + c];
            if (Y) {
                const float y = Y[(size_t)r * ldy + c];
                a0 += (double)y; a1 += (double)y * (double)x;
            } else {
                a0 += (double)x; a1 += (double)x * (double)x;
            }
        }
    }
    s0[ry][cx] = a0; s1[ry][cx] = a1;
    __syncthreads();
    if (ry == 0 && c < C) {
        for (int i = 1; i < 8; ++i) { a0 += s0[i][cx]; a1 += s1[i][cx]; }
        part[((size_t)blockIdx.y * 2 + 0) * C + c] = a0;
        part[((size_t)blockIdx.y * 2 + 1) * C + c] = a1;
    }
}

// sum of the per-chunk fp64 partials of one column: lanes take chunks lane, lane + 32, ... (ascending), then a butterfly --
// a fixed order for a given chunk count, so results are reproducible run to run
__device__ __forceinline__ void column_partials(const double *__restrict__ part, int chunks, int C, int c, int lane, double &a0,
                                                double &a1) {
    a0 = 0.0; a1 = 0.0;
    for (int i = lane; i < chunks; i += 32) { a0 += part[((size_t)i * 2 + 0) * C + c]; a1 += part[((size_t)i * 2 + 1) * C + c]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, off);
        a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    }
}

// train-mode BatchNorm1d forward statistics (torch semantics: biased variance normalises, the unbiased one updates
// running_var; momentum 0.1 in the reference's MLP, models/basic_modules.py:33).  One warp per column.
__global__ void __launch_bounds__(256) bn_finalize_fwd_kernel(const double *__restrict__ part, int chunks, int R, int C,
                                       const float *__restrict__ gamma,
                                       const float *__restrict__ beta, float eps, float momentum, float *running_mean,
                                       float *running_var, float *__restrict__ mean, float *__restrict__ invstd,
                                       float *__restrict__ scale, float *__restrict__ shift) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double a0, a1;
    column_partials(part, chunks, C, c, lane, a0, a1);
    if (lane != 0) return;
    const double mu = a0 / R;
    double var = a1 / R - mu * mu;
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)eps);
    mean[c] = (float)mu;
    invstd[c] = (float)is;
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    scale[c] = (float)(g * is);
    shift[c] = (float)(b - mu * g * is);
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
    if (running_var) {
        const double unbiased = R > 1 ? var * ((double)R / (double)(R - 1)) : var;
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// backward: dbeta = sum dy, dgamma = sum dy * xhat; coefficients of dx = a * (dy - m1 - xhat * m2)
__global__ void __launch_bounds__(256) bn_finalize_bwd_kernel(const double *__restrict__ part, int chunks, int R, int C,
                                       const float *__restrict__ gamma,
                                       const float *__restrict__ mean, const float *__restrict__ invstd,
                                       float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ coef) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double a0, a1;
    column_partials(part, chunks, C, c, lane, a0, a1);
    if (lane != 0) return;
    const double mu = mean[c], is = invstd[c];
    const double dg = is * (a1 - mu * a0);
    if (dgamma) dgamma[c] = (float)dg;
    if (dbeta) dbeta[c] = (float)a0;
    coef[c] = (float)((gamma ? (double)gamma[c] : 1.0) * is);      // a
    coef[C + c] = (float)(a0 / R);                                   // m1 = mean(dy)
    coef[2 * C + c] = (float)(dg / R);                               // m2 = mean(dy * xhat)
}

// Row-wise element kernels below: VEC = 1 handles four consecutive columns per thread with 16-byte accesses (C, the row
// strides and the base addresses multiples of 4 floats / 16 bytes -- every wide layer of the networks); the scalar form
// ran at a third to a half of the HBM rate.  `amax` (optional) receives max |output| through one ordered-int atomic per
// warp: the fp16-split tensor-core GEMM that consumes the output needs its range and would otherwise re-read it.
template <int VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float *__restrict__ x, int ldx, int R, int C,
                                                       const float *__restrict__ scale, const float *__restrict__ shift,
                                                       float *__restrict__ y, int ldy, float *amax) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)R * CW;
    float am = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / CW;
        const int c = (int)(i - r * CW) * W;
        if (VEC) {
            const float4 v = *reinterpret_cast<const float4 *>(x + (size_t)r * ldx + c);
            const float4 sc = *reinterpret_cast<const float4 *>(scale + c), sh = *reinterpret_cast<const float4 *>(shift + c);
            const float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
            *reinterpret_cast<float4 *>(y + (size_t)r * ldy + c) = o;
            am = fmaxf(fmaxf(am, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
        } else {
            const float o = fmaf(x[(size_t)r * ldx + c], scale[c], shift[c]);
            y[(size_t)r * ldy + c] = o;
            am = fmaxf(am, fabsf(o));
        }
    }
    if (amax) amax_commit_block(amax, am);
}

// dz = mask * a * (dy - m1 - xhat * m2), xhat = (x - mean) * invstd; mask = [x > 0] when the block has a ReLU in front of
// the BatchNorm (x is the ReLU output, so x > 0 <=> pre-activation > 0; torch's relu'(0) = 0)
template <int VEC>
__global__ void __launch_bounds__(256) bn_relu_bwd_kernel(const float *__restrict__ dy, int lddy, const float *__restrict__ x,
                                                          int ldx, int R, int C, const float *__restrict__ mean,
                                                          const float *__restrict__ invstd, const float *__restrict__ coef,
                                                          int relu, float *__restrict__ dz, int lddz, float *amax) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)R * CW;
    float am = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / CW;
        const int c = (int)(i - r * CW) * W;
        float xv[W], dv[W], o[W];
        if (VEC) {
            *reinterpret_cast<float4 *>(xv) = *reinterpret_cast<const float4 *>(x + (size_t)r * ldx + c);
            *reinterpret_cast<float4 *>(dv) = *reinterpret_cast<const float4 *>(dy + (size_t)r * lddy + c);
        } else {
            xv[0] = x[(size_t)r * ldx + c];
            dv[0] = dy[(size_t)r * lddy + c];
        }
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const float xhat = (xv[q] - mean[c + q]) * invstd[c + q];
            const float g = coef[c + q] * (dv[q] - coef[C + c + q] - xhat * coef[2 * C + c + q]);
            o[q] = (!relu || xv[q] > 0.f) ? g : 0.f;
            am = fmaxf(am, fabsf(o[q]));
        }
        if (VEC) *reinterpret_cast<float4 *>(dz + (size_t)r * lddz + c) = *reinterpret_cast<const float4 *>(o);
        else dz[(size_t)r * lddz + c] = o[0];
    }
    if (amax) amax_commit_block(amax, am);
}

// dz = [y > 0] * dy : backward of a bare ReLU (Linear -> ReLU without BatchNorm is not used by the rigging nets, but the
// Linear-only heads share the code path with relu = 0)
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float *__restrict__ dy, int lddy, const float *__restrict__ y, int ldy,
                                                       int R, int C, float *__restrict__ dz, int lddz) {
    const int64_t total = (int64_t)R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t r = i / C;
        dz[(size_t)r * lddz + c] = y[(size_t)r * ldy + c] > 0.f ? dy[(size_t)r * lddy + c] : 0.f;
    }
}

// ---- per-edge first layer after factorisation: h[e] = relu(P[tgt[e]] + Q[col[e]]) --------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) edge_gather_relu_kernel(const float *__restrict__ P, int ldp, const float *__restrict__ Q,
                                                               int ldq, const int32_t *__restrict__ tgt,
                                                               const int32_t *__restrict__ col, int E, int C,
                                                               float *__restrict__ h, int ldh, float *amax) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)E * CW;
    float am = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / CW);
        const int c = (int)(i - (int64_t)e * CW) * W;
        if (VEC) {
            const float4 a = *reinterpret_cast<const float4 *>(P + (size_t)tgt[e] * ldp + c);
            const float4 b = *reinterpret_cast<const float4 *>(Q + (size_t)col[e] * ldq + c);
            const float4 o = make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
            *reinterpret_cast<float4 *>(h + (size_t)e * ldh + c) = o;
            am = fmaxf(fmaxf(am, fmaxf(o.x, o.y)), fmaxf(o.z, o.w));
        } else {
            const float o = fmaxf(P[(size_t)tgt[e] * ldp + c] + Q[(size_t)col[e] * ldq + c], 0.f);
            h[(size_t)e * ldh + c] = o;
            am = fmaxf(am, o);
        }
    }
    if (amax) amax_commit_block(amax, am);
}

// dP[v] = sum over the CSR segment of v of [h > 0] dh   (one thread per (v, column group): fixed order)
template <int VEC>
__global__ void __launch_bounds__(256) edge_gather_bwd_p_kernel(const float *__restrict__ dh, int lddh, const float *__restrict__ h,
                                                                int ldh, const int32_t *__restrict__ rowptr, int N, int C,
                                                                float *__restrict__ dP, int ldp) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)N * CW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i / CW);
        const int c = (int)(i - (int64_t)v * CW) * W;
        float s[W];
#pragma unroll
        for (int q = 0; q < W; ++q) s[q] = 0.f;
        for (int e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            float hv[W], dv[W];
            if (VEC) {
                *reinterpret_cast<float4 *>(hv) = *reinterpret_cast<const float4 *>(h + (size_t)e * ldh + c);
                *reinterpret_cast<float4 *>(dv) = *reinterpret_cast<const float4 *>(dh + (size_t)e * lddh + c);
            } else {
                hv[0] = h[(size_t)e * ldh + c];
                dv[0] = dh[(size_t)e * lddh + c];
            }
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (hv[q] > 0.f) s[q] += dv[q];
        }
        if (VEC) *reinterpret_cast<float4 *>(dP + (size_t)v * ldp + c) = *reinterpret_cast<const float4 *>(s);
        else dP[(size_t)v * ldp + c] = s[0];
    }
}

// dQ[col[e]] += [h > 0] dh   (dQ zeroed by the caller; fp32 atomics: the one order-dependent sum of the path).
// VEC: one 16-byte vector reduction (red.global.add.v4.f32, sm_90+) per four columns instead of four scalar atomics.
template <int VEC>
__global__ void __launch_bounds__(256) edge_gather_bwd_q_kernel(const float *__restrict__ dh, int lddh, const float *__restrict__ h,
                                                                int ldh, const int32_t *__restrict__ col, int E, int C,
                                                                float *__restrict__ dQ, int ldq) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)E * CW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / CW);
        const int c = (int)(i - (int64_t)e * CW) * W;
        if (VEC) {
            const float4 hv = *reinterpret_cast<const float4 *>(h + (size_t)e * ldh + c);
            float4 dv = *reinterpret_cast<const float4 *>(dh + (size_t)e * lddh + c);
            dv.x = hv.x > 0.f ? dv.x : 0.f; dv.y = hv.y > 0.f ? dv.y : 0.f;
            dv.z = hv.z > 0.f ? dv.z : 0.f; dv.w = hv.w > 0.f ? dv.w : 0.f;
            if (hv.x > 0.f || hv.y > 0.f || hv.z > 0.f || hv.w > 0.f)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                             :: "l"(dQ + (size_t)col[e] * ldq + c), "f"(dv.x), "f"(dv.y), "f"(dv.z), "f"(dv.w) : "memory");
        } else {
            if (h[(size_t)e * ldh + c] > 0.f) atomicAdd(dQ + (size_t)col[e] * ldq + c, dh[(size_t)e * lddh + c]);
        }
    }
}

// ---- segmented max, short segments (the max aggregation over a vertex's in-edges: ~8 / ~17 rows per segment) --------------
// One thread per (segment, column group) walks the segment's rows in order and keeps the first maximum (strict >), with
// 16-byte loads when VEC: no shared memory and no block barriers (segmax_kernel below spends two __syncthreads per
// segment, which is right for the per-graph pooling over thousands of rows and wrong for 17).
template <int VEC>
__global__ void __launch_bounds__(256) segmax_short_kernel(const float *__restrict__ y, int ldy, const int32_t *__restrict__ ptr,
                                                           int S, int C, float *__restrict__ out, int ldo,
                                                           int32_t *__restrict__ arg, int lda) {
    constexpr int W = VEC ? 4 : 1;
    const int CW = C / W;
    const int64_t total = (int64_t)S * CW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int sgm = (int)(i / CW);
        const int c = (int)(i - (int64_t)sgm * CW) * W;
        const int lo = ptr[sgm], hi = ptr[sgm + 1];
        float best[W];
        int bi[W];
#pragma unroll
        for (int q = 0; q < W; ++q) { best[q] = 0.f; bi[q] = -1; }
        for (int r = lo; r < hi; ++r) {
            float v[W];
            if (VEC) *reinterpret_cast<float4 *>(v) = *reinterpret_cast<const float4 *>(y + (size_t)r * ldy + c);
            else v[0] = y[(size_t)r * ldy + c];
#pragma unroll
            for (int q = 0; q < W; ++q)
                if (bi[q] < 0 || v[q] > best[q]) { best[q] = v[q]; bi[q] = r; }
        }
        if (VEC) {                                                    // empty segment: 0 / -1
            *reinterpret_cast<float4 *>(out + (size_t)sgm * ldo + c) = *reinterpret_cast<const float4 *>(best);
            if (arg) *reinterpret_cast<int4 *>(arg + (size_t)sgm * lda + c) = *reinterpret_cast<const int4 *>(bi);
        } else {
            out[(size_t)sgm * ldo + c] = best[0];
            if (arg) arg[(size_t)sgm * lda + c] = bi[0];
        }
    }
}

// ---- segmented max with argmax over contiguous row segments ptr[s] .. ptr[s+1] ------------------------------------------
// block (32 columns, 8 row-lanes): lane ry scans rows ptr[s] + ry, + 8, ... keeping its FIRST maximum (strict >); the
// eight candidates merge with (larger value, then smaller row) -- i.e. the first maximal row of the segment, the rule
// of torch_scatter's scatter_max that the reference's max aggregation and its backward follow.  Empty segment -> 0 / -1.
__global__ void __launch_bounds__(256) segmax_kernel(const float *__restrict__ y, int ldy, const int32_t *__restrict__ ptr, int S,
                                                     int C, float *__restrict__ out, int ldo, int32_t *__restrict__ arg,
                                                     int lda) {
    __shared__ float sv[8][32];
    __shared__ int si[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    for (int s = blockIdx.y; s < S; s += gridDim.y) {
        const int lo = ptr[s], hi = ptr[s + 1];
        float best = neg_inf();
        int bi = -1;
        if (c < C) {
            for (int r = lo + ry; r < hi; r += 8) {
                const float v = y[(size_t)r * ldy + c];
                if (bi < 0 || v > best) { best = v; bi = r; }
            }
        }
        sv[ry][cx] = best; si[ry][cx] = bi;
        __syncthreads();
        if (ry == 0 && c < C) {
            for (int i = 1; i < 8; ++i) {
                const float v = sv[i][cx];
                const int j = si[i][cx];
                if (j >= 0 && (bi < 0 || v > best || (v == best && j < bi))) { best = v; bi = j; }
            }
            out[(size_t)s * ldo + c] = bi >= 0 ? best : 0.f;
            if (arg) arg[(size_t)s * lda + c] = bi;
        }
        __syncthreads();
    }
}

// dy (zero-filled by the caller) [arg[s, c], c] = dout[s, c]
__global__ void __launch_bounds__(256) segmax_bwd_kernel(const float *__restrict__ dout, int lddo, const int32_t *__restrict__ arg,
                                                         int lda, int S, int C, float *__restrict__ dy, int lddy) {
    const int64_t total = (int64_t)S * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int s = (int)(i / C);
        const int r = arg[(size_t)s * lda + c];
        if (r >= 0) dy[(size_t)r * lddy + c] = dout[(size_t)s * lddo + c];
    }
}

// segment pointers of a sorted key vector with every key 0..S-1 present (PyG `batch`): ptr[k] = first row of key k
__global__ void seg_ptr_kernel(const int32_t *__restrict__ keys, int N, int S, int32_t *__restrict__ ptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    if (i == N) { ptr[S] = N; return; }
    const int k = keys[i];
    if (i == 0) { for (int j = 0; j <= k && j < S; ++j) ptr[j] = 0; }
    else {
        const int kp = keys[i - 1];
        for (int j = kp + 1; j <= k && j < S; ++j) ptr[j] = i;      // empty keys in between start here too
    }
}

// out[r] = src[idx[r]]   (repeat_interleave of the per-graph feature)
__global__ void __launch_bounds__(256) row_gather_kernel(const float *__restrict__ src, int lds, const int32_t *__restrict__ idx,
                                                         int R, int C, float *__restrict__ out, int ldo) {
    const int64_t total = (int64_t)R * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int r = (int)(i / C);
        out[(size_t)r * ldo + c] = src[(size_t)idx[r] * lds + c];
    }
}

// out[s] = sum of rows ptr[s] .. ptr[s+1]   (fixed order: 8 strided lanes, then lane 0 adds them 0..7)
__global__ void __launch_bounds__(256) seg_sum_kernel(const float *__restrict__ x, int ldx, const int32_t *__restrict__ ptr, int S,
                                                      int C, float *__restrict__ out, int ldo) {
    __shared__ double sv[8][32];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    for (int s = blockIdx.y; s < S; s += gridDim.y) {
        double a = 0.0;
        if (c < C)
            for (int r = ptr[s] + ry; r < ptr[s + 1]; r += 8) a += (double)x[(size_t)r * ldx + c];
        sv[ry][cx] = a;
        __syncthreads();
        if (ry == 0 && c < C) {
            for (int i = 1; i < 8; ++i) a += sv[i][cx];
            out[(size_t)s * ldo + c] = (float)a;
        }
        __syncthreads();
    }
}

// ---- F.normalize(dim=1), out of place, and its backward: warp per row ------------------------------------------------------
__global__ void __launch_bounds__(256) normalize_fwd_kernel(const float *__restrict__ x, int ldx, int R, int C,
                                                            float *__restrict__ y, int ldy) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const float *row = x + (size_t)r * ldx;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = row[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < C; c += 32) y[(size_t)r * ldy + c] = row[c] / denom;
}

// dx = (dy - y (y . dy)) / max(|x|, eps)  for |x| > eps, dy / eps otherwise (the clamp is then constant)
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const float *__restrict__ x, int ldx, const float *__restrict__ dy,
                                                            int lddy, int R, int C, float *__restrict__ dx, int lddx) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= R) return;
    const float *row = x + (size_t)r * ldx, *g = dy + (size_t)r * lddy;
    float ss = 0.f, dot = 0.f;
    for (int c = lane; c < C; c += 32) { const float v = row[c]; ss = fmaf(v, v, ss); dot = fmaf(v, g[c], dot); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
        dot += __shfl_xor_sync(0xffffffffu, dot, off);
    }
    const float nrm = sqrtf(ss);
    if (nrm > 1e-12f) {
        const float inv = 1.f / nrm, k = dot * inv * inv * inv;          // (x . dy) / |x|^3
        for (int c = lane; c < C; c += 32) dx[(size_t)r * lddx + c] = g[c] * inv - row[c] * k;
    } else {
        for (int c = lane; c < C; c += 32) dx[(size_t)r * lddx + c] = g[c] * 1e12f;
    }
}

// ---- cls-query attention over the key-frames (models/rignet.py:36-45; only output row 0 is used) --------------------------
//   q0, kc, vc [HD]      query / key / value of the cls token (parameter-only rows)
//   Kx, Vx [N, T, HD]    keys / values of the T key-frame tokens;  HD = heads * d
//   out [N, HD] = concat_h ( a_cls,h vc_h + sum_t a_t,h Vx[n, t, h] ),  a = softmax over the T + 1 logits q0_h . k / sqrt(d)
//   att [N, heads, T + 1] is saved for the backward (slot 0 = cls)
// warp per vertex; lane l owns channels l * (HD / 32) .. of every token
constexpr int ATTN_MAX_T = 8, ATTN_MAX_CPL = 8;

__global__ void __launch_bounds__(256) attn_cls_fwd_kernel(const float *__restrict__ q0, const float *__restrict__ kc,
                                                           const float *__restrict__ vc, const float *__restrict__ Kx,
                                                           const float *__restrict__ Vx, int N, int T, int HD, int d,
                                                           float *__restrict__ out, float *__restrict__ att) {
    const int lane = threadIdx.x & 31;
    const int cpl = HD / 32, lph = d / cpl, heads = HD / d;       // channels per lane, lanes per head
    const int head = lane / lph;
    const float rs = rsqrtf((float)d);
    float q[ATTN_MAX_CPL], kcl[ATTN_MAX_CPL], vcl[ATTN_MAX_CPL];
#pragma unroll
    for (int i = 0; i < ATTN_MAX_CPL; ++i)
        if (i < cpl) { q[i] = q0[lane * cpl + i] * rs; kcl[i] = kc[lane * cpl + i]; vcl[i] = vc[lane * cpl + i]; }
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        float lg[ATTN_MAX_T + 1];
#pragma unroll
        for (int t = 0; t <= ATTN_MAX_T; ++t) {
            if (t > T) continue;
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < ATTN_MAX_CPL; ++i)
                if (i < cpl) s = fmaf(q[i], t == 0 ? kcl[i] : Kx[((size_t)n * T + (t - 1)) * HD + lane * cpl + i], s);
            for (int off = lph >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            lg[t] = s;
        }
        float mx = lg[0];
#pragma unroll
        for (int t = 1; t <= ATTN_MAX_T; ++t) if (t <= T) mx = fmaxf(mx, lg[t]);
        float den = 0.f;
#pragma unroll
        for (int t = 0; t <= ATTN_MAX_T; ++t) if (t <= T) { lg[t] = expf(lg[t] - mx); den += lg[t]; }
        const float inv = 1.f / den;
        float o[ATTN_MAX_CPL];
#pragma unroll
        for (int i = 0; i < ATTN_MAX_CPL; ++i) o[i] = 0.f;
#pragma unroll
        for (int t = 0; t <= ATTN_MAX_T; ++t) {
            if (t > T) continue;
            const float a = lg[t] * inv;
            if ((lane % lph) == 0) att[((size_t)n * heads + head) * (T + 1) + t] = a;
#pragma unroll
            for (int i = 0; i < ATTN_MAX_CPL; ++i)
                if (i < cpl) o[i] = fmaf(a, t == 0 ? vcl[i] : Vx[((size_t)n * T + (t - 1)) * HD + lane * cpl + i], o[i]);
        }
#pragma unroll
        for (int i = 0; i < ATTN_MAX_CPL; ++i)
            if (i < cpl) out[(size_t)n * HD + lane * cpl + i] = o[i];
    }
}

// backward: dKx, dVx per vertex; dq0 / dkc / dvc are sums over the vertices: per-warp partial in registers, per-CTA
// partial through shared memory, written to part [gridDim.x][3][HD]; attn_cls_reduce_kernel adds the CTAs in order
__global__ void __launch_bounds__(256) attn_cls_bwd_kernel(const float *__restrict__ q0, const float *__restrict__ kc,
                                                           const float *__restrict__ vc, const float *__restrict__ Kx,
                                                           const float *__restrict__ Vx, const float *__restrict__ att,
                                                           const float *__restrict__ dout, int N, int T, int HD, int d,
                                                           float *__restrict__ dKx, float *__restrict__ dVx,
                                                           float *__restrict__ part) {
    extern __shared__ float s_part[];                             // [8 warps][3][HD]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cpl = HD / 32, lph = d / cpl, heads = HD / d;
    const int head = lane / lph;
    const float rs = rsqrtf((float)d);
    float q[ATTN_MAX_CPL], kcl[ATTN_MAX_CPL], vcl[ATTN_MAX_CPL], aq[ATTN_MAX_CPL], ak[ATTN_MAX_CPL], av[ATTN_MAX_CPL];
#pragma unroll
    for (int i = 0; i < ATTN_MAX_CPL; ++i) {
        aq[i] = ak[i] = av[i] = 0.f;
        if (i < cpl) { q[i] = q0[lane * cpl + i]; kcl[i] = kc[lane * cpl + i]; vcl[i] = vc[lane * cpl + i]; }
    }
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; n < N; n += warps_total) {
        float g[ATTN_MAX_CPL];
#pragma unroll
        for (int i = 0; i < ATTN_MAX_CPL; ++i) if (i < cpl) g[i] = dout[(size_t)n * HD + lane * cpl + i];
        float a[ATTN_MAX_T + 1], da[ATTN_MAX_T + 1];
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t <= ATTN_MAX_T; ++t) {
            if (t > T) continue;
            a[t] = att[((size_t)n * heads + head) * (T + 1) + t];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < ATTN_MAX_CPL; ++i)
                if (i < cpl) s = fmaf(g[i], t == 0 ? vcl[i] : Vx[((size_t)n * T + (t - 1)) * HD + lane * cpl + i], s);
            for (int off = lph >> 1; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            da[t] = s;
            dot = fmaf(a[t], s, dot);
        }
#pragma unroll
        for (int t = 0; t <= ATTN_MAX_T; ++t) {
            if (t > T) continue;
            const float dl = a[t] * (da[t] - dot) * rs;           // d logit (scaled dot product)
#pragma unroll
            for (int i = 0; i < ATTN_MAX_CPL; ++i) {
                if (i >= cpl) continue;
                if (t == 0) {
                    av[i] = fmaf(a[0], g[i], av[i]);
                    ak[i] = fmaf(dl, q[i], ak[i]);
                    aq[i] = fmaf(dl, kcl[i], aq[i]);
                } else {
                    const size_t off = ((size_t)n * T + (t - 1)) * HD + lane * cpl + i;
                    dVx[off] = a[t] * g[i];
                    dKx[off] = dl * q[i];
                    aq[i] = fmaf(dl, Kx[off], aq[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ATTN_MAX_CPL; ++i) {
        if (i >= cpl) continue;
        s_part[(warp * 3 + 0) * HD + lane * cpl + i] = aq[i];
        s_part[(warp * 3 + 1) * HD + lane * cpl + i] = ak[i];
        s_part[(warp * 3 + 2) * HD + lane * cpl + i] = av[i];
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += s_part[w * 3 * HD + i];
        part[(size_t)blockIdx.x * 3 * HD + i] = s;
    }
}

__global__ void attn_cls_reduce_kernel(const float *__restrict__ part, int blocks, int HD, float *__restrict__ dq0,
                                       float *__restrict__ dkc, float *__restrict__ dvc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * HD) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += (double)part[(size_t)b * 3 * HD + i];
    float *dst = i < HD ? dq0 : (i < 2 * HD ? dkc : dvc);
    dst[i % HD] = (float)s;
}

static inline unsigned grid1d(int64_t total, int threads, int per_sm) {
    const int64_t blocks = ceil_div64(total, threads), cap = (int64_t)sm_count() * per_sm;
    return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API int morig_transpose_pad_f32(const float *src, int32_t rows, int32_t cols, int32_t lds, float *dst,
                                                 int32_t ldd, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(src && dst && rows > 0 && cols > 0 && lds >= cols && ldd >= rows, "transpose_pad: bad argument");
    transpose_pad_kernel<<<dim3(ceil_div(ldd, 32), ceil_div(cols, 32)), 256, 0, stream>>>(src, rows, cols, lds, dst, ldd);
    MORIG_LAUNCH_CHECK("transpose_pad_kernel");
    return 0;
}

static inline int wgrad_tile(int N, int K) { return (N <= 64 && K <= 64) ? 64 : 128; }

// launch geometry of a weight gradient: tensor-core kernel (csrc/wgrad_tc.cuh) for layers that fill a 128-row tile and
// have enough rows to amortise its set-up, CUDA-core kernel otherwise (narrow layers, tiny batches, MORIG_TRAIN_TC=0)
struct WgradPlan { int tc, bkw, splits, rows; };
static WgradPlan wgrad_plan(int M, int N, int K) {
    static int tc_off = -1;
    if (tc_off < 0) {
        const char *e = getenv("MORIG_TRAIN_TC");
        tc_off = (e && e[0] == '0') ? 1 : 0;
    }
    WgradPlan p{};
    p.tc = (!tc_off && M >= 2048 && N >= 64 && K >= 64) ? 1 : 0;
    if (p.tc) {
        p.bkw = K > 128 ? 256 : 128;
        const int tiles = ceil_div(N, tcw::W_BN) * ceil_div(K, p.bkw);
        int want = sm_count() / tiles;                              // one CTA per SM, at most one wave
        const int max_by_rows = M / 512;                            // at least 16 stages of 32 rows per slice
        if (want > max_by_rows) want = max_by_rows;
        if (want > 256) want = 256;
        p.splits = want < 1 ? 1 : want;
        p.rows = ceil_div(ceil_div(M, p.splits), tcw::W_ROWS) * tcw::W_ROWS;
        return p;
    }
    const int WG_T = wgrad_tile(N, K);
    const int tiles = ceil_div(N, WG_T) * ceil_div(K, WG_T);
    int want = ceil_div(2 * sm_count() * 2, tiles);                 // ~2 waves of 2 CTAs per SM
    const int max_by_rows = ceil_div(M, 4 * WG_M);                  // at least 64 rows per split
    if (want > max_by_rows) want = max_by_rows;
    if (want > 256) want = 256;
    p.splits = want < 1 ? 1 : want;
    p.rows = ceil_div(ceil_div(M, p.splits), WG_M) * WG_M;
    return p;
}

extern "C" MORIG_API int32_t morig_wgrad_splits(int32_t M, int32_t N, int32_t K) { return wgrad_plan(M, N, K).splits; }

extern "C" MORIG_API size_t morig_wgrad_workspace(int32_t M, int32_t N, int32_t K) {
    return (size_t)morig_wgrad_splits(M, N, K) * ((size_t)N * K + N) * sizeof(float);
}

template <int BKW>
static int launch_wgrad_tc(const float *dY, int lddy, const float *X, int ldx, int M, int N, int K, const WgradPlan &pl, float *part,
                           float *part_b, cudaStream_t stream) {
    auto kern = tcw::wgrad_tc_kernel<BKW>;
    constexpr int smem = tcw::WCfg<BKW>::SMEM;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_dev = dev;
    }
    kern<<<dim3(ceil_div(N, tcw::W_BN), ceil_div(K, BKW), pl.splits), tcw::W_THREADS, smem, stream>>>(dY, lddy, X, ldx, M, N, K, pl.rows,
                                                                                                     part, part_b);
    MORIG_LAUNCH_CHECK("wgrad_tc_kernel");
    return 0;
}

extern "C" MORIG_API int morig_wgrad_f32(const float *dY, int32_t lddy, const float *X, int32_t ldx, int32_t M, int32_t N,
                                         int32_t K, const float *x_scale, const float *x_shift, float *dW, int32_t lddw,
                                         float *dbias, int32_t accumulate, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dY && X && dW && M > 0 && N > 0 && K > 0 && lddy >= N && ldx >= K && lddw >= K, "wgrad: bad argument");
    WgradPlan pl = wgrad_plan(M, N, K);
    const int splits = pl.splits;
    if (ws_bytes < (size_t)splits * ((size_t)N * K + N) * sizeof(float) || !ws) {
        set_error("wgrad: workspace too small");
        return MORIG_E_WORKSPACE;
    }
    float *part = reinterpret_cast<float *>(ws);
    float *part_b = part + (size_t)splits * N * K;
    if (pl.tc && !x_scale && !x_shift) {
        if (int rc = (pl.bkw == 256) ? launch_wgrad_tc<256>(dY, lddy, X, ldx, M, N, K, pl, part, dbias ? part_b : nullptr, stream)
                                     : launch_wgrad_tc<128>(dY, lddy, X, ldx, M, N, K, pl, part, dbias ? part_b : nullptr, stream))
            return rc;
    } else {
        if (pl.tc) {                                 // column affine on X: CUDA-core kernel with the tensor-core plan's slices
            pl.rows = ceil_div(pl.rows, WG_M) * WG_M;
        }
        const int rows = pl.rows;
        const int vec_y = (lddy % 4 == 0 && aligned16p(dY)) ? 1 : 0, vec_x = (ldx % 4 == 0 && aligned16p(X)) ? 1 : 0;
        if (wgrad_tile(N, K) == 64)
            wgrad_kernel<64><<<dim3(ceil_div(N, 64), ceil_div(K, 64), splits), WG_THREADS, 0, stream>>>(
                dY, lddy, X, ldx, M, N, K, x_scale, x_shift, rows, part, dbias ? part_b : nullptr, vec_y, vec_x);
        else
            wgrad_kernel<128><<<dim3(ceil_div(N, 128), ceil_div(K, 128), splits), WG_THREADS, 0, stream>>>(
                dY, lddy, X, ldx, M, N, K, x_scale, x_shift, rows, part, dbias ? part_b : nullptr, vec_y, vec_x);
        MORIG_LAUNCH_CHECK("wgrad_kernel");
    }
    wgrad_reduce_kernel<<<grid1d((int64_t)N * K + N, 256, 8), 256, 0, stream>>>(part, dbias ? part_b : nullptr, splits, N, K, dW,
                                                                               lddw, dbias, accumulate);
    MORIG_LAUNCH_CHECK("wgrad_reduce_kernel");
    return 0;
}

static int stats_chunks(int R) {
    int chunks = sm_count() * 2;
    const int max_by_rows = ceil_div(R, 64);
    if (chunks > max_by_rows) chunks = max_by_rows;
    return chunks < 1 ? 1 : chunks;
}

extern "C" MORIG_API size_t morig_colstats_workspace(int32_t R, int32_t C) {
    return (size_t)stats_chunks(R) * 2 * C * sizeof(double);
}

static int launch_colstats(const float *X, int ldx, const float *Y, int ldy, int R, int C, double *part, int &chunks,
                           cudaStream_t stream) {
    chunks = stats_chunks(R);
    int rows = ceil_div(R, chunks);
    rows = ceil_div(rows, 8) * 8;
    chunks = ceil_div(R, rows);
    colstats_kernel<<<dim3(ceil_div(C, 32), chunks), 256, 0, stream>>>(X, ldx, Y, ldy, R, C, rows, part);
    MORIG_LAUNCH_CHECK("colstats_kernel");
    return 0;
}

/* train-mode BatchNorm1d forward over the rows of x [R, C] */
extern "C" MORIG_API int morig_bn_train_fwd(const float *x, int32_t ldx, int32_t R, int32_t C, const float *gamma,
                                            const float *beta, float eps, float momentum, float *running_mean,
                                            float *running_var, float *mean, float *invstd, float *scale, float *shift,
                                            float *y, int32_t ldy, float *y_amax, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && mean && invstd && scale && shift && R > 0 && C > 0 && ldx >= C, "bn_train_fwd: bad argument");
    if (!ws || ws_bytes < morig_colstats_workspace(R, C)) { set_error("bn_train_fwd: workspace too small"); return MORIG_E_WORKSPACE; }
    int chunks = 0;
    if (int rc = launch_colstats(x, ldx, nullptr, 0, R, C, reinterpret_cast<double *>(ws), chunks, stream)) return rc;
    bn_finalize_fwd_kernel<<<ceil_div(C * 32, 256), 256, 0, stream>>>(reinterpret_cast<double *>(ws), chunks, R, C, gamma, beta, eps,
                                                                momentum, running_mean, running_var, mean, invstd, scale,
                                                                shift);
    MORIG_LAUNCH_CHECK("bn_finalize_fwd_kernel");
    if (y) {
        MORIG_CHECK_ARG(ldy >= C, "bn_train_fwd: ldy < C");
        if (vec4_ok(C, {ldx, ldy}, {x, y, scale, shift}))
            bn_apply_kernel<1><<<grid1d((int64_t)R * C / 4, 256, 8), 256, 0, stream>>>(x, ldx, R, C, scale, shift, y, ldy, y_amax);
        else
            bn_apply_kernel<0><<<grid1d((int64_t)R * C, 256, 8), 256, 0, stream>>>(x, ldx, R, C, scale, shift, y, ldy, y_amax);
        MORIG_LAUNCH_CHECK("bn_apply_kernel");
    }
    return 0;
}

extern "C" MORIG_API int morig_col_affine(const float *x, int32_t ldx, int32_t R, int32_t C, const float *scale, const float *shift,
                                          float *y, int32_t ldy, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && y && scale && shift && R > 0 && C > 0 && ldx >= C && ldy >= C, "col_affine: bad argument");
    if (vec4_ok(C, {ldx, ldy}, {x, y, scale, shift}))
        bn_apply_kernel<1><<<grid1d((int64_t)R * C / 4, 256, 8), 256, 0, stream>>>(x, ldx, R, C, scale, shift, y, ldy, nullptr);
    else
        bn_apply_kernel<0><<<grid1d((int64_t)R * C, 256, 8), 256, 0, stream>>>(x, ldx, R, C, scale, shift, y, ldy, nullptr);
    MORIG_LAUNCH_CHECK("bn_apply_kernel");
    return 0;
}

/* backward of Linear -> [ReLU] -> BatchNorm(train) at the BatchNorm input x (= ReLU output): dz, dgamma, dbeta */
extern "C" MORIG_API int morig_bn_relu_bwd(const float *dy, int32_t lddy, const float *x, int32_t ldx, int32_t R, int32_t C,
                                           const float *gamma, const float *mean, const float *invstd, int32_t relu,
                                           float *dz, int32_t lddz, float *dgamma, float *dbeta, float *coef, float *dz_amax,
                                           void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dy && x && mean && invstd && dz && coef && R > 0 && C > 0 && lddy >= C && ldx >= C && lddz >= C,
                    "bn_relu_bwd: bad argument");
    if (!ws || ws_bytes < morig_colstats_workspace(R, C)) { set_error("bn_relu_bwd: workspace too small"); return MORIG_E_WORKSPACE; }
    int chunks = 0;
    if (int rc = launch_colstats(x, ldx, dy, lddy, R, C, reinterpret_cast<double *>(ws), chunks, stream)) return rc;
    bn_finalize_bwd_kernel<<<ceil_div(C * 32, 256), 256, 0, stream>>>(reinterpret_cast<double *>(ws), chunks, R, C, gamma, mean, invstd,
                                                                dgamma, dbeta, coef);
    MORIG_LAUNCH_CHECK("bn_finalize_bwd_kernel");
    if (vec4_ok(C, {lddy, ldx, lddz}, {dy, x, dz}))
        bn_relu_bwd_kernel<1><<<grid1d((int64_t)R * C / 4, 256, 8), 256, 0, stream>>>(dy, lddy, x, ldx, R, C, mean, invstd, coef, relu,
                                                                                    dz, lddz, dz_amax);
    else
        bn_relu_bwd_kernel<0><<<grid1d((int64_t)R * C, 256, 8), 256, 0, stream>>>(dy, lddy, x, ldx, R, C, mean, invstd, coef, relu, dz,
                                                                                 lddz, dz_amax);
    MORIG_LAUNCH_CHECK("bn_relu_bwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_relu_bwd(const float *dy, int32_t lddy, const float *y, int32_t ldy, int32_t R, int32_t C,
                                        float *dz, int32_t lddz, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dy && y && dz && R > 0 && C > 0, "relu_bwd: bad argument");
    relu_bwd_kernel<<<grid1d((int64_t)R * C, 256, 8), 256, 0, stream>>>(dy, lddy, y, ldy, R, C, dz, lddz);
    MORIG_LAUNCH_CHECK("relu_bwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_edge_gather_relu(const float *P, int32_t ldp, const float *Q, int32_t ldq, const int32_t *tgt,
                                                const int32_t *col, int32_t E, int32_t C, float *h, int32_t ldh,
                                                float *h_amax, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(P && Q && tgt && col && h && E > 0 && C > 0 && ldh >= C, "edge_gather_relu: bad argument");
    if (vec4_ok(C, {ldp, ldq, ldh}, {P, Q, h}))
        edge_gather_relu_kernel<1><<<grid1d((int64_t)E * C / 4, 256, 8), 256, 0, stream>>>(P, ldp, Q, ldq, tgt, col, E, C, h, ldh, h_amax);
    else
        edge_gather_relu_kernel<0><<<grid1d((int64_t)E * C, 256, 8), 256, 0, stream>>>(P, ldp, Q, ldq, tgt, col, E, C, h, ldh, h_amax);
    MORIG_LAUNCH_CHECK("edge_gather_relu_kernel");
    return 0;
}

/* dP [N, C] (overwritten), dQ [N, C] (overwritten) from dh [E, C] and the saved h */
extern "C" MORIG_API int morig_edge_gather_relu_bwd(const float *dh, int32_t lddh, const float *h, int32_t ldh,
                                                    const int32_t *rowptr, const int32_t *col, int32_t N, int32_t E,
                                                    int32_t C, float *dP, int32_t ldp, float *dQ, int32_t ldq,
                                                    void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dh && h && rowptr && col && dP && dQ && N > 0 && E > 0 && C > 0, "edge_gather_relu_bwd: bad argument");
    const bool vec = vec4_ok(C, {lddh, ldh, ldp, ldq}, {dh, h, dP, dQ});
    if (vec) edge_gather_bwd_p_kernel<1><<<grid1d((int64_t)N * C / 4, 256, 8), 256, 0, stream>>>(dh, lddh, h, ldh, rowptr, N, C, dP, ldp);
    else edge_gather_bwd_p_kernel<0><<<grid1d((int64_t)N * C, 256, 8), 256, 0, stream>>>(dh, lddh, h, ldh, rowptr, N, C, dP, ldp);
    MORIG_LAUNCH_CHECK("edge_gather_bwd_p_kernel");
    MORIG_CUDA(cudaMemset2DAsync(dQ, (size_t)ldq * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)N, stream));
    if (vec) edge_gather_bwd_q_kernel<1><<<grid1d((int64_t)E * C / 4, 256, 8), 256, 0, stream>>>(dh, lddh, h, ldh, col, E, C, dQ, ldq);
    else edge_gather_bwd_q_kernel<0><<<grid1d((int64_t)E * C, 256, 8), 256, 0, stream>>>(dh, lddh, h, ldh, col, E, C, dQ, ldq);
    MORIG_LAUNCH_CHECK("edge_gather_bwd_q_kernel");
    return 0;
}

extern "C" MORIG_API int morig_segmax_fwd(const float *y, int32_t ldy, const int32_t *ptr, int32_t S, int32_t C, float *out,
                                          int32_t ldo, int32_t *arg, int32_t lda, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(y && ptr && out && S > 0 && C > 0 && ldy >= C && ldo >= C, "segmax_fwd: bad argument");
    if (S >= 1024) {
        // many segments = the per-vertex aggregation over in-edges (short segments); few = the per-graph pooling
        const bool vec = vec4_ok(C, {ldy, ldo, arg ? lda : 4}, {y, out, arg});
        if (vec) segmax_short_kernel<1><<<grid1d((int64_t)S * C / 4, 256, 8), 256, 0, stream>>>(y, ldy, ptr, S, C, out, ldo, arg, lda);
        else segmax_short_kernel<0><<<grid1d((int64_t)S * C, 256, 8), 256, 0, stream>>>(y, ldy, ptr, S, C, out, ldo, arg, lda);
        MORIG_LAUNCH_CHECK("segmax_short_kernel");
        return 0;
    }
    const int gy = S < 65535 ? S : 65535;
    segmax_kernel<<<dim3(ceil_div(C, 32), gy), 256, 0, stream>>>(y, ldy, ptr, S, C, out, ldo, arg, lda);
    MORIG_LAUNCH_CHECK("segmax_kernel");
    return 0;
}

/* dy [R, C] is overwritten: zero, then dy[arg[s, c], c] = dout[s, c] */
extern "C" MORIG_API int morig_segmax_bwd(const float *dout, int32_t lddo, const int32_t *arg, int32_t lda, int32_t S, int32_t C,
                                          float *dy, int32_t lddy, int32_t R, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(dout && arg && dy && S > 0 && C > 0 && R > 0 && lddy >= C, "segmax_bwd: bad argument");
    MORIG_CUDA(cudaMemset2DAsync(dy, (size_t)lddy * sizeof(float), 0, (size_t)C * sizeof(float), (size_t)R, stream));
    segmax_bwd_kernel<<<grid1d((int64_t)S * C, 256, 8), 256, 0, stream>>>(dout, lddo, arg, lda, S, C, dy, lddy);
    MORIG_LAUNCH_CHECK("segmax_bwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_seg_ptr(const int32_t *keys, int32_t N, int32_t S, int32_t *ptr, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(keys && ptr && N > 0 && S > 0, "seg_ptr: bad argument");
    seg_ptr_kernel<<<ceil_div(N + 1, 256), 256, 0, stream>>>(keys, N, S, ptr);
    MORIG_LAUNCH_CHECK("seg_ptr_kernel");
    return 0;
}

extern "C" MORIG_API int morig_row_gather(const float *src, int32_t lds, const int32_t *idx, int32_t R, int32_t C, float *out,
                                          int32_t ldo, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(src && idx && out && R > 0 && C > 0, "row_gather: bad argument");
    row_gather_kernel<<<grid1d((int64_t)R * C, 256, 8), 256, 0, stream>>>(src, lds, idx, R, C, out, ldo);
    MORIG_LAUNCH_CHECK("row_gather_kernel");
    return 0;
}

extern "C" MORIG_API int morig_seg_sum(const float *x, int32_t ldx, const int32_t *ptr, int32_t S, int32_t C, float *out,
                                       int32_t ldo, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && ptr && out && S > 0 && C > 0, "seg_sum: bad argument");
    const int gy = S < 65535 ? S : 65535;
    seg_sum_kernel<<<dim3(ceil_div(C, 32), gy), 256, 0, stream>>>(x, ldx, ptr, S, C, out, ldo);
    MORIG_LAUNCH_CHECK("seg_sum_kernel");
    return 0;
}

extern "C" MORIG_API int morig_normalize_fwd(const float *x, int32_t ldx, int32_t R, int32_t C, float *y, int32_t ldy,
                                             void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && y && R > 0 && C > 0, "normalize_fwd: bad argument");
    normalize_fwd_kernel<<<(unsigned)ceil_div64((int64_t)R * 32, 256), 256, 0, stream>>>(x, ldx, R, C, y, ldy);
    MORIG_LAUNCH_CHECK("normalize_fwd_kernel");
    return 0;
}

extern "C" MORIG_API int morig_normalize_bwd(const float *x, int32_t ldx, const float *dy, int32_t lddy, int32_t R, int32_t C,
                                             float *dx, int32_t lddx, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(x && dy && dx && R > 0 && C > 0, "normalize_bwd: bad argument");
    normalize_bwd_kernel<<<(unsigned)ceil_div64((int64_t)R * 32, 256), 256, 0, stream>>>(x, ldx, dy, lddy, R, C, dx, lddx);
    MORIG_LAUNCH_CHECK("normalize_bwd_kernel");
    return 0;
}

static int attn_check(int N, int T, int HD, int d) {
    MORIG_CHECK_ARG(N > 0 && T >= 1 && T <= ATTN_MAX_T, "attn_cls: T=%d unsupported (1..%d)", T, ATTN_MAX_T);
    MORIG_CHECK_ARG(HD % 32 == 0 && HD / 32 <= ATTN_MAX_CPL && d > 0 && HD % d == 0 && d % (HD / 32) == 0,
                    "attn_cls: HD=%d d=%d unsupported", HD, d);
    const int lph = d / (HD / 32);
    MORIG_CHECK_ARG((lph & (lph - 1)) == 0 && lph <= 32, "attn_cls: lanes per head must be a power of two");
    return 0;
}

static int attn_blocks(int N) {
    const int blocks = ceil_div(N, 8), cap = sm_count() * 4;
    return blocks < cap ? blocks : cap;
}

extern "C" MORIG_API int morig_attn_cls_fwd(const float *q0, const float *kc, const float *vc, const float *Kx, const float *Vx,
                                            int32_t N, int32_t T, int32_t HD, int32_t d, float *out, float *att,
                                            void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(q0 && kc && vc && Kx && Vx && out && att, "attn_cls_fwd: null operand");
    if (int rc = attn_check(N, T, HD, d)) return rc;
    attn_cls_fwd_kernel<<<attn_blocks(N), 256, 0, stream>>>(q0, kc, vc, Kx, Vx, N, T, HD, d, out, att);
    MORIG_LAUNCH_CHECK("attn_cls_fwd_kernel");
    return 0;
}

extern "C" MORIG_API size_t morig_attn_cls_bwd_workspace(int32_t N, int32_t HD) {
    return (size_t)attn_blocks(N) * 3 * HD * sizeof(float);
}

extern "C" MORIG_API int morig_attn_cls_bwd(const float *q0, const float *kc, const float *vc, const float *Kx, const float *Vx,
                                            const float *att, const float *dout, int32_t N, int32_t T, int32_t HD, int32_t d,
                                            float *dq0, float *dkc, float *dvc, float *dKx, float *dVx, void *ws,
                                            size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(q0 && kc && vc && Kx && Vx && att && dout && dq0 && dkc && dvc && dKx && dVx, "attn_cls_bwd: null operand");
    if (int rc = attn_check(N, T, HD, d)) return rc;
    if (!ws || ws_bytes < morig_attn_cls_bwd_workspace(N, HD)) { set_error("attn_cls_bwd: workspace too small"); return MORIG_E_WORKSPACE; }
    const int blocks = attn_blocks(N);
    attn_cls_bwd_kernel<<<blocks, 256, (size_t)8 * 3 * HD * sizeof(float), stream>>>(q0, kc, vc, Kx, Vx, att, dout, N, T, HD, d, dKx,
                                                                                   dVx, reinterpret_cast<float *>(ws));
    MORIG_LAUNCH_CHECK("attn_cls_bwd_kernel");
    attn_cls_reduce_kernel<<<ceil_div(3 * HD, 128), 128, 0, stream>>>(reinterpret_cast<float *>(ws), blocks, HD, dq0, dkc, dvc);
    MORIG_LAUNCH_CHECK("attn_cls_reduce_kernel");
    return 0;
}
