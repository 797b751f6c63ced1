// FP32 FFMA tile engine shared by the vertex-side dense layers and the fused EdgeConv branch.
//
//   C tile = BM x BN per 256-thread CTA, BK = 16, register-prefetch double buffering,
//   8 x (BN/16) micro-tile per thread.
//   A operand:  plain rows of an activation matrix          (AMODE_PLAIN)
//               relu(P[tgt[e]] + Q[col[e]]) gathered per CSR slot e (AMODE_GATHER)  -- the per-edge
//               first layer of the EdgeConv MLP after layer-0 factorisation
//   epilogue:   bias (+ per-graph row bias) -> ReLU -> BatchNorm affine, then
//               EPI_STORE   write C and/or per-graph column max (ordered atomics)
//               EPI_SEGMAX  segmented max over the CSR target of each row (the EdgeConv aggregate)
#pragma once
#include "common.cuh"

namespace morig {

enum { AMODE_PLAIN = 0, AMODE_GATHER = 1 };
enum { EPI_STORE = 0, EPI_SEGMAX = 1 };

struct GemmP {
    // plain A
    const float *A; int lda; int a_vec;
    // gathered A (edge mode)
    const float *P, *Q; int ldpq;
    const int32_t *rowptr, *col, *tgt;
    int n_vtx_frame;            // vertices per key-frame block (edge mode)
    // B
    const float *W; int ldw;
    // epilogue
    const float *bias, *scale, *shift;
    const float *rowbias; int ldrb;
    const int32_t *batch; int n_vtx, n_graphs;
    float *C; int ldc; int c_vec;
    float *pool; int ldpool;
    int M, N, K, relu;
    // dynamic range bookkeeping of the fp16-split tensor-core engine (gemm_tc.cuh): every kernel that stores
    // activations max-reduces |value| into *amax_out; an fp16 consumer derives its operand scale from *amax_in
    const float *amax_in; float *amax_out;
    float w_inv;                // 1 / (power-of-two scale baked into the fp16 weight image)
    const float *w_inv_dev;     // ... or a device scalar holding it (image packed on the device: morig_pack_tc_f16)
};

constexpr int GEMM_THREADS = 256;
constexpr int BK = 16;

template <int BM, int BN>
struct GemmSmem {
    static constexpr int AS_LD = BM + 4;
    static constexpr int BS_LD = BN;
    static constexpr int PIPE_FLOATS = 2 * BK * AS_LD + 2 * BK * BS_LD;
    static constexpr int CS_LD = BN + 4;
    static constexpr int CS_FLOATS = BM * CS_LD;
    static constexpr int POOL_FLOATS = 16 * BN;
};

template <int BM, int BN, int EPI>
constexpr size_t gemm_smem_bytes() {
    using S = GemmSmem<BM, BN>;
    size_t f = S::PIPE_FLOATS;
    if (EPI == EPI_SEGMAX && (size_t)S::CS_FLOATS > f) f = S::CS_FLOATS;
    if (EPI == EPI_STORE && (size_t)S::POOL_FLOATS > f) f = S::POOL_FLOATS;
    return f * sizeof(float) + (EPI == EPI_SEGMAX ? BM * sizeof(int32_t) : 0);
}

template <int BM, int BN, int AMODE, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_simt_kernel(const GemmP p) {
    using S = GemmSmem<BM, BN>;
    constexpr int TM = BM / 16, TN = BN / 16;       // micro-tile
    constexpr int A_LD_PER_THREAD = BM / 64;        // float4 rows per thread (rows tid>>2 (+64))
    constexpr int B_LD_PER_THREAD = (BK * BN / 4) / GEMM_THREADS;
    static_assert(BM % 64 == 0 && (BN == 64 || BN == 128), "tile shape");
    static_assert(B_LD_PER_THREAD >= 1, "B tile too small");

    extern __shared__ __align__(16) float smem[];
    float *As = smem;                               // [2][BK][AS_LD]
    float *Bs = smem + 2 * BK * S::AS_LD;           // [2][BK][BS_LD]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int frame = (AMODE == AMODE_GATHER) ? blockIdx.z : 0;

    pdl_trigger();
    pdl_wait();
    int M = p.M;
    if (AMODE == AMODE_GATHER) {
        M = p.rowptr[p.n_vtx_frame];                // E' (device side)
        if (m0 >= M) return;
    }

    // ---- per-thread A source rows -----------------------------------------------------------
    const int a_kq = (tid & 3) * 4;
    const float *a_src0[A_LD_PER_THREAD];
    const float *a_src1[A_LD_PER_THREAD];
    bool a_ok[A_LD_PER_THREAD];
#pragma unroll
    for (int l = 0; l < A_LD_PER_THREAD; ++l) {
        const int row = (tid >> 2) + l * 64;
        const int r = m0 + row;
        a_ok[l] = r < M;
        if (AMODE == AMODE_GATHER) {
            int i = 0, j = 0;
            if (a_ok[l]) { i = p.tgt[r]; j = p.col[r]; }
            const size_t fb = (size_t)frame * p.n_vtx_frame;
            a_src0[l] = p.P + (fb + i) * (size_t)p.ldpq;
            a_src1[l] = p.Q + (fb + j) * (size_t)p.ldpq;
        } else {
            a_src0[l] = p.A + (size_t)(a_ok[l] ? r : 0) * p.lda;
            a_src1[l] = nullptr;
        }
    }
    // ---- per-thread B source ------------------------------------------------------------------
    constexpr int B_COLS4 = BN / 4;                 // float4 per B row
    const int b_n4 = (tid % B_COLS4) * 4;
    const int b_k = tid / B_COLS4;                  // + l * (GEMM_THREADS / B_COLS4)
    constexpr int B_KSTEP = GEMM_THREADS / B_COLS4;

    float4 a_reg[A_LD_PER_THREAD];
    float4 b_reg[B_LD_PER_THREAD];

    auto load_global = [&](int k0) {
#pragma unroll
        for (int l = 0; l < A_LD_PER_THREAD; ++l) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = k0 + a_kq;
            if (a_ok[l]) {
                if (AMODE == AMODE_GATHER) {
                    if (k < p.K) {
                        const float4 pv = *reinterpret_cast<const float4 *>(a_src0[l] + k);
                        const float4 qv = *reinterpret_cast<const float4 *>(a_src1[l] + k);
                        v.x = fmaxf(pv.x + qv.x, 0.f); v.y = fmaxf(pv.y + qv.y, 0.f);
                        v.z = fmaxf(pv.z + qv.z, 0.f); v.w = fmaxf(pv.w + qv.w, 0.f);
                    }
                } else if (p.a_vec) {
                    if (k < p.K) v = *reinterpret_cast<const float4 *>(a_src0[l] + k);
                } else {
                    if (k + 0 < p.K) v.x = a_src0[l][k + 0];
                    if (k + 1 < p.K) v.y = a_src0[l][k + 1];
                    if (k + 2 < p.K) v.z = a_src0[l][k + 2];
                    if (k + 3 < p.K) v.w = a_src0[l][k + 3];
                }
            }
            a_reg[l] = v;
        }
#pragma unroll
        for (int l = 0; l < B_LD_PER_THREAD; ++l) {
            const int k = k0 + b_k + l * B_KSTEP;
            const int n = n0 + b_n4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < p.K && n < p.ldw) v = *reinterpret_cast<const float4 *>(p.W + (size_t)k * p.ldw + n);
            b_reg[l] = v;
        }
    };
    auto store_smem = [&](int buf) {
        float *as = As + buf * BK * S::AS_LD;
        float *bs = Bs + buf * BK * S::BS_LD;
#pragma unroll
        for (int l = 0; l < A_LD_PER_THREAD; ++l) {
            const int row = (tid >> 2) + l * 64;
            as[(a_kq + 0) * S::AS_LD + row] = a_reg[l].x;
            as[(a_kq + 1) * S::AS_LD + row] = a_reg[l].y;
            as[(a_kq + 2) * S::AS_LD + row] = a_reg[l].z;
            as[(a_kq + 3) * S::AS_LD + row] = a_reg[l].w;
        }
#pragma unroll
        for (int l = 0; l < B_LD_PER_THREAD; ++l) {
            const int k = b_k + l * B_KSTEP;
            *reinterpret_cast<float4 *>(bs + k * S::BS_LD + b_n4) = b_reg[l];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load_global(0);
    store_smem(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_global((kt + 1) * BK);
        const float *as = As + (kt & 1) * BK * S::AS_LD;
        const float *bs = Bs + (kt & 1) * BK * S::BS_LD;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4 *>(as + k * S::AS_LD + h * (BM / 2) + ty * 4);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4 *>(bs + k * S::BS_LD + h * (BN / 2) + tx * 4);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_smem((kt + 1) & 1);
        __syncthreads();
    }

    // ---- epilogue ---------------------------------------------------------------------------------
    // local row (i) -> tile row; local col (j) -> tile col
    auto tile_row = [&](int i) { return (i >> 2) * (BM / 2) + ty * 4 + (i & 3); };
    auto tile_col = [&](int j) { return (j >> 2) * (BN / 2) + tx * 4 + (j & 3); };

    float cb[TN], cs[TN], ct[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int n = n0 + tile_col(j);
        const bool ok = n < p.N;
        cb[j] = (ok && p.bias) ? p.bias[n] : 0.f;
        cs[j] = (ok && p.scale) ? p.scale[n] : 1.f;
        ct[j] = (ok && p.shift) ? p.shift[n] : 0.f;
    }

    if (EPI == EPI_STORE) {
        const bool pooling = p.pool != nullptr;
        int g_first = 0, g_last = 0;
        if (p.batch) {
            const int r_first = m0, r_last = min(m0 + BM, M) - 1;
            g_first = (r_first / p.n_vtx) * p.n_graphs + p.batch[r_first % p.n_vtx];
            g_last = (r_last / p.n_vtx) * p.n_graphs + p.batch[r_last % p.n_vtx];
        }
        const bool uniform = g_first == g_last;
        float cmax[TN];
        float am = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) cmax[j] = neg_inf();
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = m0 + tile_row(i);
            if (r >= M) continue;
            int g = g_first;
            if (p.batch && !uniform) g = (r / p.n_vtx) * p.n_graphs + p.batch[r % p.n_vtx];
            float v[TN];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int n = n0 + tile_col(j);
                float x = acc[i][j] + cb[j];
                if (p.rowbias && n < p.N) x += p.rowbias[(size_t)g * p.ldrb + n];
                if (p.relu) x = fmaxf(x, 0.f);
                x = fmaf(x, cs[j], ct[j]);
                v[j] = x;
                if (n < p.N) am = fmaxf(am, fabsf(x));
                if (pooling) {
                    if (uniform) cmax[j] = fmaxf(cmax[j], x);
                    else if (n < p.N) atomic_max_f32(p.pool + (size_t)g * p.ldpool + n, x);
                }
            }
            if (p.C) {
#pragma unroll
                for (int h = 0; h < TN / 4; ++h) {
                    const int n = n0 + tile_col(h * 4);
                    float *dst = p.C + (size_t)r * p.ldc + n;
                    if (p.c_vec && n + 3 < p.N) {
                        *reinterpret_cast<float4 *>(dst) = make_float4(v[h * 4], v[h * 4 + 1], v[h * 4 + 2], v[h * 4 + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (n + q < p.N) dst[q] = v[h * 4 + q];
                    }
                }
            }
        }
        if (p.C) amax_commit(p.amax_out, am);
        if (pooling && uniform) {                    // CTA-wide column max, one atomic per column
            float *red = smem;                       // [16][BN], aliases the (drained) pipeline buffers
            __syncthreads();
#pragma unroll
            for (int j = 0; j < TN; ++j) red[ty * BN + tile_col(j)] = cmax[j];
            __syncthreads();
            for (int c = tid; c < BN; c += GEMM_THREADS) {
                float m = red[c];
#pragma unroll
                for (int y = 1; y < 16; ++y) m = fmaxf(m, red[y * BN + c]);
                const int n = n0 + c;
                if (n < p.N && m > neg_inf()) atomic_max_f32(p.pool + (size_t)g_first * p.ldpool + n, m);
            }
        }
    } else {                                         // EPI_SEGMAX
        float *Cs = smem;                            // [BM][CS_LD], aliases the pipeline buffers
        int32_t *s_tgt = reinterpret_cast<int32_t *>(smem + (S::CS_FLOATS > S::PIPE_FLOATS ? S::CS_FLOATS : S::PIPE_FLOATS));
        __syncthreads();
        for (int r = tid; r < BM; r += GEMM_THREADS) s_tgt[r] = (m0 + r < M) ? p.tgt[m0 + r] : -1;
        float am = 0.f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int row = tile_row(i);
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                float4 v;
                float *vv = reinterpret_cast<float *>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int j = h * 4 + q;
                    float x = acc[i][j] + cb[j];
                    x = fmaxf(x, 0.f);               // second ReLU of the edge MLP
                    vv[q] = fmaf(x, cs[j], ct[j]);   // BatchNorm #2 BEFORE the max (scale may be < 0)
                    if (m0 + row < M && n0 + tile_col(j) < p.N) am = fmaxf(am, fabsf(vv[q]));
                }
                *reinterpret_cast<float4 *>(Cs + row * S::CS_LD + tile_col(h * 4)) = v;
            }
        }
        amax_commit(p.amax_out, am);
        __syncthreads();
        constexpr int PARTS = GEMM_THREADS / BN;
        constexpr int ROWS_PER = BM / PARTS;
        const int c = tid % BN, part = tid / BN;
        const int n = n0 + c;
        if (n < p.N) {
            const int ra = part * ROWS_PER, rb = ra + ROWS_PER;
            const size_t fb = (size_t)frame * p.n_vtx_frame;
            int cur = -1;
            float m = neg_inf();
            auto flush = [&]() {
                const int lo = p.rowptr[cur], hi = p.rowptr[cur + 1];
                float *dst = p.C + (fb + cur) * (size_t)p.ldc + n;
                if (lo >= m0 + ra && hi <= m0 + rb) *dst = m;
                else atomic_max_f32(dst, m);
            };
            for (int r = ra; r < rb; ++r) {
                const int t = s_tgt[r];
                if (t < 0) break;
                if (t != cur) {
                    if (cur >= 0) flush();
                    cur = t;
                    m = neg_inf();
                }
                m = fmaxf(m, Cs[r * S::CS_LD + c]);
            }
            if (cur >= 0) flush();
        }
    }
}

}  // namespace morig
