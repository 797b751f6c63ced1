// Weight gradient on the tcgen05 tensor cores (training path):
//
//   dW[n, k] = sum_m dY[m, n] * X[m, k]          db[n] = sum_m dY[m, n]
//
// the reduction runs over the ROWS of both operands (E' edge rows or N vertex rows: 16 K - 1.4 M), so in tensor-core
// terms both operands are needed "K-major in m", i.e. transposed with respect to their row-major layout in HBM.  The
// producers do that transposition on the way into shared memory: a lane owns one COLUMN of dY (or X), reads it with
// coalesced 4-byte loads (a warp reads 128 contiguous bytes of one row), and every four consecutive rows become one
// 16-byte chunk of that column's 128-byte swizzled shared-memory row -- conflict free, because the 8 lanes of a store
// phase own 8 different rows.
//
// Arithmetic: 3xTF32 (kind::tf32 on hi = x with 13 mantissa bits cleared, lo = x - hi; A_hi B_lo + A_lo B_hi + A_hi B_hi,
// fp32 accumulation in TMEM) -- the same error class as the fp16-split kind of the forward engine (~2^-21 per product)
// but with the fp32 exponent range, so gradients need no range bookkeeping (no |max| pass over dY / X, no scales).
// The MMA runs at half the fp16 rate; the kernel is bound by it, not by the producers (3.5 instructions per element).
//
// One CTA owns a 128 (n) x BKW (k) tile of dW for a slice of the rows; slices are summed in a fixed order by
// wgrad_reduce_kernel (fp64), like the CUDA-core kernel's, so the result does not depend on the launch geometry's timing.
//   warps 0-3            producers of the dY^T image (32 columns each) + bias-gradient column sums, then the epilogue
//   warps 4-(4+BKW/32)   producers of the X^T image
//   warp  12             one elected lane issues the 12 tcgen05.mma per 32-row stage and commits stage / accumulator
#pragma once
#include "gemm_tc.cuh"

namespace morig {
namespace tcw {

using namespace tc;

constexpr int W_ROWS = 32;                   // rows (m) per stage: one 128-byte swizzled row of tf32 per operand column
constexpr int W_BN = 128;                    // dW rows (n) per CTA = TMEM lanes
constexpr int W_YWARPS = W_BN / 32;
constexpr int W_CONTROL = 12;
constexpr int W_THREADS = 512;
constexpr uint32_t W_AUX_FULL = 0, W_AUX_EMPTY = 64, W_AUX_ACC = 128, W_AUX_TMEM = 136;

template <int BKW> struct WCfg {
    static constexpr int XWARPS = BKW / 32;
    static constexpr int A_HALF = W_BN * 128, B_HALF = BKW * 128;
    static constexpr int STAGE_BYTES = 2 * A_HALF + 2 * B_HALF;              // 96 KB (BKW = 256) / 64 KB (BKW = 128)
    static constexpr int STAGES = (BKW == 256) ? 2 : 3;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 1024;          // + barriers / TMEM slot + alignment slack
    static_assert(W_YWARPS + XWARPS <= W_CONTROL, "producer warps collide with the control warp");
};

template <int BKW>
__global__ void __launch_bounds__(W_THREADS, 1) wgrad_tc_kernel(const float *__restrict__ dY, int lddy, const float *__restrict__ X,
                                                                int ldx, int M, int N, int K, int rows_per_split,
                                                                float *__restrict__ part, float *__restrict__ part_b) {
    using C = WCfg<BKW>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // swizzle atoms need 1024 B alignment
    uint8_t *smem = smem_raw + (base - raw_addr);
    uint8_t *aux = smem + C::STAGES * C::STAGE_BYTES;
    const uint32_t aux_addr = base + C::STAGES * C::STAGE_BYTES;
    auto bar_full = [&](int s) { return aux_addr + W_AUX_FULL + 8u * s; };
    auto bar_empty = [&](int s) { return aux_addr + W_AUX_EMPTY + 8u * s; };
    const uint32_t bar_acc = aux_addr + W_AUX_ACC;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aux + W_AUX_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * W_BN, k0 = blockIdx.y * BKW;
    const int m_begin = blockIdx.z * rows_per_split;
    const int m_end = min(M, m_begin + rows_per_split);
    const int n_stages = (m_end > m_begin) ? (m_end - m_begin + W_ROWS - 1) / W_ROWS : 0;

    if (warp == W_CONTROL) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; ++s) {
                mbar_init(bar_full(s), W_YWARPS + C::XWARPS);
                mbar_init(bar_empty(s), 1);
            }
            mbar_init(bar_acc, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(aux_addr + W_AUX_TMEM, BKW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    float bsum = 0.f;                                           // bias gradient of this lane's dY column (warps 0-3)
    if (warp < W_YWARPS + C::XWARPS) {
        // ================= producers: transposed hi / lo images =================
        const bool is_y = warp < W_YWARPS;
        const int lc = 32 * (is_y ? warp : warp - W_YWARPS) + lane;         // column inside the tile = shared-memory row
        const int gc = (is_y ? n0 : k0) + lc;
        const bool col_ok = gc < (is_y ? N : K);
        const float *src = (is_y ? dY : X) + (col_ok ? gc : 0);
        const size_t ld = (size_t)(is_y ? lddy : ldx);
        const uint32_t img_hi = is_y ? 0u : (uint32_t)(2 * C::A_HALF);
        const uint32_t img_lo = img_hi + (uint32_t)(is_y ? C::A_HALF : C::B_HALF);
        const uint32_t row_off = (uint32_t)lc * 128u;
        const uint32_t sw = (uint32_t)(lc & 7);
        float b0[W_ROWS], b1[W_ROWS];
        auto load = [&](float (&b)[W_ROWS], int it) {
            const int m0 = m_begin + it * W_ROWS;
#pragma unroll
            for (int r = 0; r < W_ROWS; ++r) {
                const int m = m0 + r;
                b[r] = (col_ok && m < m_end) ? src[(size_t)m * ld] : 0.f;
            }
        };
        auto store = [&](const float (&b)[W_ROWS], int it) {
            const int s = it % C::STAGES;
            mbar_wait(bar_empty(s), (uint32_t)(((it / C::STAGES) & 1) ^ 1));          // the MMAs of the previous ring turn retired
            uint8_t *st = smem + (size_t)s * C::STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < W_ROWS / 4; ++j) {
                const float4 v = make_float4(b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]);
                const float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
                const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                const uint32_t off = row_off + (((uint32_t)j ^ sw) << 4);
                *reinterpret_cast<float4 *>(st + img_hi + off) = h;
                *reinterpret_cast<float4 *>(st + img_lo + off) = l;
                if (is_y) bsum += (v.x + v.y) + (v.z + v.w);
            }
            fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(s));
        };
        if (n_stages > 0) load(b0, 0);
        for (int it = 0; it < n_stages; it += 2) {
            if (it + 1 < n_stages) load(b1, it + 1);
            store(b0, it);
            if (it + 1 < n_stages) {
                if (it + 2 < n_stages) load(b0, it + 2);
                store(b1, it + 1);
            }
        }
    } else if (warp == W_CONTROL) {
        // ================= MMA issue (one elected lane; the warp stays converged) =================
        const uint32_t idesc = make_idesc<KIND_TF32>(W_BN, BKW);
        for (int it = 0; it < n_stages; ++it) {
            const int s = it % C::STAGES;
            mbar_wait(bar_full(s), (uint32_t)((it / C::STAGES) & 1));
            tc_fence_after();
            const uint32_t a_hi = base + (uint32_t)s * C::STAGE_BYTES;
            const uint32_t lah = desc_lo(a_hi), lal = desc_lo(a_hi + C::A_HALF);
            const uint32_t lbh = desc_lo(a_hi + 2 * C::A_HALF), lbl = desc_lo(a_hi + 2 * C::A_HALF + C::B_HALF);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {                    // 8 tf32 = 32 bytes along the swizzled row
                    umma<KIND_TF32, 1>(tmem_base, desc64(lah + 2 * k), desc64(lbl + 2 * k), idesc, (it | k) != 0);
                    umma<KIND_TF32, 1>(tmem_base, desc64(lal + 2 * k), desc64(lbh + 2 * k), idesc, 1);
                    umma<KIND_TF32, 1>(tmem_base, desc64(lah + 2 * k), desc64(lbh + 2 * k), idesc, 1);
                }
                umma_commit<1>(bar_empty(s));                    // frees the stage when these MMAs retire
            }
            __syncwarp();
        }
        if (lane == 0 && n_stages > 0) umma_commit<1>(bar_acc);  // accumulator complete -> epilogue
    }

    if (warp < W_YWARPS) {
        // ================= epilogue: this CTA's partial tile -> workspace =================
        if (n_stages > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
        }
        const int n = n0 + 32 * warp + lane;                     // TMEM lane = dW row
        float *out = part + (size_t)blockIdx.z * N * K + (size_t)(n < N ? n : 0) * K + k0;
        const uint32_t tlane = tmem_base + ((uint32_t)(32 * warp) << 16);
#pragma unroll 1
        for (int c = 0; c < BKW; c += 32) {
            uint32_t v[32];
            uint32_t (&va)[16] = reinterpret_cast<uint32_t (&)[16]>(v[0]);
            uint32_t (&vb)[16] = reinterpret_cast<uint32_t (&)[16]>(v[16]);
            if (n_stages > 0) {
                tmem_ld16_issue(tlane + (uint32_t)c, va);
                tmem_ld16_issue(tlane + (uint32_t)c + 16u, vb);
                tmem_ld_wait(va, vb);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0u;          // a slice without rows contributes zeros
            }
            if (n < N) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (k0 + c + i < K) out[c + i] = __uint_as_float(v[i]);
            }
        }
        if (part_b && blockIdx.y == 0 && n < N) part_b[(size_t)blockIdx.z * N + n] = bsum;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_CONTROL) tmem_dealloc<1>(tmem_base, BKW);
}

}  // namespace tcw
}  // namespace morig
