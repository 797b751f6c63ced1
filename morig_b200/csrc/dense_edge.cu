// Host launchers for the tile engines (vertex dense layers, fused EdgeConv branch) and the narrow-channel EdgeConv
// branch kernel (H <= 32: warp per 32 CSR slots, mma.sync split-fp16, in-register segmented max).
#include <stdlib.h>
#include <cuda.h>
#include "gemm_tc.cuh"

namespace morig {

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- narrow EdgeConv branch (H = 16 / 32) --------------------------------------------------------
// Warp-level tensor-core kernel for the layers that are too small for the tcgen05 tile engine (gcu_1's x branch and
// every pos branch of the joint / mask networks: 2 E H^2 FLOPs on 16- or 32-wide rows, bound by the gathers and the
// segmented max, not by the MMAs).  One warp owns a tile of 32 consecutive CSR slots (edges sorted by target) and
//   * gathers relu(P[tgt] + Q[col]) straight into mma.sync A fragments (no shared-memory staging): lane (g, t) owns
//     the four edges 4g..4g+3 of the tile and, of each row, the H/4 contiguous channels [t H/4, (t+1) H/4) -- the
//     contraction index is permuted identically on the weight side, so every global load is a full 16-byte chunk;
//   * multiplies by the H x H second Linear with split-fp16 error compensation, the same arithmetic as the tcgen05
//     engine's fp16 kind (hi/lo fp16 splits of both operands scaled by powers of two -- the activation scale from
//     the tracked max |PQ|, the weight scale from max |W1| -- A_lo B_hi + A_hi B_lo + A_hi B_hi, fp32 accumulate;
//     mma.sync m16n8k16: the legacy TF32 shape runs at a fraction of this rate on sm_100); the split weight fragments
//     live in registers for the whole kernel;
//   * applies bias -> ReLU -> BatchNorm affine and reduces with a segmented max scan over the tile: four edges in
//     registers, then three shuffle steps across the eight lane groups; segment tails store (segments cut by the
//     tile boundary merge with the ordered-int atomic max, so the result is exact and order independent).
__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int EDGE_MMA_THREADS = 128;
constexpr int EDGE_BATCH_MAX = 4;
// several branches of the same width on the same graph (e.g. the pos branches of the three GCUs of a GCNRig, which
// all read the one pqpos buffer) share a launch: blockIdx.y selects the branch
struct EdgeBatch { morig_edge_desc d[EDGE_BATCH_MAX]; };
// TMA variant: one tensor map per branch over its PQ buffer ([rows, ldpq] fp32, box = H columns x 1 row, used with
// tile::gather4: four arbitrary rows per instruction)
struct alignas(64) EdgeMaps { CUtensorMap m[EDGE_BATCH_MAX]; };

// cp.async.bulk.tensor ... tile::gather4: rows r0..r3 of the 2-D tensor, columns [c0, c0 + box) -> 4 consecutive smem rows
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}

// TMA = false: every lane gathers its own P / Q chunks with 16-byte loads straight into MMA fragments.
// TMA = true : the neighbour rows Q[col[e]] of a tile are staged in shared memory by the TMA engine (tile::gather4,
//              4 rows per instruction, issued by the 8 group-leader lanes, completion on a per-warp mbarrier); two
//              stages per warp, so the gather of tile i+1 is in flight while tile i is multiplied.
template <int H, bool TMA>
__global__ void __launch_bounds__(EDGE_MMA_THREADS) edge_mma_kernel(const __grid_constant__ EdgeBatch batch,
                                                                    const __grid_constant__ EdgeMaps maps) {
    const morig_edge_desc &d = batch.d[blockIdx.y];
    constexpr int KT = H / 16, NT = H / 8;          // k-tiles and n-tiles of the m16n8k16 shape
    constexpr int KPL = H / 4;                      // contraction values of one row held by one lane
    constexpr int CPL = 2 * NT;                     // output columns held by one lane: 8 nt + 2 t + {0, 1}
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ float2 s_bias[H / 2], s_scale[H / 2], s_shift[H / 2];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    pdl_trigger();
    // the branch's weights and per-channel constants are parameters (packed before the forward started): staging them
    // overlaps the predecessor's tail; pdl_wait() comes before the first read of data produced inside the forward
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        reinterpret_cast<float *>(s_bias)[i] = d.b1[i];
        reinterpret_cast<float *>(s_scale)[i] = d.scale[i];
        reinterpret_cast<float *>(s_shift)[i] = d.shift[i];
    }
    // weight fragments: the four contraction slots {2t, 2t+1, 2t+8, 2t+9} of k-tile kt read the physical indices
    // KPL t + 4 kt + {0, 1, 2, 3} (the activation fragments use the same permutation)
    float wv[KT][NT][4];
    float wmax = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                wv[kt][nt][s4] = d.W1[(size_t)(KPL * t + 4 * kt + s4) * d.ldw + 8 * nt + g];
                wmax = fmaxf(wmax, fabsf(wv[kt][nt][s4]));
            }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(FULL, wmax, off));   // the warp holds all of W1
    // power-of-two operand scales: max|W1| * ws in [2^14, 2^15); relu(P + Q) * as < 2^15
    auto pow2 = [](int sh) { sh = sh < -100 ? -100 : (sh > 100 ? 100 : sh); return __uint_as_float((uint32_t)(sh + 127) << 23); };
    auto expo = [](float x) { int e = (int)((__float_as_uint(x) >> 23) & 0xffu); return (e == 0 || e == 255) ? 127 : e; };
    pdl_wait();
    const int w_sh = 14 - (expo(wmax) - 127);
    const int a_sh = 14 - (expo(d.pq_amax ? *d.pq_amax : 1.f) - 126);
    const float w_scale = pow2(w_sh), a_scale = pow2(a_sh);
    const float inv = pow2(-w_sh) * pow2(-a_sh);
    uint32_t bh[KT][NT][2], bl[KT][NT][2];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const float w0 = wv[kt][nt][2 * s2] * w_scale, w1 = wv[kt][nt][2 * s2 + 1] * w_scale;
                const __half2 hi = __floats2half2_rn(w0, w1);
                const float2 hf = __half22float2(hi);
                const __half2 lo = __floats2half2_rn(w0 - hf.x, w1 - hf.y);
                bh[kt][nt][s2] = *reinterpret_cast<const uint32_t *>(&hi);
                bl[kt][nt][s2] = *reinterpret_cast<const uint32_t *>(&lo);
            }
    __syncthreads();

    const int N = d.N;
    const int Ep = d.rowptr[N];                     // real edge count (E' lives on the device only)
    const int n_tiles = (Ep + 31) >> 5;
    const long long total = (long long)n_tiles * d.n_frames;
    const int warps_total = gridDim.x * (EDGE_MMA_THREADS / 32);
    const bool vec2 = ((d.ldo | d.out_off) & 1) == 0 && (reinterpret_cast<uintptr_t>(d.out) & 7u) == 0;
    const float *Pb = d.PQ + d.p_off + KPL * t;
    const float *Qb = d.PQ + d.q_off + KPL * t;
    float am = 0.f;

    // indices of a tile: this lane's four edges (target = segment key, source), plus the keys of the slots just before
    // and after the tile (lanes 0 / 31) that tell whether its first / last segment is cut by the tile boundary.
    // They are fetched one tile ahead, so the gathers of a tile never wait for its index loads.
    auto load_idx = [&](long long id, int (&key)[4], int (&cj)[4], int &ext) {
        ext = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) { key[j] = -1; cj[j] = 0; }
        if (id >= total) return;
        const int s0 = (int)(id % n_tiles) * 32;
        const int e0 = s0 + 4 * g;
        if (e0 + 3 < Ep) {
            const int4 k4 = *reinterpret_cast<const int4 *>(d.tgt + e0);
            const int4 c4 = *reinterpret_cast<const int4 *>(d.col + e0);
            key[0] = k4.x; key[1] = k4.y; key[2] = k4.z; key[3] = k4.w;
            cj[0] = c4.x; cj[1] = c4.y; cj[2] = c4.z; cj[3] = c4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (e0 + j < Ep) { key[j] = d.tgt[e0 + j]; cj[j] = d.col[e0 + j]; }
            }
        }
        if (lane == 0) ext = s0 > 0 ? d.tgt[s0 - 1] : -2;
        if (lane == 31 && s0 + 32 < Ep) ext = d.tgt[s0 + 32];
    };
    long long id = (long long)blockIdx.x * (EDGE_MMA_THREADS / 32) + (threadIdx.x >> 5);
    int nkey[4], ncj[4], next_;
    load_idx(id, nkey, ncj, next_);
    // TMA staging: [warp][stage][P | Q][32 edges][H] floats + one mbarrier per (warp, stage)
    extern __shared__ __align__(128) uint8_t tma_smem[];
    constexpr uint32_t ROW_BYTES = H * 4, STAGE_BYTES = 32 * ROW_BYTES;      // the tile's 32 neighbour rows Q[col[e]]
    const int wib = threadIdx.x >> 5;
    const uint32_t stage0 = tc::smem_u32(tma_smem) + (uint32_t)wib * 2u * STAGE_BYTES;
    const uint32_t bar0 = tc::smem_u32(tma_smem) + (uint32_t)(EDGE_MMA_THREADS / 32) * 2u * STAGE_BYTES + (uint32_t)wib * 16u;
    auto issue = [&](long long tid_, const int (&k4)[4], const int (&c4)[4], int st) {
        // group leaders (t == 0) each fetch the neighbour rows of their four edges; lane 0 announces the bytes.
        // (The target rows P[tgt[e]] stay on the per-lane load path: consecutive edges share their target, so those
        //  loads hit L1, which the TMA engine bypasses -- measured: staging P too costs 40 % more time.)
        const uint32_t bar = bar0 + 8u * st;
        (void)k4;
        if (lane == 0) tc::mbar_arrive_expect_tx(bar, STAGE_BYTES);
        __syncwarp();
        if (t == 0) {
            const int fbr = (int)(tid_ / n_tiles) * N;
            const uint32_t dst = stage0 + (uint32_t)st * STAGE_BYTES + (uint32_t)(4 * g) * ROW_BYTES;
            tma_gather4(dst, &maps.m[blockIdx.y], d.q_off, fbr + c4[0], fbr + c4[1], fbr + c4[2], fbr + c4[3], bar);
        }
    };
    if (TMA) {
        if (lane == 0) { tc::mbar_init(bar0, 1); tc::mbar_init(bar0 + 8u, 1); tc::fence_barrier_init(); }
        __syncwarp();
        if (id < total) issue(id, nkey, ncj, 0);
    }
    int it = 0;
    for (; id < total; id += warps_total, ++it) {
        const int f = (int)(id / n_tiles);
        const size_t fb = (size_t)f * N;
        int key[4], cj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { key[j] = nkey[j]; cj[j] = ncj[j]; }
        const int ext = next_;
        load_idx(id + warps_total, nkey, ncj, next_);
        // ---- gather: x[j] = relu(P[tgt] + Q[col]), this lane's KPL channels ----
        float x[4][KPL];
        if (TMA) {
            const int st = it & 1;
            float4 pa[4][KPL / 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                 // target rows: per-lane loads (L1 hits inside a segment)
                const float *pp = Pb + (fb + (size_t)max(key[j], 0)) * d.ldpq;
#pragma unroll
                for (int v = 0; v < KPL / 4; ++v) pa[j][v] = *reinterpret_cast<const float4 *>(pp + 4 * v);
            }
            tc::mbar_wait(bar0 + 8u * st, (uint32_t)((it >> 1) & 1));
            const uint8_t *sp = tma_smem + ((size_t)wib * 2 + st) * STAGE_BYTES + (size_t)(4 * g) * ROW_BYTES + (size_t)KPL * t * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int v = 0; v < KPL / 4; ++v) {
                    const float4 a = pa[j][v];
                    const float4 b = *reinterpret_cast<const float4 *>(sp + j * ROW_BYTES + 16 * v);
                    x[j][4 * v + 0] = fmaxf(a.x + b.x, 0.f) * a_scale;
                    x[j][4 * v + 1] = fmaxf(a.y + b.y, 0.f) * a_scale;
                    x[j][4 * v + 2] = fmaxf(a.z + b.z, 0.f) * a_scale;
                    x[j][4 * v + 3] = fmaxf(a.w + b.w, 0.f) * a_scale;
                }
            }
            __syncwarp();                               // every lane has read stage st^1's previous contents long ago;
            if (id + warps_total < total) issue(id + warps_total, nkey, ncj, st ^ 1);   // refill the other stage now
        } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float *pp = Pb + (fb + (size_t)max(key[j], 0)) * d.ldpq;
            const float *qq = Qb + (fb + (size_t)cj[j]) * d.ldpq;
#pragma unroll
            for (int v = 0; v < KPL / 4; ++v) {
                const float4 a = *reinterpret_cast<const float4 *>(pp + 4 * v);
                const float4 b = *reinterpret_cast<const float4 *>(qq + 4 * v);
                x[j][4 * v + 0] = fmaxf(a.x + b.x, 0.f) * a_scale;
                x[j][4 * v + 1] = fmaxf(a.y + b.y, 0.f) * a_scale;
                x[j][4 * v + 2] = fmaxf(a.z + b.z, 0.f) * a_scale;
                x[j][4 * v + 3] = fmaxf(a.w + b.w, 0.f) * a_scale;
            }
        }
        }
        // ---- segment structure of the tile (identical in the four lanes of a group) ----
        const int key_prev = __shfl_up_sync(FULL, key[3], 4);
        const int key_next = __shfl_down_sync(FULL, key[0], 4);
        bool head[4], tail[4];
        head[0] = (g == 0) || key[0] != key_prev;
        tail[3] = (g == 7) || key[3] != key_next;
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            head[j] = key[j] != key[j - 1];
            tail[j - 1] = head[j];
        }
        const int first_key = __shfl_sync(FULL, key[0], 0);
        const int last_key = __shfl_sync(FULL, key[3], 31);
        // segments cut by the tile boundary merge through atomics
        const bool cut_first = first_key >= 0 && first_key == __shfl_sync(FULL, ext, 0);
        const bool cut_last = last_key >= 0 && last_key == __shfl_sync(FULL, ext, 31);

        // ---- second Linear on the tensor cores: two m16 tiles (edges {0,1} and {2,3} of every lane) ----
        float z[4][CPL];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                uint32_t ah[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                 // a0: row g, a1: row g + 8, a2 / a3: the slots 2t+8, 2t+9
                    const float v0 = x[2 * m + (i & 1)][4 * kt + 2 * (i >> 1)];
                    const float v1 = x[2 * m + (i & 1)][4 * kt + 2 * (i >> 1) + 1];
                    // hi = value with the 13 low mantissa bits cleared (11 significant bits: exact in fp16 inside the
                    // range the scale guarantees), lo = value - hi (exact in fp32)
                    const float h0 = tc::tf32_hi(v0), h1 = tc::tf32_hi(v1);
                    ah[i] = tc::pack_h2(h0, h1);
                    al[i] = tc::pack_h2(v0 - h0, v1 - h1);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    mma_f16_16x8x16(acc[nt], al[0], al[1], al[2], al[3], bh[kt][nt][0], bh[kt][nt][1]);
                    mma_f16_16x8x16(acc[nt], ah[0], ah[1], ah[2], ah[3], bl[kt][nt][0], bl[kt][nt][1]);
                    mma_f16_16x8x16(acc[nt], ah[0], ah[1], ah[2], ah[3], bh[kt][nt][0], bh[kt][nt][1]);
                }
            }
            // bias -> ReLU -> BatchNorm affine (before any max: the scale may be negative)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 b = s_bias[4 * nt + t], sc = s_scale[4 * nt + t], sh = s_shift[4 * nt + t];
                z[2 * m][2 * nt] = fmaf(fmaxf(fmaf(acc[nt][0], inv, b.x), 0.f), sc.x, sh.x);
                z[2 * m][2 * nt + 1] = fmaf(fmaxf(fmaf(acc[nt][1], inv, b.y), 0.f), sc.y, sh.y);
                z[2 * m + 1][2 * nt] = fmaf(fmaxf(fmaf(acc[nt][2], inv, b.x), 0.f), sc.x, sh.x);
                z[2 * m + 1][2 * nt + 1] = fmaf(fmaxf(fmaf(acc[nt][3], inv, b.y), 0.f), sc.y, sh.y);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (key[j] >= 0) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) am = fmaxf(am, fabsf(z[j][c]));
            }

        // ---- segmented inclusive max scan over the 32 edges ----
#pragma unroll
        for (int j = 1; j < 4; ++j)
            if (!head[j]) {
#pragma unroll
                for (int c = 0; c < CPL; ++c) z[j][c] = fmaxf(z[j][c], z[j - 1][c]);
            }
        float v[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) v[c] = z[3][c];
        bool F = head[0] | head[1] | head[2] | head[3];       // a segment starts inside this lane's edges
#pragma unroll
        for (int dlt = 1; dlt < 8; dlt <<= 1) {
            const bool Fp = __shfl_up_sync(FULL, F, 4 * dlt);
            const bool take = (g >= dlt) && !F;
#pragma unroll
            for (int c = 0; c < CPL; ++c) {
                const float vp = __shfl_up_sync(FULL, v[c], 4 * dlt);
                if (take) v[c] = fmaxf(v[c], vp);
            }
            if (g >= dlt) F = F | Fp;
        }
        {
            bool open = g > 0;                                    // edges before the lane's first head continue a segment
            float cin[CPL];
#pragma unroll
            for (int c = 0; c < CPL; ++c) cin[c] = __shfl_up_sync(FULL, v[c], 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                open = open && !head[j];
                if (open) {
#pragma unroll
                    for (int c = 0; c < CPL; ++c) z[j][c] = fmaxf(z[j][c], cin[c]);
                }
            }
        }
        // ---- segment tails hold the maxima ----
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!tail[j] || key[j] < 0) continue;
            const bool atomic = (cut_first && key[j] == first_key) || (cut_last && key[j] == last_key);
            for (int r = 0; r < d.out_repeat; ++r) {
                float *dst = d.out + ((size_t)(f + r) * N + key[j]) * (size_t)d.ldo + d.out_off + 2 * t;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (atomic) {
                        atomic_max_f32(dst + 8 * nt, z[j][2 * nt]);
                        atomic_max_f32(dst + 8 * nt + 1, z[j][2 * nt + 1]);
                    } else if (vec2) {
                        *reinterpret_cast<float2 *>(dst + 8 * nt) = make_float2(z[j][2 * nt], z[j][2 * nt + 1]);
                    } else {
                        dst[8 * nt] = z[j][2 * nt];
                        dst[8 * nt + 1] = z[j][2 * nt + 1];
                    }
                }
            }
        }
    }
    amax_commit(d.out_amax, am);
}

static bool narrow_tma_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MORIG_NARROW_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// tensor map of a PQ buffer for tile::gather4: dims {ldpq, rows} fp32, row stride ldpq * 4 bytes, box {H, 1}
static int make_pq_map(CUtensorMap *map, const morig_edge_desc &d, int H) {
    const cuuint64_t dims[2] = {(cuuint64_t)d.ldpq, (cuuint64_t)d.N * (cuuint64_t)(d.out_repeat == 1 ? d.n_frames : 1)};
    const cuuint64_t strides[1] = {(cuuint64_t)d.ldpq * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)H, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    // the driver entry point is resolved through the runtime, so the library carries no link-time dependency on
    // libcuda.so (it must load on machines without a driver: the build check runs on a CPU-only box)
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void *fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return MORIG_E_BADARG;
        }
        encode = reinterpret_cast<encode_fn>(fp);
    }
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d.PQ), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for PQ [%llu x %d]", (int)r, (unsigned long long)dims[1], d.ldpq);
        return MORIG_E_BADARG;
    }
    return 0;
}

template <int H, bool TMA>
static int launch_edge_mma_impl(const morig_edge_desc *descs, int count, cudaStream_t stream) {
    constexpr size_t smem = TMA ? (size_t)(EDGE_MMA_THREADS / 32) * (2 * 32 * H * 4 + 16) : 0;
    static thread_local int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        if (smem > 48 * 1024) MORIG_CUDA(cudaFuncSetAttribute(edge_mma_kernel<H, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MORIG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, edge_mma_kernel<H, TMA>, EDGE_MMA_THREADS, smem));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
    }
    EdgeBatch b;
    EdgeMaps maps;
    for (int i = 0; i < EDGE_BATCH_MAX; ++i) {
        b.d[i] = descs[i < count ? i : 0];
        if (TMA) { if (int rc = make_pq_map(&maps.m[i], b.d[i], H)) return rc; }
    }
    const morig_edge_desc &d = descs[0];
    const long long tiles = (long long)ceil_div(d.E_max, 32) * d.n_frames;          // upper bound (E' <= E_max)
    const long long want = ceil_div64(tiles, EDGE_MMA_THREADS / 32);
    long long cap = (long long)sm_count() * blocks_per_sm / count;                  // persistent: warps stride over tiles
    if (cap < 1) cap = 1;
    MORIG_CUDA(launch_pdl(edge_mma_kernel<H, TMA>, dim3((unsigned)(want < cap ? want : cap), (unsigned)count), dim3(EDGE_MMA_THREADS), smem,
                          stream, b, maps));
    return 0;
}

template <int H>
static int launch_edge_mma(const morig_edge_desc *descs, int count, cudaStream_t stream) {
    // TMA needs 16-byte aligned box starts (p_off / q_off multiples of 4 floats: checked by check_narrow) and row strides
    if (narrow_tma_enabled() && descs[0].ldpq % 4 == 0) return launch_edge_mma_impl<H, true>(descs, count, stream);
    return launch_edge_mma_impl<H, false>(descs, count, stream);
}

static int check_narrow(const morig_edge_desc *d) {
    MORIG_CHECK_ARG(d && d->PQ && d->rowptr && d->col && d->tgt && d->W1 && d->b1 && d->scale && d->shift && d->out,
                    "edgeconv_fwd: null operand");
    MORIG_CHECK_ARG(d->N > 0 && d->E_max >= d->N && d->n_frames >= 1 && d->out_repeat >= 1, "edgeconv_fwd: bad sizes");
    MORIG_CHECK_ARG(d->out_repeat == 1 || d->n_frames == 1, "edgeconv_fwd: out_repeat needs n_frames == 1");
    MORIG_CHECK_ARG(d->ldpq % 4 == 0 && d->p_off % 4 == 0 && d->q_off % 4 == 0 && aligned16(d->PQ),
                    "edgeconv_fwd: PQ must be 16B aligned (ldpq, p_off, q_off multiples of 4)");
    MORIG_CHECK_ARG(aligned16(d->col) && aligned16(d->tgt), "edgeconv_fwd: col / tgt must be 16B aligned");
    MORIG_CHECK_ARG(d->pq_amax, "edgeconv_fwd: the narrow (H <= 32) kernel needs pq_amax");
    return 0;
}

// ---- skinny dense layer (M <= 32 rows) -------------------------------------------------------------
// The per-graph layers (e.g. the [B, 1024] x [1024, 1024] bias derived from the pooled global feature,
// models/rignet.py:64-65) have a handful of rows: the cost is streaming the weight matrix once, i.e. DRAM latency,
// so the work is spread over as many CTAs as possible while every weight read stays a full 128-byte line.
// A cluster of 8 CTAs owns 32 output columns; CTA r of the cluster takes every 8th group of 8 k-rows (split-K), its
// eight k-lanes keep the M partial sums of their column in registers, and the partials are added in a fixed order:
// first across the k-lanes through shared memory, then across the cluster through distributed shared memory by
// rank 0, which applies the epilogue.  Deterministic, no workspace.
constexpr int SKINNY_MAX_M = 32;
constexpr int SKINNY_COLS = 32;
constexpr int SKINNY_KL = 8;                        // k-lanes per CTA
constexpr int SKINNY_SPLIT = 8;                     // CTAs per cluster (split-K)
constexpr int SKINNY_THREADS = SKINNY_COLS * SKINNY_KL;

__global__ void __cluster_dims__(1, SKINNY_SPLIT, 1) __launch_bounds__(SKINNY_THREADS) dense_skinny_kernel(const GemmP p) {
    extern __shared__ float s_mem[];                // [KL][M][COLS] k-lane partials; then [M][COLS] CTA partial at the front
    const int col = threadIdx.x % SKINNY_COLS, kl = threadIdx.x / SKINNY_COLS;
    pdl_trigger();
    pdl_wait();
    const int rank = blockIdx.y;                    // cluster dims (1, 8, 1): rank in the cluster == blockIdx.y
    const int n = blockIdx.x * SKINNY_COLS + col;
    const bool n_ok = n < p.N;
    float acc[SKINNY_MAX_M];
#pragma unroll
    for (int m = 0; m < SKINNY_MAX_M; ++m) acc[m] = 0.f;
    // k = 4 * (kl + 8 * (rank + 8 * i)) .. + 3: a warp reads four full 128-byte lines of W per step
    for (int k4 = 4 * (kl + SKINNY_KL * rank); k4 < p.K; k4 += 4 * SKINNY_KL * SKINNY_SPLIT) {      // K % 4 == 0
        float w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = n_ok ? p.W[(size_t)(k4 + j) * p.ldw + n] : 0.f;
#pragma unroll
        for (int m = 0; m < SKINNY_MAX_M; ++m) {
            if (m < p.M) {
                const float4 a = *reinterpret_cast<const float4 *>(p.A + (size_t)m * p.lda + k4);   // warp-uniform address
                acc[m] = fmaf(a.x, w[0], fmaf(a.y, w[1], fmaf(a.z, w[2], fmaf(a.w, w[3], acc[m]))));
            }
        }
    }
#pragma unroll
    for (int m = 0; m < SKINNY_MAX_M; ++m)
        if (m < p.M) s_mem[(kl * p.M + m) * SKINNY_COLS + col] = acc[m];
    __syncthreads();
    float part[SKINNY_MAX_M / SKINNY_KL];           // this thread's share of the [M][COLS] CTA partial
#pragma unroll
    for (int i = 0; i < SKINNY_MAX_M / SKINNY_KL; ++i) {
        const int m = kl + SKINNY_KL * i;
        float sum = 0.f;
        if (m < p.M)
            for (int g = 0; g < SKINNY_KL; ++g) sum += s_mem[(g * p.M + m) * SKINNY_COLS + col];      // fixed order
        part[i] = sum;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SKINNY_MAX_M / SKINNY_KL; ++i) {
        const int m = kl + SKINNY_KL * i;
        if (m < p.M) s_mem[m * SKINNY_COLS + col] = part[i];
    }
    // cluster barrier (release / acquire): every CTA's partial is visible to rank 0
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    float am = 0.f;
    if (rank == 0) {
        const uint32_t base = (uint32_t)__cvta_generic_to_shared(s_mem);
#pragma unroll
        for (int i = 0; i < SKINNY_MAX_M / SKINNY_KL; ++i) {
            const int m = kl + SKINNY_KL * i;
            if (m >= p.M) continue;
            float sum = 0.f;
            for (uint32_t r = 0; r < SKINNY_SPLIT; ++r) {                                           // fixed order
                uint32_t addr;
                float v;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(base + 4u * (m * SKINNY_COLS + col)), "r"(r));
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
                sum += v;
            }
            if (n_ok) {
                float x = sum + (p.bias ? p.bias[n] : 0.f);
                if (p.relu) x = fmaxf(x, 0.f);
                const float z = fmaf(x, p.scale ? p.scale[n] : 1.f, p.shift ? p.shift[n] : 0.f);
                p.C[(size_t)m * p.ldc + n] = z;
                am = fmaxf(am, fabsf(z));
            }
        }
    }
    // no CTA may exit while rank 0 still reads its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    amax_commit(p.amax_out, am);
}

// ---- tiny-K dense layer (K <= 8: the first Linear on 3-D positions / flows) -----------------------------------
// [M, K] x [K, N] with K of 3 or 4 is a streaming write of M x N outputs: one thread per (row, 4 consecutive
// columns), weights through L1, float4 stores.
constexpr int SMALLK_MAX = 8;

__global__ void __launch_bounds__(256) dense_smallk_kernel(const GemmP p) {
    pdl_trigger();
    pdl_wait();
    const int nq = p.N >> 2;                                    // N % 4 == 0 (checked by the launcher)
    const long long total = (long long)p.M * nq, stride = (long long)gridDim.x * blockDim.x;
    float am = 0.f;
    // grid-stride: the operand-range atomic at the end is one per warp of a bounded grid, not one per 32 outputs
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int m = (int)(idx / nq), n = (int)(idx - (long long)m * nq) * 4;
        float a[SMALLK_MAX];
#pragma unroll
        for (int k = 0; k < SMALLK_MAX; ++k) a[k] = k < p.K ? p.A[(size_t)m * p.lda + k] : 0.f;
        float4 acc = p.bias ? *reinterpret_cast<const float4 *>(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < SMALLK_MAX; ++k) {
            if (k < p.K) {
                const float4 w = *reinterpret_cast<const float4 *>(p.W + (size_t)k * p.ldw + n);
                acc.x = fmaf(a[k], w.x, acc.x); acc.y = fmaf(a[k], w.y, acc.y);
                acc.z = fmaf(a[k], w.z, acc.z); acc.w = fmaf(a[k], w.w, acc.w);
            }
        }
        if (p.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        if (p.scale) {
            const float4 sc = *reinterpret_cast<const float4 *>(p.scale + n), sh = *reinterpret_cast<const float4 *>(p.shift + n);
            acc.x = fmaf(acc.x, sc.x, sh.x); acc.y = fmaf(acc.y, sc.y, sh.y);
            acc.z = fmaf(acc.z, sc.z, sh.z); acc.w = fmaf(acc.w, sc.w, sh.w);
        }
        *reinterpret_cast<float4 *>(p.C + (size_t)m * p.ldc + n) = acc;
        am = fmaxf(am, fmaxf(fmaxf(fabsf(acc.x), fabsf(acc.y)), fmaxf(fabsf(acc.z), fabsf(acc.w))));
    }
    amax_commit(p.amax_out, am);
}

template <int BM, int BN, int AMODE, int EPI>
static int launch_gemm(const GemmP &p, dim3 grid, cudaStream_t stream, const char *name) {
    auto kern = gemm_simt_kernel<BM, BN, AMODE, EPI>;
    constexpr size_t smem = gemm_smem_bytes<BM, BN, EPI>();
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_dev = dev;
    }
    MORIG_CUDA(launch_pdl(kern, grid, dim3(GEMM_THREADS), smem, stream, p));
    (void)name;
    return 0;
}


static long long *g_trace = nullptr;      // debug timeline buffer (device), set by morig_debug_set_trace

static bool env_flag(const char *name) {
    const char *e = getenv(name);
    return e && e[0] == '1';
}

template <int KIND> static void fill_tcp(tc::TcP &tp, const GemmP &p, const void *blob, int BN, int frames) {
    tp.g = p;
    tp.Bblob = reinterpret_cast<const float *>(blob);
    tp.nK = ceil_div(p.K, tc::KindCfg<KIND>::KSTAGE);
    tp.ntn = ceil_div(p.N, BN);
    tp.ntm = ceil_div(p.M, tc::BM);
    tp.frames = frames;
    tp.trace = nullptr;
}

// cta_group::2 variant (UMMA 256 x BN on a 2-CTA cluster): halves the weight traffic and the MMA operand reads per SM
template <int KIND, int BN, int AMODE, int EPI>
static int launch_tc2(const GemmP &p, const void *blob, int frames, cudaStream_t stream, const char *name) {
    auto kern = tc::tc2_gemm_kernel<KIND, BN, AMODE, EPI>;
    constexpr int smem = tc::SMEM_BYTES;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_dev = dev;
    }
    tc::TcP tp;
    fill_tcp<KIND>(tp, p, blob, BN, frames);
    using C2 = tc::Cfg2<BN>;
    tp.resident_b = (tp.ntn == 1 && C2::res_stages(tp.nK) >= 2 && !env_flag("MORIG_NO_RESB")) ? 1 : 0;
    tp.stages = tp.resident_b ? C2::res_stages(tp.nK) : C2::STAGES;
    tp.trace = g_trace;
    const long long pairs = (long long)tp.ntn * ceil_div(tp.ntm, 2) * frames;
    const int clusters_max = sm_count() / 2;
    const unsigned grid = 2u * (unsigned)(pairs < clusters_max ? pairs : clusters_max);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(tc::THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    MORIG_CUDA(cudaLaunchKernelEx(&cfg, kern, tp));
    MORIG_LAUNCH_CHECK(name);
    return 0;
}

template <int KIND, int BN, int AMODE, int EPI>
static int launch_tc(const GemmP &p, const void *blob, int frames, cudaStream_t stream, const char *name) {
    if constexpr (BN == 256) {
        if (!env_flag("MORIG_NO_2CTA")) {
            // enough pair-tiles to give every 2-CTA cluster of the machine at least one
            const long long pairs = (long long)ceil_div(p.N, BN) * ceil_div(ceil_div(p.M, tc::BM), 2) * frames;
            if (pairs >= sm_count() / 2) return launch_tc2<KIND, BN, AMODE, EPI>(p, blob, frames, stream, name);
        }
    }
    auto kern = tc::tc_gemm_kernel<KIND, BN, AMODE, EPI>;
    constexpr int smem = tc::SMEM_BYTES;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MORIG_CUDA(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MORIG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_dev = dev;
    }
    tc::TcP tp;
    fill_tcp<KIND>(tp, p, blob, BN, frames);
    using C = tc::Cfg<BN>;
    tp.resident_b = (tp.ntn == 1 && C::res_stages(tp.nK) >= 2 && !env_flag("MORIG_NO_RESB")) ? 1 : 0;
    tp.stages = tp.resident_b ? C::res_stages(tp.nK) : C::STAGES;
    tp.trace = g_trace;
    const long long tiles = (long long)tp.ntn * tp.ntm * frames;
    const int sms = sm_count();
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);     // persistent: one CTA per SM
    MORIG_CUDA(launch_pdl(kern, dim3(grid), dim3(tc::THREADS), (size_t)smem, stream, tp));
    (void)name;
    return 0;
}

template <int BN, int AMODE, int EPI>
static int launch_tc_kind(int kind, const GemmP &p, const void *blob, int frames, cudaStream_t stream, const char *name) {
    if (kind == tc::KIND_F16) return launch_tc<tc::KIND_F16, BN, AMODE, EPI>(p, blob, frames, stream, name);
    return launch_tc<tc::KIND_TF32, BN, AMODE, EPI>(p, blob, frames, stream, name);
}

static bool tc_disabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("MORIG_NO_TC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

// ---- device-side packer of the fp16-split weight image (training path) ------------------------------------------------
__global__ void __launch_bounds__(256) pack_absmax_kernel(const float *__restrict__ src, int lds, int rows, int cols, float *amax) {
    float am = 0.f;
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        am = fmaxf(am, fabsf(src[(size_t)(i / cols) * lds + (i % cols)]));
    amax_commit(amax, am);
}

__global__ void __launch_bounds__(256) pack_tc_f16_kernel(const float *__restrict__ src, int lds, int N, int K, int transposed, int bn,
                                                          int nK, int64_t total, __half *__restrict__ blob,
                                                          const float *__restrict__ amax, float *__restrict__ w_inv_dev) {
    // scale 2^j with max|W| 2^j in [2^14, 2^15)  (packing.pack_tc_blob: j = 14 - floor(log2(amax)))
    const uint32_t bits = __float_as_uint(*amax);
    int e = (int)((bits >> 23) & 0xffu);
    int j = (e == 0 || e == 255) ? 0 : 14 - (e - 127);
    j = j < -100 ? -100 : (j > 100 ? 100 : j);
    const float scale = __uint_as_float((uint32_t)(j + 127) << 23);
    if (blockIdx.x == 0 && threadIdx.x == 0) *w_inv_dev = __uint_as_float((uint32_t)(127 - j) << 23);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        // idx enumerates the padded logical matrix [nt * bn, nK * 64] row-major
        const int k = (int)(idx % (nK * 64));
        const int n = (int)(idx / (nK * 64));
        float w = 0.f;
        if (n < N && k < K) w = transposed ? src[(size_t)k * lds + n] : src[(size_t)n * lds + k];
        const float ws = w * scale;
        const __half hi = __float2half_rn(ws);
        const __half lo = __float2half_rn(ws - __half2float(hi));
        const int nt = n / bn, r = n % bn, kc = k / 64, kk = k % 64, chunk = kk / 8, el = kk % 8;
        const size_t base = ((((size_t)nt * nK + kc) * 2) * bn + r) * 64 + (size_t)((chunk ^ (r & 7)) * 8 + el);
        blob[base] = hi;
        blob[base + (size_t)bn * 64] = lo;
    }
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API size_t morig_pack_tc_f16_bytes(int32_t N, int32_t K, int32_t bn) {
    return (size_t)ceil_div(N, bn) * ceil_div(K, 64) * 2 * bn * 128;
}

extern "C" MORIG_API int morig_pack_tc_f16(const float *src, int32_t lds, int32_t N, int32_t K, int32_t transposed, int32_t bn, void *blob,
                                           float *w_inv_dev, float *amax_scratch, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(src && blob && w_inv_dev && amax_scratch && N > 0 && K > 0 && (bn == 128 || bn == 256), "pack_tc_f16: bad argument");
    MORIG_CUDA(cudaMemsetAsync(amax_scratch, 0, sizeof(float), stream));
    const int rows = transposed ? K : N, cols = transposed ? N : K;
    const int64_t nel = (int64_t)rows * cols;
    const int64_t b1 = ceil_div64(nel, 256 * 4);
    pack_absmax_kernel<<<(unsigned)(b1 > 296 ? 296 : b1), 256, 0, stream>>>(src, lds, rows, cols, amax_scratch);
    MORIG_LAUNCH_CHECK("pack_absmax_kernel");
    const int nK = ceil_div(K, 64);
    const int64_t total = (int64_t)ceil_div(N, bn) * bn * nK * 64;
    const int64_t b2 = ceil_div64(total, 256);
    pack_tc_f16_kernel<<<(unsigned)(b2 > 148 * 8 ? 148 * 8 : b2), 256, 0, stream>>>(src, lds, N, K, transposed, bn, nK, total,
                                                                                 reinterpret_cast<__half *>(blob), amax_scratch, w_inv_dev);
    MORIG_LAUNCH_CHECK("pack_tc_f16_kernel");
    return 0;
}

// debug only (not declared in the public header): device buffer of 3 * 2048 * 2 int64 receiving the role timeline
extern "C" MORIG_API void morig_debug_set_trace(long long *buf) { g_trace = buf; }

extern "C" MORIG_API int morig_dense_fwd(const morig_dense_desc *d, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(d && d->A && d->W, "dense_fwd: null operand");
    MORIG_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "dense_fwd: M=%d N=%d K=%d", d->M, d->N, d->K);
    MORIG_CHECK_ARG(d->ldw % 4 == 0 && d->ldw >= d->N && aligned16(d->W), "dense_fwd: W must be 16B aligned with ldw%%4==0");
    MORIG_CHECK_ARG(d->C || d->pool, "dense_fwd: no output");
    MORIG_CHECK_ARG(!d->rowbias || d->batch, "dense_fwd: rowbias needs batch");
    MORIG_CHECK_ARG(!d->batch || (d->n_vtx > 0 && d->n_graphs > 0 && d->M % d->n_vtx == 0),
                    "dense_fwd: M=%d not a multiple of n_vtx=%d", d->M, d->n_vtx);
    GemmP p{};
    p.A = d->A; p.lda = d->lda;
    p.a_vec = (d->lda % 4 == 0 && d->K % 4 == 0 && aligned16(d->A)) ? 1 : 0;
    p.W = d->W; p.ldw = d->ldw;
    p.bias = d->bias; p.scale = d->scale; p.shift = d->shift;
    p.rowbias = d->rowbias; p.ldrb = d->ldrb;
    p.batch = d->batch; p.n_vtx = d->batch ? d->n_vtx : d->M; p.n_graphs = d->batch ? d->n_graphs : 1;
    p.C = d->C; p.ldc = d->ldc;
    p.c_vec = (d->C && d->ldc % 4 == 0 && aligned16(d->C)) ? 1 : 0;
    p.pool = d->pool; p.ldpool = d->ldpool;
    p.M = d->M; p.N = d->N; p.K = d->K; p.relu = d->relu;
    p.amax_in = d->a_amax; p.amax_out = d->c_amax; p.w_inv = 1.f;
    if (d->K <= SMALLK_MAX && d->N % 4 == 0 && p.c_vec && d->C && !d->pool && !d->rowbias && aligned16(d->W) &&
        (!d->bias || aligned16(d->bias)) && (!d->scale || (d->shift && aligned16(d->scale) && aligned16(d->shift)))) {
        const long long blocks = ceil_div64((long long)d->M * (d->N / 4), 256), cap = (long long)sm_count() * 8;
        MORIG_CUDA(launch_pdl(dense_smallk_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), 0, stream, p));
        return 0;
    }
    if (d->M <= SKINNY_MAX_M && d->K >= 64 && p.a_vec && d->C && !d->pool && !d->rowbias) {
        const size_t smem = (size_t)SKINNY_KL * d->M * SKINNY_COLS * sizeof(float);
        MORIG_CUDA(launch_pdl(dense_skinny_kernel, dim3(ceil_div(d->N, SKINNY_COLS), SKINNY_SPLIT), dim3(SKINNY_THREADS), smem, stream, p));
        return 0;
    }
    if (d->Wtc && p.a_vec && !tc_disabled()) {
        MORIG_CHECK_ARG(aligned16(d->Wtc), "dense_fwd: Wtc must be 16B aligned");
        MORIG_CHECK_ARG(ceil_div(d->M, 128) <= 65535, "dense_fwd: M=%d too large for one launch", d->M);
        MORIG_CHECK_ARG((uint64_t)d->M * (uint64_t)d->lda < (1ull << 32), "dense_fwd: M*lda exceeds 32-bit element offsets");
        MORIG_CHECK_ARG((uint64_t)d->M * (uint64_t)(d->C ? d->ldc : 1) < (1ull << 32), "dense_fwd: M*ldc exceeds 32-bit element offsets");
        MORIG_CHECK_ARG(d->tc_kind == tc::KIND_TF32 || d->tc_kind == tc::KIND_F16, "dense_fwd: tc_kind=%d", d->tc_kind);
        if (d->tc_kind == tc::KIND_F16) {
            MORIG_CHECK_ARG(d->a_amax && (d->tc_w_inv > 0.f || d->tc_w_inv_dev), "dense_fwd: the fp16 kind needs a_amax and tc_w_inv");
            p.w_inv = d->tc_w_inv;
            p.w_inv_dev = d->tc_w_inv_dev;
        }
        switch (d->tc_bn) {
            case 128: return launch_tc_kind<128, AMODE_PLAIN, EPI_STORE>(d->tc_kind, p, d->Wtc, 1, stream, "tc_dense<128>");
            case 256: return launch_tc_kind<256, AMODE_PLAIN, EPI_STORE>(d->tc_kind, p, d->Wtc, 1, stream, "tc_dense<256>");
            default:  MORIG_CHECK_ARG(false, "dense_fwd: tc_bn=%d unsupported (128,256)", d->tc_bn);
        }
    }
    if (d->N <= 64) {
        dim3 grid(ceil_div(d->M, 128), ceil_div(d->N, 64), 1);
        return launch_gemm<128, 64, AMODE_PLAIN, EPI_STORE>(p, grid, stream, "dense<128,64>");
    }
    dim3 grid(ceil_div(d->M, 128), ceil_div(d->N, 128), 1);
    return launch_gemm<128, 128, AMODE_PLAIN, EPI_STORE>(p, grid, stream, "dense<128,128>");
}

extern "C" MORIG_API int morig_edgeconv_fwd(const morig_edge_desc *d, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(d && d->PQ && d->rowptr && d->col && d->tgt && d->W1 && d->b1 && d->scale && d->shift && d->out,
                    "edgeconv_fwd: null operand");
    MORIG_CHECK_ARG(d->N > 0 && d->E_max >= d->N && d->n_frames >= 1 && d->out_repeat >= 1, "edgeconv_fwd: bad sizes");
    MORIG_CHECK_ARG(d->out_repeat == 1 || d->n_frames == 1, "edgeconv_fwd: out_repeat needs n_frames == 1");
    const int H = d->H;
    if (H == 16 || H == 32) {
        if (int rc = check_narrow(d)) return rc;
        return H == 16 ? launch_edge_mma<16>(d, 1, stream) : launch_edge_mma<32>(d, 1, stream);
    }
    MORIG_CHECK_ARG(H == 64 || H == 128 || H == 256, "edgeconv_fwd: H=%d unsupported (16,32,64,128,256)", H);
    MORIG_CHECK_ARG(d->out_repeat == 1, "edgeconv_fwd: out_repeat only for H<=32");
    MORIG_CHECK_ARG(d->ldpq % 4 == 0 && d->p_off % 4 == 0 && d->q_off % 4 == 0 && aligned16(d->PQ),
                    "edgeconv_fwd: PQ must be 16B aligned (ldpq, p_off, q_off multiples of 4)");
    MORIG_CHECK_ARG(d->ldw % 4 == 0 && d->ldw >= H && aligned16(d->W1), "edgeconv_fwd: W1 alignment");
    GemmP p{};
    p.P = d->PQ + d->p_off; p.Q = d->PQ + d->q_off; p.ldpq = d->ldpq;
    p.rowptr = d->rowptr; p.col = d->col; p.tgt = d->tgt; p.n_vtx_frame = d->N;
    p.W = d->W1; p.ldw = d->ldw;
    p.bias = d->b1; p.scale = d->scale; p.shift = d->shift;
    p.C = d->out + d->out_off; p.ldc = d->ldo;
    p.M = d->E_max; p.N = H; p.K = H; p.relu = 1;
    p.amax_in = d->pq_amax; p.amax_out = d->out_amax; p.w_inv = 1.f;
    if (d->W1tc && !tc_disabled()) {
        MORIG_CHECK_ARG(aligned16(d->W1tc), "edgeconv_fwd: W1tc must be 16B aligned");
        MORIG_CHECK_ARG(ceil_div(d->E_max, 128) <= 65535, "edgeconv_fwd: E=%d too large for one launch", d->E_max);
        MORIG_CHECK_ARG((uint64_t)d->N * (uint64_t)d->n_frames * (uint64_t)d->ldpq < (1ull << 32),
                        "edgeconv_fwd: PQ exceeds 32-bit element offsets");
        MORIG_CHECK_ARG((uint64_t)d->N * (uint64_t)d->n_frames * (uint64_t)d->ldo < (1ull << 32),
                        "edgeconv_fwd: out exceeds 32-bit element offsets");
        MORIG_CHECK_ARG(d->tc_kind == tc::KIND_TF32 || d->tc_kind == tc::KIND_F16, "edgeconv_fwd: tc_kind=%d", d->tc_kind);
        if (d->tc_kind == tc::KIND_F16) {
            MORIG_CHECK_ARG(d->pq_amax && d->tc_w_inv > 0.f, "edgeconv_fwd: the fp16 kind needs pq_amax and tc_w_inv");
            p.w_inv = d->tc_w_inv;
        }
        const int kind = d->tc_kind;
        if (H <= 128) return launch_tc_kind<128, AMODE_GATHER, EPI_SEGMAX>(kind, p, d->W1tc, d->n_frames, stream, "tc_edge<128>");
        return launch_tc_kind<256, AMODE_GATHER, EPI_SEGMAX>(kind, p, d->W1tc, d->n_frames, stream, "tc_edge<256>");
    }
    if (H == 64) {
        dim3 grid(ceil_div(d->E_max, 128), 1, d->n_frames);
        return launch_gemm<128, 64, AMODE_GATHER, EPI_SEGMAX>(p, grid, stream, "edge<128,64>");
    }
    dim3 grid(ceil_div(d->E_max, 128), H / 128, d->n_frames);
    return launch_gemm<128, 128, AMODE_GATHER, EPI_SEGMAX>(p, grid, stream, "edge<128,128>");
}

extern "C" MORIG_API int morig_edgeconv_fwd_batch(const morig_edge_desc *d, int32_t count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(d && count >= 1 && count <= EDGE_BATCH_MAX, "edgeconv_fwd_batch: count=%d unsupported (1..%d)", count,
                    EDGE_BATCH_MAX);
    const int H = d[0].H;
    MORIG_CHECK_ARG(H == 16 || H == 32, "edgeconv_fwd_batch: H=%d unsupported (16, 32)", H);
    for (int i = 0; i < count; ++i) {
        if (int rc = check_narrow(d + i)) return rc;
        MORIG_CHECK_ARG(d[i].H == H && d[i].rowptr == d[0].rowptr && d[i].col == d[0].col && d[i].tgt == d[0].tgt &&
                        d[i].N == d[0].N && d[i].E_max == d[0].E_max && d[i].n_frames == d[0].n_frames,
                        "edgeconv_fwd_batch: branch %d differs in width, graph or key-frame count", i);
    }
    return H == 16 ? launch_edge_mma<16>(d, count, stream) : launch_edge_mma<32>(d, count, stream);
}
