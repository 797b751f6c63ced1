// tcgen05 tile engine (sm_100a): persistent, warp-specialised 3xTF32 GEMM with fp32 accumulation in TMEM.
//
//   D[128 x BN] (+)= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi        (kind::tf32, cta_group::1, M=128, N=BN, K=8)
//
// fp32 inputs are split x = hi + lo with hi = rna_tf32(x); dropping lo*lo leaves fp32-class results (the 1e-4
// absolute tolerance of the path rules out plain TF32/BF16).
//
// One CTA per SM walks tiles t = blockIdx.x, +gridDim.x, ... ; three roles run decoupled through mbarriers:
//   warps 0-3  A producers: 32 k-columns per stage written straight into 128B-swizzled K-major shared memory;
//              plain activation rows, or relu(P[tgt[e]] + Q[col[e]]) gathered per CSR slot (fused EdgeConv)
//   warp  8    one elected lane: cp.async.bulk of the pre-split / pre-swizzled weight image of each
//              (n-tile, k-chunk) (TMA engine, mbarrier complete_tx) and the 12 tcgen05.mma per stage;
//              tcgen05.commit releases stages and publishes accumulators
//   warps 4-7  epilogue: tcgen05.ld of one of the TWO accumulator buffers (so the next tile's MMAs overlap),
//              bias (+ per-graph bias) -> ReLU -> BatchNorm affine, then
//                 store rows / per-graph column max (warp shuffles + ordered atomics) /
//                 segmented max over the CSR target: in-register segmented scan across the 32 rows of the warp,
//                 segment tails transposed through a 4 KB per-warp staging tile for coalesced row stores;
//                 segments crossing a warp's 32 rows merge with ordered-int atomic max (exact, deterministic)
#pragma once
#include "gemm_simt.cuh"

namespace morig {
namespace tc {

constexpr int BM = 128;
constexpr int KC = 32;                       // fp32 k-columns per stage = one 128-byte swizzle row
constexpr int A_HALF_BYTES = BM * 128;       // 16 KB: hi (then lo) image of the A stage
constexpr int PRODUCER_WARPS = 8;
constexpr int EPILOGUE_WARPS = 4;
constexpr int CONTROL_WARP = PRODUCER_WARPS + EPILOGUE_WARPS;
constexpr int THREADS = 32 * (PRODUCER_WARPS + EPILOGUE_WARPS + 1);
constexpr int ROWS_PER_THREAD = BM / (PRODUCER_WARPS * 4);       // 4 rows, one 16-byte chunk each
constexpr int ROW_STEP = PRODUCER_WARPS * 4;                     // rows handled by one warp-wide instruction group

template <int BN> struct Cfg {
    static constexpr int B_HALF_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * B_HALF_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int STG_LD = 33;                                        // staging tile row stride (floats)
    static constexpr int STG_BYTES = EPILOGUE_WARPS * 32 * STG_LD * 4;       // 16.5 KB
    static constexpr int AUX_BYTES = 1024;                                   // barriers, tmem pointer, row keys
    static constexpr int SMEM_BYTES = PIPE_BYTES + STG_BYTES + AUX_BYTES + 1024;   // + slack for 1024B alignment
    static constexpr int TMEM_COLS = 2 * BN;                                 // two accumulator buffers
    static constexpr int A_STAGE_BYTES = 2 * A_HALF_BYTES;
    static constexpr int B_CHUNK_BYTES = 2 * B_HALF_BYTES;
    static constexpr int PIPE_BUDGET = PIPE_BYTES;                           // 192 KB for every BN
    // resident-B mode (all k-chunks of the weight image stay in shared memory for the whole kernel):
    // possible when the CTA only ever sees one n-tile and the image leaves room for >= 2 A stages
    static constexpr int res_stages(int nK) {
        const int left = PIPE_BUDGET - nK * B_CHUNK_BYTES;
        const int s = left / A_STAGE_BYTES;
        return s > 4 ? 4 : s;
    }
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// A pipeline bug must surface as a launch failure, never as a hung GPU: trap after ~2 s of waiting.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B (8 rows x 128 B atoms, 1024 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=BN
template <int BN> __device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

struct TcP {
    GemmP g;                 // shared operand/epilogue description (W/ldw unused here)
    const float *Bblob;      // [n_tiles][nK][hi|lo][BN*32] pre-swizzled weight images
    int nK;                  // k-chunks of 32
    int ntn, ntm, frames;    // tile grid (ntm is an upper bound in gather mode)
    int stages;              // A (and, when streaming, B) ring depth
    int resident_b;          // 1: the whole weight image is loaded once and kept in shared memory
};

struct TileCoord { int n_tile, m0, frame; };

template <int BN, int AMODE, int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const TcP tp) {
    using C = Cfg<BN>;
    const GemmP &p = tp.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // swizzle atoms need 1024 B alignment
    uint8_t *smem = smem_raw + (base - raw_addr);
    const int nK = tp.nK;
    const int S = tp.stages;
    const bool resb = tp.resident_b != 0;
    // stage s: A image at a_off(s); B image of the stage (streaming) or of k-chunk kc (resident) at b_off(.)
    const uint32_t a_stride = resb ? (uint32_t)C::A_STAGE_BYTES : (uint32_t)C::STAGE_BYTES;
    const uint32_t b_region = resb ? (uint32_t)(S * C::A_STAGE_BYTES) : (uint32_t)C::A_STAGE_BYTES;
    const uint32_t b_stride = resb ? (uint32_t)C::B_CHUNK_BYTES : (uint32_t)C::STAGE_BYTES;
    float *stg_all = reinterpret_cast<float *>(smem + C::PIPE_BYTES);
    uint8_t *aux = smem + C::PIPE_BYTES + C::STG_BYTES;
    const uint32_t aux_addr = base + C::PIPE_BYTES + C::STG_BYTES;
    // aux: a_full[S] @0, b_full[S] @64, mma_done[S] @128, acc_full[2] @192, acc_empty[2] @208, tmem ptr @224
    auto bar_a = [&](int s) { return aux_addr + 8u * s; };
    auto bar_b = [&](int s) { return aux_addr + 64u + 8u * s; };
    auto bar_m = [&](int s) { return aux_addr + 128u + 8u * s; };
    auto bar_accf = [&](int b) { return aux_addr + 192u + 8u * b; };
    auto bar_acce = [&](int b) { return aux_addr + 208u + 8u * b; };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aux + 224);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int M = p.M, ntm = tp.ntm;
    if (AMODE == AMODE_GATHER) {
        M = p.rowptr[p.n_vtx_frame];                 // E' lives on the device only
        ntm = (M + BM - 1) / BM;
    }
    const int total_tiles = tp.ntn * ntm * tp.frames;
    auto decode = [&](int t) {
        TileCoord c;
        c.n_tile = t % tp.ntn;
        const int r = t / tp.ntn;
        c.m0 = (r % ntm) * BM;
        c.frame = r / ntm;
        return c;
    };

    if (warp == CONTROL_WARP) {
        if (lane == 0) {
            for (int s = 0; s < 4; ++s) {
                mbar_init(bar_a(s), PRODUCER_WARPS);
                mbar_init(bar_b(s), 1);
                mbar_init(bar_m(s), 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar_accf(b), 1);
                mbar_init(bar_acce(b), EPILOGUE_WARPS);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(aux_addr + 224u, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == CONTROL_WARP) {
        // ================= control warp: B bulk copies + MMA issue (one elected lane) =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BN>();
            const uint32_t b_bytes = 2u * C::B_HALF_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob);
            const int my_tiles = (total_tiles > (int)blockIdx.x) ? (total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            // All ring positions are tracked incrementally (no div/mod on this latency-critical thread).
            // fetch cursor: next (tile, k-chunk) whose weight image has to be requested, and its stage
            int f_li = 0, f_kc = 0, f_s = 0, f_ntile = (my_tiles > 0) ? decode((int)blockIdx.x).n_tile : 0;
            auto fetch_next = [&]() {                              // streaming mode only
                const uint32_t dst = base + b_region + f_s * b_stride;
                mbar_arrive_expect_tx(bar_b(f_s), b_bytes);
                bulk_g2s(dst, gB + ((size_t)f_ntile * nK + f_kc) * b_bytes, b_bytes, bar_b(f_s));
                if (++f_s == S) f_s = 0;
                if (++f_kc == nK) {
                    f_kc = 0;
                    ++f_li;
                    if (f_li < my_tiles) f_ntile = decode((int)blockIdx.x + f_li * (int)gridDim.x).n_tile;
                }
            };
            if (resb) {
                // one n-tile for the whole kernel: fetch every k-chunk of the image once
                if (my_tiles > 0) {
                    mbar_arrive_expect_tx(bar_b(0), b_bytes * (uint32_t)nK);
                    for (int kc = 0; kc < nK; ++kc)
                        bulk_g2s(base + b_region + kc * b_stride, gB + ((size_t)f_ntile * nK + kc) * b_bytes, b_bytes,
                                 bar_b(0));
                    mbar_wait(bar_b(0), 0);
                }
            } else {
                for (int i = 0; i < S - 1 && f_li < my_tiles; ++i) fetch_next();
            }
            int s = 0, prev_s = 0;
            uint32_t ph = 0, prev_ph = 0;
            bool first = true;
            for (int li = 0; li < my_tiles; ++li) {
                const int buf = li & 1;
                mbar_wait(bar_acce(buf), ((li >> 1) & 1) ^ 1);     // accumulator buffer drained by the epilogue
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kc = 0; kc < nK; ++kc) {
                    mbar_wait(bar_a(s), ph);
                    if (!resb) mbar_wait(bar_b(s), ph);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * a_stride, a_lo = a_hi + A_HALF_BYTES;
                    const uint32_t b_hi = base + b_region + (resb ? kc : s) * b_stride, b_lo = b_hi + C::B_HALF_BYTES;
#pragma unroll
                    for (int k = 0; k < KC / 8; ++k) {
                        const uint32_t ko = k * 32;              // 8 tf32 = 32 bytes along the swizzled row
                        const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
                        const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
                        umma_tf32(tmem_d, dal, dbh, idesc, (kc | k) != 0);
                        umma_tf32(tmem_d, dah, dbl, idesc, 1);
                        umma_tf32(tmem_d, dah, dbh, idesc, 1);
                    }
                    umma_commit(bar_m(s));                       // frees stage s when these MMAs retire
                    if (!resb && f_li < my_tiles) {
                        // the stage to refill was last read by the PREVIOUS chunk's MMAs
                        if (!first) mbar_wait(bar_m(prev_s), prev_ph);
                        fetch_next();
                    }
                    first = false;
                    prev_s = s; prev_ph = ph;
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(bar_accf(buf));                      // accumulator complete -> epilogue
            }
        }
        __syncwarp();
    } else if (warp < PRODUCER_WARPS) {
        // ================= producer warps: A stage images =================
        // Thread -> 16-byte chunk c of rows row0 + ROW_STEP*ps.  One warp instruction covers 4 consecutive CSR slots,
        // which mostly share P[tgt] (one coalesced line).  Raw operands of the NEXT chunk are in flight in registers
        // while the current chunk is combined, split and stored: the relu(P+Q) combine is deferred to the store
        // step so that issuing the loads never blocks.
        const int c = tid & 7;
        const int row0 = tid >> 3;
        int s = 0;                                   // ring position, tracked incrementally
        uint32_t wait_ph = 1;                        // parity of "stage s is free" (passes on a fresh barrier)
        float4 pa[ROWS_PER_THREAD], qa[ROWS_PER_THREAD], pb[ROWS_PER_THREAD], qb[ROWS_PER_THREAD];
        const float *src0[ROWS_PER_THREAD];
        const float *src1[ROWS_PER_THREAD];
        int ni[ROWS_PER_THREAD], nj[ROWS_PER_THREAD];    // gather indices of the NEXT tile, fetched a tile ahead
        uint32_t okmask = 0;

        auto fetch_indices = [&](const TileCoord &t) {
#pragma unroll
            for (int ps = 0; ps < ROWS_PER_THREAD; ++ps) {
                const int r = t.m0 + row0 + ROW_STEP * ps;
                ni[ps] = 0; nj[ps] = 0;
                if (AMODE == AMODE_GATHER && r < M) { ni[ps] = p.tgt[r]; nj[ps] = p.col[r]; }
            }
        };
        auto setup_rows = [&](const TileCoord &t) {      // consumes ni/nj of this tile
            okmask = 0;
#pragma unroll
            for (int ps = 0; ps < ROWS_PER_THREAD; ++ps) {
                const int r = t.m0 + row0 + ROW_STEP * ps;
                const bool ok = r < M;
                okmask |= (ok ? 1u : 0u) << ps;
                if (AMODE == AMODE_GATHER) {
                    const size_t fb = (size_t)t.frame * p.n_vtx_frame;
                    src0[ps] = p.P + (fb + ni[ps]) * (size_t)p.ldpq + 4 * c;
                    src1[ps] = p.Q + (fb + nj[ps]) * (size_t)p.ldpq + 4 * c;
                } else {
                    src0[ps] = p.A + (size_t)(ok ? r : 0) * p.lda + 4 * c;
                    src1[ps] = nullptr;
                }
            }
        };
        auto load_raw = [&](int kc, float4 (&pd)[ROWS_PER_THREAD], float4 (&qd)[ROWS_PER_THREAD]) {
            const int k = kc * KC + 4 * c;
#pragma unroll
            for (int ps = 0; ps < ROWS_PER_THREAD; ++ps) {
                pd[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
                qd[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (((okmask >> ps) & 1u) && k < p.K) {
                    pd[ps] = *reinterpret_cast<const float4 *>(src0[ps] + kc * KC);
                    if (AMODE == AMODE_GATHER) qd[ps] = *reinterpret_cast<const float4 *>(src1[ps] + kc * KC);
                }
            }
        };
        auto store_stage = [&](const float4 (&pd)[ROWS_PER_THREAD], const float4 (&qd)[ROWS_PER_THREAD]) {
            mbar_wait(bar_m(s), wait_ph);            // MMAs that read this stage one ring turn ago have retired
            uint8_t *a_hi = smem + s * a_stride;
            uint8_t *a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
            for (int ps = 0; ps < ROWS_PER_THREAD; ++ps) {
                const int row = row0 + ROW_STEP * ps;
                const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
                float4 v = pd[ps];
                if (AMODE == AMODE_GATHER) {
                    v.x = fmaxf(v.x + qd[ps].x, 0.f); v.y = fmaxf(v.y + qd[ps].y, 0.f);
                    v.z = fmaxf(v.z + qd[ps].z, 0.f); v.w = fmaxf(v.w + qd[ps].w, 0.f);
                }
                float4 h, l;
                h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                *reinterpret_cast<float4 *>(a_hi + off) = h;
                *reinterpret_cast<float4 *>(a_lo + off) = l;
            }
            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a(s));
            if (++s == S) { s = 0; wait_ph ^= 1; }
        };

        int t = blockIdx.x, kc = 0;
        bool in_a = true;                            // which register set holds the current chunk
        if (t < total_tiles) {
            const TileCoord t0 = decode(t);
            fetch_indices(t0);
            setup_rows(t0);
            load_raw(0, pa, qa);
            if (t + (int)gridDim.x < total_tiles) fetch_indices(decode(t + gridDim.x));
        }
        while (t < total_tiles) {
            // coordinates of the next chunk of this CTA's stream
            int kc_n = kc + 1, t_n = t;
            if (kc_n == nK) { kc_n = 0; t_n = t + gridDim.x; }
            const bool has_next = t_n < total_tiles;
            if (has_next && kc_n == 0) {
                setup_rows(decode(t_n));             // indices were fetched one tile ago
                if (t_n + (int)gridDim.x < total_tiles) fetch_indices(decode(t_n + gridDim.x));
            }
            if (in_a) {
                if (has_next) load_raw(kc_n, pb, qb);
                store_stage(pa, qa);
            } else {
                if (has_next) load_raw(kc_n, pa, qa);
                store_stage(pb, qb);
            }
            in_a = !in_a;
            kc = kc_n;
            t = t_n;
        }
    } else {
        // ================= epilogue warps =================
        // tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns).  Each 32x32 block is transposed
        // through a 4 KB per-warp staging tile so that lanes become COLUMNS: the per-column epilogue constants then
        // live in registers, the 32 rows of the column are processed from registers with compile-time indices, a
        // running max is restarted at segment heads and flushed at segment tails (CSR target for the EdgeConv,
        // graph id for pooling), and every global access is a coalesced 128-byte row segment.
        const int q = warp & 3;                      // TMEM lane quarter of this warp (warps 8..11 -> 0..3)
        float *stg = stg_all + q * 32 * C::STG_LD;
        int32_t *s_key = reinterpret_cast<int32_t *>(aux + 256) + q * 32;
        int li = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
            const TileCoord tcd = decode(t);
            const int buf = li & 1;
            const int n0 = tcd.n_tile * BN;
            const int rbase = tcd.m0 + q * 32;       // first row (CSR slot) of this warp
            const int r = rbase + lane;
            const bool row_ok = r < M;
            const int valid_rows = min(32, max(0, M - rbase));
            // segment key of this lane's row (rows are sorted by it); independent of the accumulator
            int key = -1;
            if (row_ok) {
                if (EPI == EPI_SEGMAX) key = p.tgt[r];
                else key = p.batch ? (r / p.n_vtx) * p.n_graphs + p.batch[r % p.n_vtx] : 0;
            }
            const int key_up = __shfl_up_sync(0xffffffffu, key, 1);
            const int key_dn = __shfl_down_sync(0xffffffffu, key, 1);
            const bool is_head = (lane == 0) || (key != key_up);
            const bool is_tail = (key >= 0) && ((lane == 31) || (key != key_dn));
            const uint32_t head_mask = __ballot_sync(0xffffffffu, is_head);
            const uint32_t tail_mask = __ballot_sync(0xffffffffu, is_tail);
            bool complete = false;                   // segment entirely inside this warp's 32 rows -> plain store
            if (EPI == EPI_SEGMAX && is_tail) complete = p.rowptr[key] >= rbase && p.rowptr[key + 1] <= rbase + 32;
            const uint32_t complete_mask = __ballot_sync(0xffffffffu, complete);
            const size_t frame_base = (EPI == EPI_SEGMAX) ? (size_t)tcd.frame * p.n_vtx_frame : 0;
            const bool one_group = (tail_mask & (tail_mask - 1)) == 0;      // at most one segment in these rows
            const int key0 = __shfl_sync(0xffffffffu, key, 0);
            __syncwarp();
            s_key[lane] = key;
            __syncwarp();

            mbar_wait(bar_accf(buf), (uint32_t)((li >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int cb = 0; cb < BN / 32; ++cb) {
                const int col0 = cb * 32;
                float xr[32];
                {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + col0), v);
                    __syncwarp();                    // previous block's reads of the staging tile are done
#pragma unroll
                    for (int j = 0; j < 32; ++j) stg[lane * C::STG_LD + j] = v[j];
                    __syncwarp();
                }
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) xr[rr] = stg[rr * C::STG_LD + lane];   // lane = column from here on
                const int nl = n0 + col0 + lane;
                const bool nl_ok = nl < p.N;
                float bias_l = (nl_ok && p.bias) ? p.bias[nl] : 0.f;
                const float scale_l = (nl_ok && p.scale) ? p.scale[nl] : 1.f;
                const float shift_l = (nl_ok && p.shift) ? p.shift[nl] : 0.f;
                const bool relu = (EPI == EPI_SEGMAX) || p.relu;
                const bool rowbias_slow = (EPI == EPI_STORE) && p.rowbias && !one_group;
                if (EPI == EPI_STORE && p.rowbias && one_group && nl_ok && key0 >= 0)
                    bias_l += p.rowbias[(size_t)key0 * p.ldrb + nl];
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    float x = xr[rr] + bias_l;
                    if (rowbias_slow) {              // rows of several graphs in one warp: rare
                        const int g = s_key[rr];
                        if (g >= 0 && nl_ok) x += p.rowbias[(size_t)g * p.ldrb + nl];
                    }
                    if (relu) x = fmaxf(x, 0.f);
                    xr[rr] = fmaf(x, scale_l, shift_l);              // BatchNorm affine BEFORE any max (scale may be < 0)
                }
                if (EPI == EPI_STORE && p.C && nl_ok) {
                    float *dst = p.C + (size_t)rbase * p.ldc + nl;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        if (rr < valid_rows) *dst = xr[rr];
                        dst += p.ldc;
                    }
                }
                if (EPI == EPI_SEGMAX || p.pool) {
                    float m = neg_inf();
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) {
                        m = ((head_mask >> rr) & 1u) ? xr[rr] : fmaxf(m, xr[rr]);
                        if ((tail_mask >> rr) & 1u) {                   // warp-uniform
                            const int k_rr = s_key[rr];
                            if (nl_ok) {
                                if (EPI == EPI_SEGMAX) {
                                    float *dst = p.C + (frame_base + k_rr) * (size_t)p.ldc + nl;
                                    if ((complete_mask >> rr) & 1u) *dst = m;
                                    else atomic_max_f32(dst, m);
                                } else {
                                    atomic_max_f32(p.pool + (size_t)k_rr * p.ldpool + nl, m);
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acce(buf));           // buffer may be overwritten by tile li + 2
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CONTROL_WARP) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace tc
}  // namespace morig
