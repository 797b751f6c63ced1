// tcgen05 tile engine (sm_100a): persistent, warp-specialised split-precision GEMM with fp32 accumulation in TMEM.
//
//   D[128 x BN] (+)= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
//
// fp32 inputs are split x = hi + lo; dropping lo*lo leaves fp32-class results (the 1e-4 absolute tolerance of the
// path rules out plain TF32/BF16/FP16).  Two operand kinds share all of the code below:
//   KIND_TF32  hi = x with 13 mantissa bits cleared, lo = x - hi          tcgen05.mma kind::tf32, K =  8 / instruction
//   KIND_F16   hi = fp16(x*s) (11 significant bits), lo = fp16(x*s - hi)  tcgen05.mma kind::f16,  K = 16 / instruction
//              -> the same 22 significant bits per operand at twice the MMA rate and half the shared-memory bytes.
//              fp16 has a 5-bit exponent, so operands are scaled by powers of two: weights at pack time, activations
//              by s = 2^(15 - ceil(log2(amax))) where amax is the running max |value| of the source buffer that every
//              producing kernel maintains (GemmP::amax_in / amax_out); the epilogue multiplies by the exact inverse.
//              Elements down to 2^-17 of the buffer's max keep all 22 bits (fp16 subnormals resolve 2^-24 * 2^-15).
//
// The MMA is issued "transposed": the WEIGHT image is the tensor core's A operand (M = 128 output channels per CTA)
// and the activation stage its B operand (N = 128 rows per CTA), so the accumulator in TMEM is D^T -- TMEM lane =
// output channel, TMEM column = row (edge / vertex) of the tile.  tcgen05.ld (32 lanes x 32 columns) then hands every
// epilogue lane ONE CHANNEL and 32 consecutive rows in registers: per-channel constants are per-lane scalars, the
// segmented max over rows runs over registers with compile-time indices, and every global access is a coalesced
// 128-byte row segment -- no shared-memory transposition (an earlier version staged every 32x32 block through shared
// memory; timeline traces showed the epilogue warps, not the MMAs, bounding the fused EdgeConv kernels).
//
// One CTA per SM walks tiles; the roles run decoupled through mbarriers:
//   warps 0-7   A producers: 32 k-columns per stage unit written straight into 128B-swizzled K-major shared memory;
//               plain activation rows (producer_role: four 16-register load buffers = two stages of loads in flight), or
//               relu(P[tgt[e]] + Q[col[e]]) gathered per CSR slot (producer_gather_role, the fused EdgeConv: tile body
//               unrolled over the K / 32 units, ReLU folded into the fp16 conversion)
//   warp  16    one elected lane issues the 12 tcgen05.mma per 128 channels and stage; tcgen05.commit releases stages
//               and publishes accumulators
//   warp  17    streaming mode: cp.async.bulk of the pre-split / pre-swizzled weight image of each (n-tile, k-chunk)
//               (TMA engine, mbarrier complete_tx), requested as soon as the ring slot's previous MMAs retired --
//               independent of the issuer, which a late chunk would otherwise keep from requesting the next one.
//               (Resident mode: the issuer loads the whole image once.)
//   warps 8-15  epilogue, two per TMEM lane quarter (= 32 channels), working on one of the 2-4 accumulator buffers while
//               the next tiles' MMAs run:
//               dense layers (epilogue_role): alternating 32-row blocks; bias (+ per-graph bias) -> ReLU -> BatchNorm
//                 affine and coalesced row stores / per-graph column max
//               fused EdgeConv (epilogue_segmax_role): a contiguous half of the tile's rows per warp; running max over
//                 the CSR segments in registers, written back to TMEM, segment tails read back by runtime column address;
//                 bias -> ReLU -> BatchNorm once per segment; segments that leave the warp's range merge with the
//                 ordered-int atomic max: exact and order independent, hence deterministic
//
// Two kernels share the role code:
//   tc_gemm_kernel<BN,...>   cta_group::1, BN / 128 UMMAs of 128 x 128 per k-step; weight image streamed per stage or,
//                            when the CTA sees a single n-tile and the image fits, kept resident in shared memory
//   tc2_gemm_kernel<...>     cta_group::2 on a 2-CTA cluster, UMMA 256 x 256: CTA r holds channels [128 r, 128 r + 128)
//                            of the n-tile (its half of every weight chunk) and produces rows [128 r, 128 r + 128) of
//                            the 256-row pair tile (the pair shares the activation operand through the tensor core's
//                            cross-CTA path), which halves the L2 -> SM weight traffic of the streamed layers
#pragma once
#include <cuda_fp16.h>
#include "gemm_simt.cuh"
#include "tile_iter.cuh"

namespace morig {
namespace tc {

enum { KIND_TF32 = 0, KIND_F16 = 1 };

constexpr int KC = 32;                       // fp32 k-columns per producer ring unit
// one stage = 128-byte swizzle rows: 32 tf32 (one ring unit) or 64 fp16 (two ring units)
template <int KIND> struct KindCfg {
    static constexpr int UPS = (KIND == KIND_F16) ? 2 : 1;      // ring units per stage
    static constexpr int KSTAGE = KC * UPS;                     // k-columns per stage
};
constexpr int A_HALF_BYTES = BM * 128;       // 16 KB: hi (then lo) image of the A stage
constexpr int A_STAGE_BYTES = 2 * A_HALF_BYTES;
constexpr int PRODUCER_WARPS = 8;
constexpr int EPILOGUE_WARPS = 8;                                // two per TMEM lane quarter, alternating 32-row blocks
constexpr int CONTROL_WARP = PRODUCER_WARPS + EPILOGUE_WARPS;
constexpr int THREADS = 32 * (PRODUCER_WARPS + EPILOGUE_WARPS + 4);    // control warp + 3 idle warps: a full warpgroup
// Register budget: 20 warps x 96 registers at launch (640 threads cap the launch allocation at 61440 of the 65536
// registers, and setmaxnreg can only redistribute what the CTA was given); the roles then rebalance with setmaxnreg
// (which works on whole warpgroups, hence the 3 idle warps next to the control warp).  Every scheduler hosts 2 producer
// warps + 2 epilogue warps + 1 control-group warp, and 2 P + 2 E + C must stay within 5 x 96 registers per lane:
//   fused EdgeConv (gather producers: two 32-register load buffers, eight row pointers, prefetched indices; segmented-max
//       epilogue with TMEM write-back: one 32-row block in registers): P = 128, E = 88, C = 48
//       (a third load buffer at P = 144 / E = 72 was measured: no gain -- the producers are bound by instruction issue and
//       dependent-issue latency, not by the ~1000-cycle gather latency)
//   dense layers (plain producers: 16 registers per buffer; store / pool epilogue with two 32-register arrays):
//       P = 112, E = 104, C = 48
// Both data roles are latency-bound straight-line code, so two warps of each per scheduler (hiding each other's stalls)
// matter more than deeper per-thread rings.
constexpr int REGS_CONTROL = 48;
template <int AMODE> struct RoleCfg {
    static constexpr int REGS_PRODUCER = (AMODE == AMODE_GATHER) ? 128 : 112;
    static constexpr int REGS_EPILOGUE = (AMODE == AMODE_GATHER) ? 88 : 104;
    static_assert(2 * REGS_PRODUCER + 2 * REGS_EPILOGUE + REGS_CONTROL <= 5 * 96, "setmaxnreg budget exceeds the launch allocation");
};
constexpr int ROWS_PER_THREAD = BM / (PRODUCER_WARPS * 4);       // 4 rows, one 16-byte fp32 chunk each
constexpr int AUX_BYTES = 1024;                                  // barriers, tmem pointer
constexpr int PIPE_BYTES = 192 * 1024;                           // operand ring (every configuration)
constexpr int KEYS_PER_WARP = 4 * 32 + 8;                        // row keys of a warp's (up to) 4 row blocks + before/after pairs
constexpr int KEYS_BYTES = EPILOGUE_WARPS * KEYS_PER_WARP * 4;
constexpr int SMEM_BYTES = PIPE_BYTES + AUX_BYTES + KEYS_BYTES + 1024;   // + slack for 1024 B alignment

template <int BN> struct Cfg {
    static constexpr int B_HALF_BYTES = BN * 128;
    static constexpr int B_CHUNK_BYTES = 2 * B_HALF_BYTES;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_CHUNK_BYTES;
    static constexpr int STAGES = PIPE_BYTES / STAGE_BYTES;                  // 2 / 3 for BN = 256 / 128
    static constexpr int NH = BN / 128;                                      // 128-channel accumulators per buffer
    static constexpr int NBUF = 512 / BN;                                    // accumulator buffers: all 512 TMEM columns
    static constexpr int TMEM_COLS = NBUF * BN;                              // (4 for BN = 128, 2 for BN = 256)
    // resident-B mode (all k-chunks of the weight image stay in shared memory for the whole kernel):
    // possible when the CTA only ever sees one n-tile and the image leaves room for >= 2 A stages
    static constexpr int res_stages(int nK) {
        const int left = PIPE_BYTES - nK * B_CHUNK_BYTES;
        const int s = left / A_STAGE_BYTES;
        return s > 4 ? 4 : s;
    }
};

// aux block layout (byte offsets): barriers are 8 bytes each
constexpr uint32_t AUX_A_FULL = 0, AUX_B_FULL = 64, AUX_MMA_DONE = 128, AUX_ACC_FULL = 192, AUX_ACC_EMPTY = 224,
                   AUX_TMEM_PTR = 256, AUX_B_PEER = 264;

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default semantics (release at
// CTA scope) on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR in front of every arrive, which drains
// the producer's prefetched loads once per stage; what the remote waiter consumes is shared memory written before a
// fence.proxy.async (or TMEM reads retired before tcgen05.fence::before_thread_sync), which that fence already orders.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits use the default acquire at CTA scope also for barriers that are signalled from the peer CTA or by a multicast
// tcgen05.commit: `.acquire.cluster` makes ptxas append CCTL.IVALL (an L1 invalidation of the whole SM, i.e. of the
// P[tgt] lines the producers keep hitting) to every successful wait.
// try_wait carries a suspend-time hint: the hardware parks the warp until the phase completes or the hint (in ns)
// expires, so a waiting role does not burn issue slots of the producers / epilogue warps sharing its scheduler
// (ncu: 9 % of all executed instructions were wait-loop iterations before the hint was added).
template <bool CLUSTER>
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
// A pipeline bug must surface as a launch failure, never as a hung GPU: trap after ~2 s of waiting.
template <bool CLUSTER = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait<CLUSTER>(bar, parity)) return;
    uint32_t spins = 0;
    const long long t0 = clock64();
    while (!mbar_try_wait<CLUSTER>(bar, parity)) {
        if ((++spins & 63u) == 0 && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int CTAS> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    if (CTAS == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CTAS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit: the mbarrier gets one arrival when all tcgen05.mma issued so far by this thread have retired.
// cta_group::2 multicasts the arrival to the barrier at the same offset in both CTAs of the pair.
template <int CTAS> __device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CTAS == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
    }
}
#define MORIG_UMMA(GROUP, KINDSTR)                                                                   \
    asm volatile(                                                                                   \
        "{\n\t"                                                                                     \
        ".reg .pred p;\n\t"                                                                         \
        "setp.ne.b32 p, %4, 0;\n\t"                                                                 \
        "tcgen05.mma.cta_group::" GROUP ".kind::" KINDSTR " [%0], %1, %2, %3, p;\n\t"               \
        "}" ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory")
template <int KIND, int CTAS>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if (KIND == KIND_TF32) {
        if (CTAS == 1) MORIG_UMMA("1", "tf32"); else MORIG_UMMA("2", "tf32");
    } else {
        if (CTAS == 1) MORIG_UMMA("1", "f16"); else MORIG_UMMA("2", "f16");
    }
}
#undef MORIG_UMMA
// asynchronous TMEM -> register load of 32 lanes x 16 columns; the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// waits for every tcgen05.ld of this thread; the "+r" operands tie every later use of both halves to this wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&a)[16], uint32_t (&b)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]),
                   "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
                 :: "memory");
}

// register -> TMEM store of 32 lanes x 16 columns, and the wait that retires every tcgen05.st of this thread
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
           "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one TMEM column (a RUNTIME column address: the one place where a row of the tile can be picked dynamically without
// indexing registers) -> one register per lane; issue + wait in one statement, so the value is final on return
__device__ __forceinline__ float tmem_ld1_sync(uint32_t taddr) {
    uint32_t r;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// shared-memory matrix descriptor (K-major, SWIZZLE_128B: 8-row x 128 B atoms, 1024 B apart):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major), [32,46) stride
//   byte offset >> 4 (next 8-row group), [46,48) version = 1 (Blackwell), [61,64) layout = 2 (SWIZZLE_128B).
// Split in its two 32-bit halves: only the low half depends on the address, and stepping 8 tf32 (32 bytes) along the
// swizzled row adds 2 to it.
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; }

// instruction descriptor: D = f32 (bits 4-5 = 1), A and B format (bits 7-9, 10-12: 0 = f16, 2 = tf32), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
template <int KIND> __device__ __forceinline__ uint32_t make_idesc(int m, int n) {
    const uint32_t fmt = (KIND == KIND_F16) ? 0u : 2u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// Activation scale of the fp16 kind: a power of two s with amax * s < 2^15 (gather mode: relu(P + Q) <= 2 amax), and
// the factor that undoes it (and the weight image's scale) in the epilogue.  TF32 needs neither.
template <int KIND>
__device__ __forceinline__ void operand_scales(const GemmP &p, bool gather, float &a_scale, float &inv) {
    a_scale = 1.f; inv = 1.f;
    if (KIND == KIND_F16) {
        const uint32_t bits = __float_as_uint(p.amax_in ? *p.amax_in : 1.f);
        int e = (int)((bits >> 23) & 0xffu);            // amax < 2^(e - 126)
        if (e == 0 || e == 255) e = 126;                // zero / denormal / non-finite source: no scaling
        int sh = (gather ? 14 : 15) - (e - 126);
        sh = sh < -100 ? -100 : (sh > 100 ? 100 : sh);
        a_scale = __uint_as_float((uint32_t)(sh + 127) << 23);
        inv = __uint_as_float((uint32_t)(127 - sh) << 23) * (p.w_inv_dev ? *p.w_inv_dev : p.w_inv);
    }
}
// packed fp32 pairs (sm_100 FADD2 / FMUL2): one instruction for two lanes of the producers' element-wise work
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return *reinterpret_cast<const float2 *>(&r);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return *reinterpret_cast<const float2 *>(&r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t *>(&a)), "l"(*reinterpret_cast<const uint64_t *>(&b)));
    return *reinterpret_cast<const float2 *>(&r);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t *>(&a)),
        "l"(*reinterpret_cast<const uint64_t *>(&b)), "l"(*reinterpret_cast<const uint64_t *>(&c)));
    return *reinterpret_cast<const float2 *>(&r);
}
__device__ __forceinline__ uint32_t pack_h2(float lo_k, float hi_k) {
    const __half2 h = __floats2half2_rn(lo_k, hi_k);    // lower k in the low half (K-major, little endian)
    return *reinterpret_cast<const uint32_t *>(&h);
}

struct TcP {
    GemmP g;                 // shared operand/epilogue description (W/ldw unused here)
    const float *Bblob;      // [n_tiles][nK][hi|lo][BN*32] pre-swizzled weight images
    int nK;                  // k-chunks of 32
    int ntn, ntm, frames;    // tile grid (ntm is an upper bound in gather mode)
    int stages;              // A (and, when streaming, B) ring depth
    int resident_b;          // 1: the whole weight image is loaded once and kept in shared memory
    long long *trace;        // debug timeline (CTA 0): [role 0..2][2048] (tag, clock64) pairs, or NULL
};

// role timeline for scripts/tc_trace.py; compiled in only with -DMORIG_TRACE (MORIG_TRACE=1 python -m morig_b200.build)
struct Tracer {
    long long *buf; int n;
    __device__ __forceinline__ void operator()(int tag) {
#ifdef MORIG_TRACE
        if (buf && n < 1024) { buf[2 * n] = tag; buf[2 * n + 1] = clock64(); ++n; }
#else
        (void)tag;
#endif
    }
};

// ================= producer warps: A stage images =================
// The producers work in ring UNITS of 32 k-columns: thread -> 4 consecutive k-columns (one float4) of 4 rows.  A TF32
// stage is one unit (16-byte hi and lo chunks), an FP16 stage two units (8-byte chunks).  Two register buffers rotate:
// store unit u from one buffer, publish the stage, then refill the same buffer with the loads of unit u + 2 -- so
// one to two units (16-32 KB per SM) are always in flight, and the proxy fence that precedes every publish never has
// to drain loads younger than a full unit.  In gather mode the relu(P[tgt] + Q[col]) combine happens at store time;
// gather indices are fetched one tile ahead.  Loads are unpredicated: rows past the end re-read the last valid row
// (their accumulators are never looked at) and k-columns past K re-read columns 0-3 (the weight image is zero there).
// `arrive(s)` publishes stage s to the MMA issuer.
//
// tile row of (role warp w, lane group g = lane >> 3, pass ps): rows {a, a+1, a+4, a+5}.  An 8-byte fp16 store covers
// half a swizzled row and bit 2 of the row decides which half of the banks it lands in, so this mix keeps every
// store instruction at its minimum of 2 (fp16) / 4 (tf32) wavefronts; adjacent rows mostly share their P[tgt] line.
__device__ __forceinline__ int producer_row(int w, int g, int ps) {
    const int b = w + PRODUCER_WARPS * ps;          // 32 blocks of 4 rows
    return 8 * (b >> 1) + 2 * (b & 1) + (g & 1) + 4 * (g >> 1);
}

template <int KIND, int AMODE, class Arrive>
__device__ __forceinline__ void producer_role(const GemmP &p, float a_scale, uint8_t *smem, uint32_t a_stride,
                                              uint32_t aux_addr, int S, int nK, int M, const TileMap &tm, int tid,
                                              int lane, Arrive arrive, long long *trace = nullptr) {
    constexpr int UPS = KindCfg<KIND>::UPS;
    constexpr int RPT = ROWS_PER_THREAD;
    constexpr bool GATHER = (AMODE == AMODE_GATHER);
    Tracer tr{(trace && blockIdx.x == 0 && tid == 0) ? trace : nullptr, 0};
    const int c = tid & 7;
    const int w = tid >> 5, g = (tid >> 3) & 3;
    const int nU = nK * UPS;                     // ring units per tile
    int s = 0;                                   // stage ring position
    uint32_t wait_ph = 1;                        // parity of "stage s is free" (passes on a fresh barrier)

    // shared-memory byte offset of this thread's chunk in each of its rows (unit 0 of a stage; unit 1 of an fp16
    // stage is the other half of the swizzled row: offset ^ 64)
    uint32_t soff[RPT];
#pragma unroll
    for (int ps = 0; ps < RPT; ++ps) {
        const int row = producer_row(w, g, ps);
        soff[ps] = (KIND == KIND_F16) ? (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + ((c & 1) << 3))
                                      : (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
    }

    // load cursor: (tile, unit) whose loads are issued next, row pointers of that tile, indices of the following one
    int lt = tm.first, lu = 0;
    const float *rp[RPT];
    const float *rq[GATHER ? RPT : 1];
    int ni[GATHER ? RPT : 1], nj[GATHER ? RPT : 1];
    const int last_row = M - 1;

    auto fetch_idx = [&](int t) {
        const int m0 = tm.decode(t).m0;
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            const int r = min(m0 + producer_row(w, g, ps), last_row);
            ni[GATHER ? ps : 0] = p.tgt[r];
            nj[GATHER ? ps : 0] = p.col[r];
        }
    };
    auto issue = [&](float4 (&pd)[RPT], float4 (&qd)[GATHER ? RPT : 1]) {
        if (lt >= tm.total) return;
        if (lu == 0) {
            const TileCoord tc_ = tm.decode(lt);
            if (GATHER) {
                if (lt == tm.first) fetch_idx(lt);
                const size_t fb = (size_t)tc_.frame * p.n_vtx_frame;
#pragma unroll
                for (int ps = 0; ps < RPT; ++ps) {
                    rp[ps] = p.P + (fb + ni[ps]) * (size_t)p.ldpq + 4 * c;
                    rq[ps] = p.Q + (fb + nj[ps]) * (size_t)p.ldpq + 4 * c;
                }
                if (lt + tm.step < tm.total) fetch_idx(lt + tm.step);      // consumed a tile later
            } else {
#pragma unroll
                for (int ps = 0; ps < RPT; ++ps)
                    rp[ps] = p.A + (size_t)min(tc_.m0 + producer_row(w, g, ps), last_row) * p.lda + 4 * c;
            }
        }
        const int koff = (lu * KC + 4 * c < p.K) ? lu * KC : -4 * c;
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            pd[ps] = *reinterpret_cast<const float4 *>(rp[ps] + koff);
            if (GATHER) qd[ps] = *reinterpret_cast<const float4 *>(rq[ps] + koff);
        }
        if (++lu == nU) { lu = 0; lt += tm.step; }
    };
    // US = unit inside the stage: with two rotating buffers and an even unit count per tile, buffer 0 always holds
    // unit 0 and buffer 1 unit 1 of an fp16 stage, so it is a compile-time constant
    // (a literal at both call sites: the lambda is inlined and `us` folds away)
    auto store_unit = [&](const int us_arg, const float4 (&pd)[RPT], const float4 (&qd)[GATHER ? RPT : 1]) {
        const int us = (KIND == KIND_F16) ? us_arg : 0;
        if (us == 0) {
            tr(1);
            mbar_wait(aux_addr + AUX_MMA_DONE + 8u * s, wait_ph);    // MMAs that read this stage one ring turn ago retired
            tr(2);
        }
        uint8_t *a_hi = smem + s * a_stride;
        uint8_t *a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            // two packed halves of the thread's four k-columns
            float2 va = make_float2(pd[ps].x, pd[ps].y), vb = make_float2(pd[ps].z, pd[ps].w);
            if (GATHER) {
                va = add2(va, make_float2(qd[ps].x, qd[ps].y));
                vb = add2(vb, make_float2(qd[ps].z, qd[ps].w));
                va.x = fmaxf(va.x, 0.f); va.y = fmaxf(va.y, 0.f); vb.x = fmaxf(vb.x, 0.f); vb.y = fmaxf(vb.y, 0.f);
            }
            if (KIND == KIND_F16) {
                const float2 s2 = make_float2(a_scale, a_scale);
                va = mul2(va, s2); vb = mul2(vb, s2);
            }
            // hi = x with the 13 low mantissa bits cleared: 11 significant bits, exact both as TF32 and (inside the
            // normal range the scale guarantees) as FP16; lo = x - hi is exact in fp32 and |lo| < 2^-10 |x|, so
            // hi*hi + hi*lo + lo*hi reproduces x*y to ~2^-21.
            const float2 ha = make_float2(tf32_hi(va.x), tf32_hi(va.y)), hb = make_float2(tf32_hi(vb.x), tf32_hi(vb.y));
            const float2 la = sub2(va, ha), lb = sub2(vb, hb);
            if (KIND == KIND_F16) {
                const uint32_t off = soff[ps] ^ (us ? 64u : 0u);     // unit 1 = the other half of the swizzled row
                *reinterpret_cast<uint2 *>(a_hi + off) = make_uint2(pack_h2(ha.x, ha.y), pack_h2(hb.x, hb.y));
                *reinterpret_cast<uint2 *>(a_lo + off) = make_uint2(pack_h2(la.x, la.y), pack_h2(lb.x, lb.y));
            } else {
                *reinterpret_cast<float4 *>(a_hi + soff[ps]) = make_float4(ha.x, ha.y, hb.x, hb.y);
                *reinterpret_cast<float4 *>(a_lo + soff[ps]) = make_float4(la.x, la.y, lb.x, lb.y);
            }
        }
        if (us == UPS - 1) {
            fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) arrive(s);
            tr(3);
            if (++s == S) { s = 0; wait_ph ^= 1; }
        }
    };

    // NB register buffers rotate (an even number, and a tile has an even number of units, so buffer b always holds unit
    // b & 1 of an fp16 stage).  Plain rows need 16 registers per buffer: four of them keep two whole stages of loads in
    // flight -- with the weight chunks no longer late (dedicated fetch warp) the dense layers wait for these loads.
    constexpr int NB = GATHER ? 2 : 4;
    float4 pb[NB][RPT];
    float4 qb[NB][GATHER ? RPT : 1];
    int remaining = tm.my_tiles() * nU;                       // units still to be stored
#pragma unroll
    for (int b = 0; b < NB; ++b) issue(pb[b], qb[b]);
    while (remaining > 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (remaining > b) {
                store_unit(b & 1, pb[b], qb[b]);
                issue(pb[b], qb[b]);
            }
        }
        remaining -= NB;
    }
}

// ---- gather producers of the fused EdgeConv (AMODE_GATHER): relu(P[tgt[e]] + Q[col[e]]) per CSR slot ---------------------
// The same ring protocol as producer_role, specialised so that nothing is computed per unit that can be known earlier:
//   * NU = K / 32 units per tile is a template parameter (H = 64 / 128 / 256 -> 2 / 4 / 8): the tile body is fully
//     unrolled, the unit's column offset is an immediate of the load instruction, the fp16 half-row (unit & 1) and the
//     register buffer are compile-time, the eight row pointers are formed once per tile;
//   * the ReLU rides on the fp32 -> fp16 conversion (cvt.rn.relu.f16x2.f32): hi = x with the low mantissa bits cleared and
//     lo = x - hi have the sign of x (truncation), so for x < 0 both convert to +0 and for x >= 0 nothing changes --
//     bit-identical to splitting relu(x), four FMNMX per 16-byte chunk cheaper;
//   * tile coordinates advance incrementally (TileIter).
// ncu (source view) had the old producers at ~200 instructions per unit and warp against ~80 of arithmetic; with the
// TMEM write-back epilogue the producers set the pace of both fused EdgeConv kernels.  Measured and dropped: a third
// register buffer (loads three units ahead) and an L2 prefetch of the rows of the tile after the next one (half of
// the kernel's L2 requests miss in a cold-cache ncu replay, but inside the step the operand is largely L2-resident):
// neither moved the kernels.
__device__ __forceinline__ uint32_t pack_h2_relu(float lo_k, float hi_k) {
    uint32_t r;                                          // upper half <- first source, lower half <- second
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_k), "f"(lo_k));
    return r;
}

template <int KIND, int NU, class Arrive>
__device__ __forceinline__ void producer_gather_role(const GemmP &p, float a_scale, uint8_t *smem, uint32_t a_stride,
                                                     uint32_t aux_addr, int S, int M, const TileMap &tm, int tid, int lane,
                                                     Arrive arrive, long long *trace = nullptr) {
    constexpr int UPS = KindCfg<KIND>::UPS;
    constexpr int RPT = ROWS_PER_THREAD;
    static_assert(NU >= 2 && NU % 2 == 0, "two register buffers alternate over an even number of units per tile");
    Tracer tr{(trace && blockIdx.x == 0 && tid == 0) ? trace : nullptr, 0};
    const int c = tid & 7;
    const int w = tid >> 5, g = (tid >> 3) & 3;
    int s = 0;                                   // stage ring position
    uint32_t wait_ph = 1;                        // parity of "stage s is free" (passes on a fresh barrier)

    uint32_t soff[RPT];                          // see producer_role
#pragma unroll
    for (int ps = 0; ps < RPT; ++ps) {
        const int row = producer_row(w, g, ps);
        soff[ps] = (KIND == KIND_F16) ? (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + ((c & 1) << 3))
                                      : (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
    }

    TileIter it;                                 // the tile whose units are being stored
    it.init(tm);
    if (!it.valid()) return;
    const int last_row = M - 1;
    int ni[RPT], nj[RPT];                        // gather indices, fetched one tile ahead
    const float *rp[RPT], *rq[RPT];              // this thread's 16-byte column of its four P / Q rows
    auto fetch_idx = [&](int m0) {               // volatile asm: the loads must issue HERE, a tile ahead of their use
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            const int r = min(m0 + producer_row(w, g, ps), last_row);      // rows past the end re-read the last valid row
            asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(ni[ps]) : "l"(p.tgt + r));
            asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(nj[ps]) : "l"(p.col + r));
        }
    };
    auto set_rows = [&](int frame) {             // element offsets fit 32 bits (checked by the launchers)
        const uint32_t fb = (uint32_t)frame * (uint32_t)p.n_vtx_frame;
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            rp[ps] = p.P + ((fb + (uint32_t)ni[ps]) * (uint32_t)p.ldpq + 4u * c);
            rq[ps] = p.Q + ((fb + (uint32_t)nj[ps]) * (uint32_t)p.ldpq + 4u * c);
        }
    };
    auto issue = [&](float4 (&pd)[RPT], float4 (&qd)[RPT], const int unit) {      // `unit` is a literal at every call site
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            pd[ps] = *reinterpret_cast<const float4 *>(rp[ps] + unit * KC);
            qd[ps] = *reinterpret_cast<const float4 *>(rq[ps] + unit * KC);
        }
    };
    auto store_unit = [&](const int us, const float4 (&pd)[RPT], const float4 (&qd)[RPT]) {   // `us` is a literal too
        if (us == 0) {
            tr(1);
            mbar_wait(aux_addr + AUX_MMA_DONE + 8u * s, wait_ph);    // MMAs that read this stage one ring turn ago retired
            tr(2);
        }
        uint8_t *a_hi = smem + s * a_stride;
        uint8_t *a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
        for (int ps = 0; ps < RPT; ++ps) {
            float2 va = add2(make_float2(pd[ps].x, pd[ps].y), make_float2(qd[ps].x, qd[ps].y));
            float2 vb = add2(make_float2(pd[ps].z, pd[ps].w), make_float2(qd[ps].z, qd[ps].w));
            if (KIND == KIND_F16) {
                const float2 s2 = make_float2(a_scale, a_scale);
                va = mul2(va, s2); vb = mul2(vb, s2);
                const float2 ha = make_float2(tf32_hi(va.x), tf32_hi(va.y)), hb = make_float2(tf32_hi(vb.x), tf32_hi(vb.y));
                const float2 la = sub2(va, ha), lb = sub2(vb, hb);
                const uint32_t off = soff[ps] ^ (us ? 64u : 0u);     // unit 1 = the other half of the swizzled row
                *reinterpret_cast<uint2 *>(a_hi + off) = make_uint2(pack_h2_relu(ha.x, ha.y), pack_h2_relu(hb.x, hb.y));
                *reinterpret_cast<uint2 *>(a_lo + off) = make_uint2(pack_h2_relu(la.x, la.y), pack_h2_relu(lb.x, lb.y));
            } else {
                va.x = fmaxf(va.x, 0.f); va.y = fmaxf(va.y, 0.f); vb.x = fmaxf(vb.x, 0.f); vb.y = fmaxf(vb.y, 0.f);
                const float2 ha = make_float2(tf32_hi(va.x), tf32_hi(va.y)), hb = make_float2(tf32_hi(vb.x), tf32_hi(vb.y));
                const float2 la = sub2(va, ha), lb = sub2(vb, hb);
                *reinterpret_cast<float4 *>(a_hi + soff[ps]) = make_float4(ha.x, ha.y, hb.x, hb.y);
                *reinterpret_cast<float4 *>(a_lo + soff[ps]) = make_float4(la.x, la.y, lb.x, lb.y);
            }
        }
        if (us == UPS - 1) {
            fence_proxy_async();                 // generic-proxy stores -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) arrive(s);
            tr(3);
            if (++s == S) { s = 0; wait_ph ^= 1; }
        }
    };

    float4 pb0[RPT], pb1[RPT], qb0[RPT], qb1[RPT];
    fetch_idx(it.m0());
    set_rows(it.frame);
    {
        TileIter nx = it;
        nx.next();
        if (nx.valid()) fetch_idx(nx.m0());
    }
    issue(pb0, qb0, 0);
    issue(pb1, qb1, 1);
    // Invariant at the top of the loop: units 0 / 1 of tile `it` are in flight in buffers 0 / 1, ni / nj hold the indices
    // of the tile after it.  Unit u is stored from buffer u & 1, which is then refilled with unit u + 2 -- of this tile
    // (same row pointers, immediate column offset) or, for the last two units, of the next tile, whose row pointers
    // replace the current ones as soon as every load of the current tile has been issued.
    for (;;) {
        bool next_valid = false;
#pragma unroll
        for (int lu = 0; lu < NU; ++lu) {
            if (lu & 1) store_unit(UPS - 1, pb1, qb1);
            else store_unit(0, pb0, qb0);
            if (lu + 2 < NU) {
                if (lu & 1) issue(pb1, qb1, lu + 2);
                else issue(pb0, qb0, lu + 2);
            } else {
                if (lu + 2 == NU) {
                    it.next();
                    next_valid = it.valid();
                    if (next_valid) {
                        set_rows(it.frame);
                        TileIter nx = it;
                        nx.next();
                        if (nx.valid()) fetch_idx(nx.m0());                // consumed a tile later
                    }
                }
                if (next_valid) {
                    if (lu & 1) issue(pb1, qb1, lu + 2 - NU);
                    else issue(pb0, qb0, lu + 2 - NU);
                }
            }
        }
        if (!next_valid) break;
    }
}

// ================= epilogue warps =================
// TMEM lane = output channel, TMEM column = row of the tile: tcgen05.ld gives every lane one channel and 16 consecutive
// rows per instruction.  A warp owns the 32 channels of its lane quarter and every second 32-row block of the tile.
// Rows of a block are walked from registers with compile-time indices; the segment structure of the block (bit masks
// of segment heads / tails from a ballot over the row keys) is warp-uniform, so restarting the running max and
// flushing a finished segment are uniform branches.  `release(buf)` hands the accumulator buffer back to the issuer.

// w[r] for a warp-uniform runtime r: a jump over 32 register moves (registers cannot be indexed dynamically)
__device__ __forceinline__ float pick32(const float (&w)[32], int r) {
    float x;
    switch (r) {
#define MORIG_PICK(i) case i: x = w[i]; break;
        MORIG_PICK(0) MORIG_PICK(1) MORIG_PICK(2) MORIG_PICK(3) MORIG_PICK(4) MORIG_PICK(5) MORIG_PICK(6) MORIG_PICK(7)
        MORIG_PICK(8) MORIG_PICK(9) MORIG_PICK(10) MORIG_PICK(11) MORIG_PICK(12) MORIG_PICK(13) MORIG_PICK(14)
        MORIG_PICK(15) MORIG_PICK(16) MORIG_PICK(17) MORIG_PICK(18) MORIG_PICK(19) MORIG_PICK(20) MORIG_PICK(21)
        MORIG_PICK(22) MORIG_PICK(23) MORIG_PICK(24) MORIG_PICK(25) MORIG_PICK(26) MORIG_PICK(27) MORIG_PICK(28)
        MORIG_PICK(29) MORIG_PICK(30)
#undef MORIG_PICK
        default: x = w[31]; break;
    }
    return x;
}

template <int CTAS, int BN, int EPI, class Release>
__device__ __forceinline__ void epilogue_role(const GemmP &p, float inv, uint32_t aux_addr, int *keys_smem,
                                              uint32_t tmem_base, int M, const TileMap &tm, int rank, int warp, int lane,
                                              Release release, long long *trace = nullptr) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int NH = (CTAS == 1) ? BN / 128 : 1;   // 128-channel accumulators per buffer in this CTA's TMEM
    constexpr int NRB = (CTAS == 1) ? 4 : 8;         // 32-row blocks per tile (128 rows, or the 256 rows of a pair)
    constexpr int RBW = NRB / 2;                     // row blocks per warp
    constexpr int BUF_COLS = (CTAS == 1) ? BN : 256; // TMEM columns per accumulator buffer
    constexpr int NBUF = 512 / BUF_COLS;             // accumulator buffers in flight (4 or 2)
    constexpr int NBUF_LOG = (NBUF == 4) ? 2 : 1;
    const int e = warp - PRODUCER_WARPS;             // epilogue warp 0..7
    const int q = warp & 3;                          // TMEM lane quarter this warp may read (hardware: warp id % 4)
    const int half = e >> 2;
    Tracer tr{(trace && blockIdx.x == 0 && e == 0 && lane == 0) ? trace + 2 * 2048 : nullptr, 0};
    const bool need_keys = (EPI == EPI_SEGMAX) || p.pool != nullptr || p.rowbias != nullptr;
    const bool has_rb = (EPI == EPI_STORE) && p.rowbias != nullptr;
    const bool want_max = (EPI == EPI_STORE) && p.pool != nullptr;
    const float relu_floor = ((EPI == EPI_SEGMAX) || p.relu) ? 0.f : neg_inf();

    // Row keys (CSR target of an edge row / graph of a vertex row) are fetched one tile ahead into registers, so that no
    // global-memory latency sits between two tiles, and parked in a per-warp shared-memory strip for the walk (lane =
    // row there, while the walk needs them by row index).  ext: lane 0 / 31 hold the key of the row before / after the
    // block (-2 / -1 outside the matrix: never equal to a key).
    int *kstrip = keys_smem + e * KEYS_PER_WARP;
    auto key_of = [&](int r) -> int {
        if (EPI == EPI_SEGMAX) return p.tgt[r];
        return p.batch ? (r / p.n_vtx) * p.n_graphs + p.batch[r % p.n_vtx] : 0;
    };
    int nkey[RBW], next_[RBW];
    auto load_keys = [&](int t) {
#pragma unroll
        for (int i = 0; i < RBW; ++i) { nkey[i] = -1; next_[i] = -1; }
        if (!need_keys || t >= tm.total) return;
        const int r0 = tm.decode(t).m0 - (CTAS == 2 ? rank * BM : 0) + half * 32;
#pragma unroll
        for (int i = 0; i < RBW; ++i) {
            const int r = r0 + 64 * i + lane;
            if (r < M) nkey[i] = key_of(r);
            if (EPI == EPI_SEGMAX) {
                if (lane == 0) next_[i] = (r > 0) ? ((r - 1 < M) ? key_of(r - 1) : -1) : -2;
                if (lane == 31 && r + 1 < M) next_[i] = key_of(r + 1);
            }
        }
    };
    load_keys(tm.first);
    float amax_l = 0.f;                              // max |stored value| seen by this lane (GemmP::amax_out)

    int li = 0;
    for (int t = tm.first; t < tm.total; t += tm.step, ++li) {
        const TileCoord tcd = tm.decode(t);
        const int buf = li & (NBUF - 1);
        const int row0 = tcd.m0 - (CTAS == 2 ? rank * BM : 0);
        const int n0 = tcd.n_tile * ((CTAS == 1) ? BN : 256) + (CTAS == 2 ? rank * 128 : 0) + q * 32 + lane;
        if (need_keys) {
            __syncwarp();                            // the previous tile's reads of the strip are done
#pragma unroll
            for (int i = 0; i < RBW; ++i) {
                kstrip[32 * i + lane] = nkey[i];
                if (lane == 0) kstrip[4 * 32 + 2 * i] = next_[i];
                if (lane == 31) kstrip[4 * 32 + 2 * i + 1] = next_[i];
            }
            __syncwarp();
            load_keys(t + tm.step);                  // in flight during this tile
        }
        const uint32_t frame_base = (EPI == EPI_SEGMAX) ? (uint32_t)(tcd.frame * p.n_vtx_frame) : 0u;
        // per-channel constants of the tile's channels: requested BEFORE the wait for the accumulator, so that their
        // L2 latency hides behind it (loaded per 32-row unit they cost the short-K layers ~400 cycles per unit)
        float bias_t[NH], scale_t[NH], shift_t[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const int nl = n0 + h * 128;
            const bool ok = nl < p.N;
            bias_t[h] = (ok && p.bias) ? p.bias[nl] : 0.f;
            scale_t[h] = (ok && p.scale) ? p.scale[nl] : 1.f;
            shift_t[h] = (ok && p.shift) ? p.shift[nl] : 0.f;
        }

        tr(10);
        mbar_wait(aux_addr + AUX_ACC_FULL + 8u * buf, (uint32_t)((li >> NBUF_LOG) & 1));
        tr(11);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BUF_COLS);
        uint32_t v[32];
        uint32_t (&va)[16] = reinterpret_cast<uint32_t (&)[16]>(v[0]);
        uint32_t (&vb)[16] = reinterpret_cast<uint32_t (&)[16]>(v[16]);
        tmem_ld16_issue(tbase + (uint32_t)(half * 32), va);
        tmem_ld16_issue(tbase + (uint32_t)(half * 32 + 16), vb);
        // The unit loop is deliberately NOT unrolled: twenty warps of three roles share the instruction cache.
#pragma unroll 1
        for (int u = 0; u < RBW * NH; ++u) {
            const int i = u / NH, h = u - i * NH;
            const int rb = half + 2 * i;
            const int rbase = row0 + rb * 32;        // first row of the block
            const int valid_rows = min(32, max(0, M - rbase));
            const int nl = n0 + h * 128;             // this lane's output channel
            const bool nl_ok = nl < p.N;
            const float bias_l = (NH == 1 || h == 0) ? bias_t[0] : bias_t[NH - 1];
            const float scale_l = (NH == 1 || h == 0) ? scale_t[0] : scale_t[NH - 1];
            const float shift_l = (NH == 1 || h == 0) ? shift_t[0] : shift_t[NH - 1];
            // segment structure of the block: warp-uniform masks
            const int *kblk = kstrip + 32 * i;
            uint32_t tails = 0x80000000u;
            bool first_cut = false, last_cut = false;
            int first_tail = 31;
            if (need_keys) {
                const int key = kblk[lane];
                const int key_dn = __shfl_down_sync(FULL, key, 1);
                tails = __ballot_sync(FULL, (lane == 31) || (key != key_dn));
                if (EPI == EPI_SEGMAX) {
                    first_cut = kblk[0] == kstrip[4 * 32 + 2 * i];           // segment continues from the previous block
                    last_cut = kblk[31] == kstrip[4 * 32 + 2 * i + 1];       // ... into the next block
                    first_tail = __ffs(tails) - 1;
                }
            }
            const uint32_t heads = (tails << 1) | 1u;
            float *crow = (EPI == EPI_STORE && p.C && nl_ok) ? p.C + (size_t)rbase * p.ldc + nl : nullptr;
            // output offsets fit 32 bits (checked by the launchers)
            const float sinv_l = scale_l < 0.f ? -inv : inv;
            auto flush = [&](int r, float m) {       // the segment ending at row r is complete (warp-uniform call)
                const int k_seg = kblk[r];
                if (k_seg < 0 || !nl_ok) return;
                if (EPI == EPI_SEGMAX) {
                    // m = max over the segment's rows of the raw accumulator = sigma * (h W1'), sigma = sign of the
                    // BatchNorm scale (folded into the weight image).  bias -> ReLU -> BatchNorm affine is monotone in
                    // sigma * z (non-decreasing for scale >= 0; for scale < 0 the extreme is the minimum of z, i.e. the
                    // maximum of -z), also in floating point, so applying it ONCE to the segment's extreme gives bit for
                    // bit the maximum of the per-edge values -- and the per-element epilogue shrinks to the running max.
                    m = fmaf(fmaxf(fmaf(m, sinv_l, bias_l), 0.f), scale_l, shift_l);
                    float *dst = p.C + ((frame_base + (uint32_t)k_seg) * (uint32_t)p.ldc + (uint32_t)nl);
                    if ((first_cut && r == first_tail) || (last_cut && r == 31)) atomic_max_f32(dst, m);
                    else *dst = m;
                    amax_l = fmaxf(amax_l, fabsf(m));                         // only the maxima are stored
                } else {
                    atomic_max_f32(p.pool + ((uint32_t)k_seg * (uint32_t)p.ldpool + (uint32_t)nl), m);
                }
            };
            auto head_bias = [&](int r) -> float {   // per-graph bias of the segment starting at row r (uniform call)
                const int k_seg = kblk[r];
                return (has_rb && k_seg >= 0 && nl_ok) ? p.rowbias[(uint32_t)k_seg * (uint32_t)p.ldrb + (uint32_t)nl] : 0.f;
            };
            tmem_ld_wait(va, vb);
            tr(13);
            float w[32];
            if (EPI == EPI_SEGMAX) {
                // Straight-line code over the RAW accumulators (see flush): the running max restarts at segment heads;
                // the value after the last row of a segment is its extreme.
#pragma unroll
                for (int r = 0; r < 32; ++r) w[r] = __uint_as_float(v[r]);
#pragma unroll
                for (int r = 1; r < 32; ++r) w[r] = ((heads >> r) & 1u) ? w[r] : fmaxf(w[r - 1], w[r]);
                for (uint32_t tl = tails; tl; tl &= tl - 1) {
                    const int r = __ffs(tl) - 1;
                    flush(r, pick32(w, r));
                }
            } else if (tails == 0x80000000u && valid_rows == 32) {
                // dense layer, one segment (graph) in the block: straight-line stores, one pooled max at the end
                const float b_cur = bias_l + (has_rb ? head_bias(0) : 0.f);
                float am0 = 0.f, am1 = 0.f;
                const float2 inv2 = make_float2(inv, inv), bias2 = make_float2(b_cur, b_cur);
                const float2 scale2 = make_float2(scale_l, scale_l), shift2 = make_float2(shift_l, shift_l);
#pragma unroll
                for (int r = 0; r < 32; r += 2) {             // two rows per packed FFMA2
                    float2 x = fma2(make_float2(__uint_as_float(v[r]), __uint_as_float(v[r + 1])), inv2, bias2);
                    x.x = fmaxf(x.x, relu_floor); x.y = fmaxf(x.y, relu_floor);
                    x = fma2(x, scale2, shift2);
                    w[r] = x.x; w[r + 1] = x.y;
                    am0 = fmaxf(am0, fabsf(x.x)); am1 = fmaxf(am1, fabsf(x.y));
                    if (crow) { crow[(size_t)r * p.ldc] = x.x; crow[(size_t)(r + 1) * p.ldc] = x.y; }
                }
                if (nl_ok) amax_l = fmaxf(amax_l, fmaxf(am0, am1));
                if (want_max) {
#pragma unroll
                    for (int st = 16; st > 0; st >>= 1)
#pragma unroll
                        for (int r = 0; r < st; ++r) w[r] = fmaxf(w[r], w[r + st]);
                    flush(31, w[0]);
                }
            } else {
                // dense layer, general block (ragged last tile, block straddling two graphs): row by row
                float b_cur = bias_l, m = neg_inf(), am = 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    if ((heads >> r) & 1u) b_cur = bias_l + head_bias(r);       // uniform branch
                    const float z = fmaf(fmaxf(fmaf(__uint_as_float(v[r]), inv, b_cur), relu_floor), scale_l, shift_l);
                    if (r < valid_rows) {
                        am = fmaxf(am, fabsf(z));
                        if (crow) crow[(size_t)r * p.ldc] = z;
                    }
                    if (want_max) {
                        m = ((heads >> r) & 1u) ? z : fmaxf(m, z);
                        if ((tails >> r) & 1u) flush(r, m);                   // uniform branch
                    }
                }
                if (nl_ok) amax_l = fmaxf(amax_l, am);
            }
            if (u + 1 < RBW * NH) {
                // next unit: same row block, next channel half -- or the next row block
                const int un = u + 1, in_ = un / NH, hn = un - in_ * NH;
                const uint32_t tnext = tbase + (uint32_t)(hn * 128 + (half + 2 * in_) * 32);
                tmem_ld16_issue(tnext, va);
                tmem_ld16_issue(tnext + 16u, vb);
            }
            tr(16);
        }
        tr(17);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release(buf);             // buffer may be overwritten by tile li + NBUF
        tr(12);
    }
    if (EPI == EPI_SEGMAX || p.C) amax_commit(p.amax_out, amax_l);
}

// ---- fused-EdgeConv epilogue (EPI_SEGMAX): segmented max over the CSR target of the tile's rows -------------------
// Same TMEM view as above (lane = channel, column = edge row).  A warp owns the 32 channels of its lane quarter and a
// CONTIGUOUS range of RBW 32-row blocks (half the tile), so that a segment running across its blocks is carried in a
// register and only segments that leave the warp's range need the atomic merge.  Per 32 x 32 block:
//   1. tcgen05.ld the raw accumulators; running max along the rows IN REGISTERS (one predicated FMNMX per row, the
//      head mask is warp-uniform), seeded with the carry when the block continues the previous block's last segment;
//   2. tcgen05.st the running maxima back over the accumulator columns (the buffer is ours until it is released);
//   3. for every segment tail r of the block (2-5 per block): tcgen05.ld of ONE column at the runtime address of r --
//      TMEM is addressable by a register, registers are not -- then bias -> ReLU -> BatchNorm once per segment (see
//      the monotonicity argument at `flush` in epilogue_role) and one coalesced 128-byte row store.
// An earlier version picked the tail's value out of 32 registers with a branch tree: ~450 cycles per tail (ncu: branch
// resolution + instruction-cache misses), 2100-2500 cycles per block, which made the epilogue -- not the MMAs or the
// gathers -- the pace-setter of both fused EdgeConv kernels (role timeline, scripts/tc_trace.py).
template <int CTAS, int BN, class Release>
__device__ __forceinline__ void epilogue_segmax_role(const GemmP &p, float inv, uint32_t aux_addr, int *keys_smem,
                                                     uint32_t tmem_base, int M, const TileMap &tm, int rank, int warp,
                                                     int lane, Release release, long long *trace = nullptr) {
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int NH = (CTAS == 1) ? BN / 128 : 1;   // 128-channel accumulators per buffer in this CTA's TMEM
    constexpr int NRB = (CTAS == 1) ? 4 : 8;         // 32-row blocks per tile
    constexpr int RBW = NRB / 2;                     // contiguous row blocks per warp
    constexpr int BUF_COLS = (CTAS == 1) ? BN : 256;
    constexpr int NBUF = 512 / BUF_COLS;
    constexpr int NBUF_LOG = (NBUF == 4) ? 2 : 1;
    static_assert(RBW * 32 + 2 <= KEYS_PER_WARP, "key strip too small");
    const int e = warp - PRODUCER_WARPS;             // epilogue warp 0..7
    const int q = warp & 3;                          // TMEM lane quarter this warp may access (hardware: warp id % 4)
    const int half = e >> 2;                         // which half of the tile's rows
    Tracer tr{(trace && blockIdx.x == 0 && e == 0 && lane == 0) ? trace + 2 * 2048 : nullptr, 0};

    // Row keys (CSR targets) of the warp's rows are fetched one tile ahead into registers and parked in a per-warp
    // shared-memory strip for the walk: strip[32 i + lane] = key of row lane of block i; strip[128] / strip[129] = key of
    // the row before / after the warp's range (-2 / -1 outside the matrix: never equal to a real key).
    int *kstrip = keys_smem + e * KEYS_PER_WARP;
    int nkey[RBW], kext;
    auto load_keys = [&](int t) {
#pragma unroll
        for (int i = 0; i < RBW; ++i) nkey[i] = -1;
        kext = -1;
        if (t >= tm.total) return;
        const int r0 = tm.decode(t).m0 - (CTAS == 2 ? rank * BM : 0) + half * (32 * RBW);
#pragma unroll
        for (int i = 0; i < RBW; ++i) {
            const int r = r0 + 32 * i + lane;
            if (r < M) nkey[i] = p.tgt[r];
        }
        if (lane == 0) kext = (r0 > 0) ? ((r0 - 1 < M) ? p.tgt[r0 - 1] : -1) : -2;
        if (lane == 31 && r0 + 32 * RBW < M) kext = p.tgt[r0 + 32 * RBW];
    };
    load_keys(tm.first);
    float amax_l = 0.f;                              // max |stored value| seen by this lane (GemmP::amax_out)

    int li = 0;
    for (int t = tm.first; t < tm.total; t += tm.step, ++li) {
        const TileCoord tcd = tm.decode(t);
        const int buf = li & (NBUF - 1);
        const int n0 = tcd.n_tile * ((CTAS == 1) ? BN : 256) + (CTAS == 2 ? rank * 128 : 0) + q * 32 + lane;
        __syncwarp();                                // the previous tile's reads of the strip are done
#pragma unroll
        for (int i = 0; i < RBW; ++i) kstrip[32 * i + lane] = nkey[i];
        if (lane == 0) kstrip[128] = kext;
        if (lane == 31) kstrip[129] = kext;
        __syncwarp();
        load_keys(t + tm.step);                      // in flight during this tile
        // per-channel constants and the output column of this lane (L1 / L2 hits; their latency hides behind the wait
        // for the accumulator).  sigma = sign of the BatchNorm scale is folded into the weight image, see `flush`.
        // (NH = 2, the small-batch fallback, reloads them per unit instead of holding two sets in 72 registers.)
        struct LaneConsts { float bias, scale, shift, sinv; float *cb; };
        auto lane_consts = [&](int h) {
            LaneConsts k;
            const int nl = n0 + h * 128;
            const bool ok = nl < p.N;
            k.bias = ok ? p.bias[nl] : 0.f;
            k.scale = ok ? p.scale[nl] : 1.f;
            k.shift = ok ? p.shift[nl] : 0.f;
            k.sinv = k.scale < 0.f ? -inv : inv;
            // output offsets fit 32 bits (checked by the launchers)
            k.cb = ok ? p.C + ((uint32_t)(tcd.frame * p.n_vtx_frame) * (uint32_t)p.ldc + (uint32_t)nl) : nullptr;
            return k;
        };
        LaneConsts k0 = lane_consts(0);

        tr(10);
        mbar_wait(aux_addr + AUX_ACC_FULL + 8u * buf, (uint32_t)((li >> NBUF_LOG) & 1));
        tr(11);
        tc_fence_after();
        // TMEM address of (lane quarter, accumulator buffer, first row of the warp's range)
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BUF_COLS + half * (32 * RBW));
        uint32_t v[32];
        uint32_t (&va)[16] = reinterpret_cast<uint32_t (&)[16]>(v[0]);
        uint32_t (&vb)[16] = reinterpret_cast<uint32_t (&)[16]>(v[16]);
        tmem_ld16_issue(tbase, va);
        tmem_ld16_issue(tbase + 16u, vb);
        float carry[NH];                             // running max of the segment that is open at the end of a block
#pragma unroll
        for (int h = 0; h < NH; ++h) carry[h] = 0.f;
        bool open_cut = false;                       // the segment open at the START of the block began before the range
        // The unit loop is deliberately NOT unrolled: twenty warps of three roles share the instruction cache.
#pragma unroll 1
        for (int i = 0; i < RBW; ++i) {
            // segment structure of block i: warp-uniform masks (identical for both channel halves)
            const int *kblk = kstrip + 32 * i;
            const int key = kblk[lane];
            const int key_dn = __shfl_down_sync(FULL, key, 1);
            const uint32_t tails = __ballot_sync(FULL, (lane == 31) || (key != key_dn));
            const uint32_t heads = (tails << 1) | 1u;
            const bool first_cut = kblk[0] == ((i == 0) ? kstrip[128] : kblk[-1]);            // continues from the previous row
            const bool last_cut = kblk[31] == ((i == RBW - 1) ? kstrip[129] : kblk[32]);      // ... into the next row
            const int first_tail = __ffs(tails) - 1;
            const bool cont = (i > 0) && first_cut;          // row 0 continues the segment carried in `carry`
            if (i == 0) open_cut = first_cut;
            const bool defer = last_cut && (i < RBW - 1);    // the last segment goes on in this warp's next block
#pragma unroll
            for (int h = 0; h < NH; ++h) {                   // NH = 2 only in the small-batch cta_group::1 fallback of H = 256
            const uint32_t tcol = tbase + (uint32_t)(h * 128 + i * 32);

            tmem_ld_wait(va, vb);
            tr(13);
            // running max along the rows, restarted at segment heads (uniform predicates)
            float w[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) w[r] = __uint_as_float(v[r]);
            if (cont) w[0] = fmaxf(carry[h], w[0]);
#pragma unroll
            for (int r = 1; r < 32; ++r)
                if (!((heads >> r) & 1u)) w[r] = fmaxf(w[r - 1], w[r]);
            if (defer) carry[h] = w[31];
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = __float_as_uint(w[r]);
            tmem_st16(tcol, va);
            tmem_st16(tcol + 16u, vb);
            tmem_st_wait();
            // flush the finished segments: the value after the last row of a segment is its extreme
            const LaneConsts kh = (NH == 1) ? k0 : lane_consts(h);
            const float bias_h = kh.bias, scale_h = kh.scale, shift_h = kh.shift, sinv_h = kh.sinv;
            float *const cb = kh.cb;
            for (uint32_t tl = defer ? (tails & 0x7fffffffu) : tails; tl; tl &= tl - 1) {
                const int r = __ffs(tl) - 1;
                const int k_seg = kblk[r];
                float m = tmem_ld1_sync(tcol + (uint32_t)r);
                const bool ok = k_seg >= 0 && cb != nullptr;     // not: rows past the end of the matrix / channels past N
                m = fmaf(fmaxf(fmaf(m, sinv_h, bias_h), 0.f), scale_h, shift_h);
                float *dst = cb + (uint32_t)(k_seg < 0 ? 0 : k_seg) * (uint32_t)p.ldc;
                // a segment that leaves the warp's range is merged with the ordered-int atomic max (exact and order
                // independent, hence deterministic; same rule as atomic_max_f32), everything else is a plain store.
                // Predicated, not branched: the lanes of a warp differ in `ok` and in the sign, and a divergent region
                // between two warp-synchronous tcgen05.ld costs a reconvergence barrier per tail.
                const bool at = (r == first_tail && first_cut && open_cut) || (r == 31 && last_cut);
                const float v = m + 0.0f;                        // canonicalise -0.0
                const int p_st = (ok && !at) ? 1 : 0, p_mx = (ok && at && v >= 0.0f) ? 1 : 0, p_mn = (ok && at && !(v >= 0.0f)) ? 1 : 0;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p0, p1, p2;\n\t"
                    "setp.ne.s32 p0, %2, 0;\n\t"
                    "setp.ne.s32 p1, %3, 0;\n\t"
                    "setp.ne.s32 p2, %4, 0;\n\t"
                    "@p0 st.global.f32 [%0], %1;\n\t"
                    "@p1 red.global.max.s32 [%0], %5;\n\t"
                    "@p2 red.global.min.u32 [%0], %5;\n\t"
                    "}" ::"l"(dst), "f"(m), "r"(p_st), "r"(p_mx), "r"(p_mn), "r"(__float_as_int(v)) : "memory");
                amax_l = ok ? fmaxf(amax_l, fabsf(m)) : amax_l;  // only the maxima are stored
            }
            if (h + 1 < NH || i + 1 < RBW) {                  // next unit: the other channel half, or the next row block
                const uint32_t tnext = tbase + (uint32_t)((h + 1 < NH) ? (h + 1) * 128 + i * 32 : (i + 1) * 32);
                tmem_ld16_issue(tnext, va);
                tmem_ld16_issue(tnext + 16u, vb);
            }
            tr(16);
            }
            open_cut = defer && (first_tail == 31) && first_cut && open_cut;
        }
        tr(17);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release(buf);             // buffer may be overwritten by tile li + NBUF
        tr(12);
    }
    amax_commit(p.amax_out, amax_l);
}

// =============================================================================================================
// cta_group::1 kernel: BN / 128 UMMAs of 128 (channels) x 128 (rows) per k-step
// =============================================================================================================
template <int KIND, int BN, int AMODE, int EPI>
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
    // stage s: A image at s * a_stride; B image of the stage (streaming) or of k-chunk kc (resident)
    const uint32_t a_stride = resb ? (uint32_t)A_STAGE_BYTES : (uint32_t)C::STAGE_BYTES;
    const uint32_t b_region = resb ? (uint32_t)(S * A_STAGE_BYTES) : (uint32_t)A_STAGE_BYTES;
    const uint32_t b_stride = resb ? (uint32_t)C::B_CHUNK_BYTES : (uint32_t)C::STAGE_BYTES;
    uint8_t *aux = smem + PIPE_BYTES;
    const uint32_t aux_addr = base + PIPE_BYTES;
    auto bar_a = [&](int s) { return aux_addr + AUX_A_FULL + 8u * s; };
    auto bar_b = [&](int s) { return aux_addr + AUX_B_FULL + 8u * s; };
    auto bar_m = [&](int s) { return aux_addr + AUX_MMA_DONE + 8u * s; };
    auto bar_accf = [&](int b) { return aux_addr + AUX_ACC_FULL + 8u * b; };
    auto bar_acce = [&](int b) { return aux_addr + AUX_ACC_EMPTY + 8u * b; };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aux + AUX_TMEM_PTR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_trigger();                                   // the next kernel of the stream may start its own set-up

    if (warp == CONTROL_WARP) {
        if (lane == 0) {
            for (int s = 0; s < 4; ++s) {
                mbar_init(bar_a(s), PRODUCER_WARPS);
                mbar_init(bar_b(s), 1);
                mbar_init(bar_m(s), 1);
            }
            for (int b = 0; b < C::NBUF; ++b) {
                mbar_init(bar_accf(b), 1);
                mbar_init(bar_acce(b), EPILOGUE_WARPS);      // every epilogue warp releases the buffer
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<1>(aux_addr + AUX_TMEM_PTR, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above only touched shared memory / TMEM: from here on the predecessor's results are needed
    pdl_wait();
    int M = p.M, ntm = tp.ntm;
    if (AMODE == AMODE_GATHER) {
        M = p.rowptr[p.n_vtx_frame];                 // E' lives on the device only
        ntm = (M + BM - 1) / BM;
    }
    TileMap tm;
    tm.ntn = tp.ntn; tm.ntm = ntm; tm.total = tp.ntn * ntm * tp.frames;
    tm.first = blockIdx.x; tm.step = gridDim.x; tm.mult = 1; tm.rank = 0;
    float a_scale, inv;
    operand_scales<KIND>(p, AMODE == AMODE_GATHER, a_scale, inv);

    if (warp >= CONTROL_WARP) {
        // ================= control warp: B bulk copies + MMA issue (one elected lane) =================
        reg_dec<REGS_CONTROL>();
        // The whole control warp runs the loop (uniform control flow keeps the address arithmetic on the uniform
        // datapath); only the elected lane issues the asynchronous operations.
        if (warp == CONTROL_WARP) {
            const bool leader = lane == 0;
            Tracer tr{(tp.trace && blockIdx.x == 0 && leader) ? tp.trace + 2048 : nullptr, 0};
            const uint32_t idesc = make_idesc<KIND>(128, BM);      // M = 128 output channels, N = the tile's 128 rows
            const uint32_t b_bytes = (uint32_t)C::B_CHUNK_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob);
            const int my_tiles = tm.my_tiles();
            const int f_ntile = (my_tiles > 0) ? tm.decode(tm.first).n_tile : 0;
            if (resb) {
                // one n-tile for the whole kernel: fetch every k-chunk of the image once
                if (my_tiles > 0) {
                    if (leader) {
                        mbar_arrive_expect_tx(bar_b(0), b_bytes * (uint32_t)nK);
                        for (int kc = 0; kc < nK; ++kc)
                            bulk_g2s(base + b_region + kc * b_stride, gB + ((size_t)f_ntile * nK + kc) * b_bytes, b_bytes,
                                     bar_b(0));
                    }
                    mbar_wait(bar_b(0), 0);
                }
            }
            // (streaming mode: the weight chunks are requested by the dedicated warp below, see tc2_gemm_kernel)
            int s = 0;
            uint32_t ph = 0;
            for (int li = 0; li < my_tiles; ++li) {
                const int buf = li & (C::NBUF - 1);
                tr(20);
                mbar_wait(bar_acce(buf), ((li / C::NBUF) & 1) ^ 1);     // accumulator buffer drained by the epilogue
                tr(21);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                for (int kc = 0; kc < nK; ++kc) {
                    mbar_wait(bar_a(s), ph);
                    tr(22);
                    if (!resb) mbar_wait(bar_b(s), ph);
                    tr(23);
                    tc_fence_after();
                    const uint32_t a_hi = base + s * a_stride;
                    const uint32_t b_hi = base + b_region + (resb ? kc : s) * b_stride;
                    const uint32_t lah = desc_lo(a_hi), lal = desc_lo(a_hi + A_HALF_BYTES);
                    const uint32_t lbh = desc_lo(b_hi), lbl = desc_lo(b_hi + C::B_HALF_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                // 8 tf32 / 16 fp16 = 32 bytes along the swizzled row
#pragma unroll
                            for (int h = 0; h < C::NH; ++h) {        // weights = A operand (rows 128 h .. of the image)
                                const uint32_t wo = (uint32_t)(h * 128 * 128) >> 4;
                                const uint32_t td = tmem_d + (uint32_t)(h * 128);
                                umma<KIND, 1>(td, desc64(lbh + wo + 2 * k), desc64(lal + 2 * k), idesc, (kc | k) != 0);
                                umma<KIND, 1>(td, desc64(lbl + wo + 2 * k), desc64(lah + 2 * k), idesc, 1);
                                umma<KIND, 1>(td, desc64(lbh + wo + 2 * k), desc64(lah + 2 * k), idesc, 1);
                            }
                        }
                        umma_commit<1>(bar_m(s));                    // frees stage s when these MMAs retire
                    }
                    tr(24);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                if (leader) umma_commit<1>(bar_accf(buf));       // accumulator complete -> epilogue
            }
        } else if (warp == CONTROL_WARP + 1 && !resb) {
            // ================= weight-chunk fetcher (streaming mode): follows the ring, never the MMA issuer =================
            const bool leader = lane == 0;
            const uint32_t b_bytes = (uint32_t)C::B_CHUNK_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob);
            const int my_tiles = tm.my_tiles();
            int s = 0;
            uint32_t free_ph = 1;                                  // passes on the fresh barriers of the first ring turn
            for (int li = 0; li < my_tiles; ++li) {
                const int n_tile = tm.decode(tm.first + li * tm.step).n_tile;
                for (int kc = 0; kc < nK; ++kc) {
                    mbar_wait(bar_m(s), free_ph);                  // the MMAs that read this slot one ring turn ago retired
                    if (leader) {
                        mbar_arrive_expect_tx(bar_b(s), b_bytes);
                        bulk_g2s(base + b_region + s * b_stride, gB + ((size_t)n_tile * nK + kc) * b_bytes, b_bytes, bar_b(s));
                    }
                    if (++s == S) { s = 0; free_ph ^= 1; }
                }
            }
        }
    } else if (warp < PRODUCER_WARPS) {
        reg_inc<RoleCfg<AMODE>::REGS_PRODUCER>();
        auto arrive = [&](int s) { mbar_arrive(bar_a(s)); };
        if constexpr (AMODE == AMODE_GATHER) {           // K = H in {64, 128} (BN = 128) or 256: K / 32 units per tile
            if constexpr (BN == 128) {
                if (p.K <= 64) producer_gather_role<KIND, 2>(p, a_scale, smem, a_stride, aux_addr, S, M, tm, tid, lane, arrive, tp.trace);
                else producer_gather_role<KIND, 4>(p, a_scale, smem, a_stride, aux_addr, S, M, tm, tid, lane, arrive, tp.trace);
            } else {
                producer_gather_role<KIND, 8>(p, a_scale, smem, a_stride, aux_addr, S, M, tm, tid, lane, arrive, tp.trace);
            }
        } else {
            producer_role<KIND, AMODE>(p, a_scale, smem, a_stride, aux_addr, S, nK, M, tm, tid, lane, arrive, tp.trace);
        }
    } else {
        if constexpr (RoleCfg<AMODE>::REGS_EPILOGUE >= 96) reg_inc<RoleCfg<AMODE>::REGS_EPILOGUE>();
        else reg_dec<RoleCfg<AMODE>::REGS_EPILOGUE>();
        auto release = [&](int b) { mbar_arrive(bar_acce(b)); };
        if constexpr (EPI == EPI_SEGMAX)
            epilogue_segmax_role<1, BN>(p, inv, aux_addr, reinterpret_cast<int *>(aux + AUX_BYTES), tmem_base, M, tm, 0, warp,
                                        lane, release, tp.trace);
        else
            epilogue_role<1, BN, EPI>(p, inv, aux_addr, reinterpret_cast<int *>(aux + AUX_BYTES), tmem_base, M, tm, 0, warp,
                                      lane, release, tp.trace);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CONTROL_WARP) tmem_dealloc<1>(tmem_base, C::TMEM_COLS);
}

// =============================================================================================================
// cta_group::2 kernel: 2-CTA cluster, UMMA 256 (channels) x 256 (rows).  CTA rank r of the pair holds channels
// [128 r, 128 r + 128) of the n-tile (rows of every weight chunk) in its TMEM lanes and produces rows (2*mp + r)*128..
// of the pair-tile.  Only the leader (rank 0) issues MMAs; its barriers collect the arrivals of both CTAs, and
// tcgen05.commit multicasts completions to both.
// =============================================================================================================
template <int BN2> struct Cfg2 {
    static constexpr int HALF_B = (BN2 / 2) * 128;                              // this CTA's rows of the hi (or lo) image
    static constexpr int B_BYTES = 2 * HALF_B;                                  // this CTA's share of one weight chunk
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_BYTES;                 // 32 KB A + half weight chunk
    static constexpr int STAGES = (PIPE_BYTES / STAGE_BYTES) > 4 ? 4 : (PIPE_BYTES / STAGE_BYTES);   // 3 (BN 256) / 4 (BN 128)
    // resident-B mode: the pair sees one n-tile for the whole kernel and each CTA keeps its half of all k-chunks
    static constexpr int res_stages(int nK) {
        const int left = PIPE_BYTES - nK * B_BYTES;
        const int s = left / A_STAGE_BYTES;
        return s > 4 ? 4 : s;
    }
};

template <int KIND, int BN2, int AMODE, int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc2_gemm_kernel(const TcP tp) {
    static_assert(BN2 == 256, "the pair kernel needs 128 channels (TMEM lanes) per CTA");
    using C2 = Cfg2<BN2>;
    const GemmP &p = tp.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (base - raw_addr);
    const int nK = tp.nK;
    const int S = tp.stages;
    const bool resb = tp.resident_b != 0;
    constexpr uint32_t HALF_B = C2::HALF_B;
    // stage s: A image at s * a_stride; this CTA's half weight chunk behind it (streaming) or, resident, chunk kc
    // at b_region + kc * B_BYTES
    const uint32_t a_stride = resb ? (uint32_t)A_STAGE_BYTES : (uint32_t)C2::STAGE_BYTES;
    const uint32_t b_region = resb ? (uint32_t)(S * A_STAGE_BYTES) : (uint32_t)A_STAGE_BYTES;
    const uint32_t b_stride = resb ? (uint32_t)C2::B_BYTES : (uint32_t)C2::STAGE_BYTES;
    uint8_t *aux = smem + PIPE_BYTES;
    const uint32_t aux_addr = base + PIPE_BYTES;
    auto bar_a = [&](int s) { return aux_addr + AUX_A_FULL + 8u * s; };        // leader: 8 producer warps of the pair
    auto bar_b = [&](int s) { return aux_addr + AUX_B_FULL + 8u * s; };        // local: this CTA's half chunk landed
    auto bar_bp = [&](int s) { return aux_addr + AUX_B_PEER + 8u * s; };       // leader: the peer's half landed
    auto bar_m = [&](int s) { return aux_addr + AUX_MMA_DONE + 8u * s; };      // local copy of the multicast commit
    auto bar_accf = [&](int b) { return aux_addr + AUX_ACC_FULL + 8u * b; };   // local copy of the multicast commit
    auto bar_acce = [&](int b) { return aux_addr + AUX_ACC_EMPTY + 8u * b; };  // leader: 16 epilogue warps of the pair
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aux + AUX_TMEM_PTR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    pdl_trigger();

    if (warp == CONTROL_WARP) {
        if (lane == 0) {
            for (int s = 0; s < 4; ++s) {
                mbar_init(bar_a(s), 2 * PRODUCER_WARPS);
                mbar_init(bar_b(s), 1);
                mbar_init(bar_bp(s), 1);
                mbar_init(bar_m(s), 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar_accf(b), 1);
                mbar_init(bar_acce(b), 2 * EPILOGUE_WARPS);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<2>(aux_addr + AUX_TMEM_PTR, 2 * BN2);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // both CTAs' barriers exist before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                      // set-up done; now the predecessor's results are needed
    int M = p.M;
    if (AMODE == AMODE_GATHER) M = p.rowptr[p.n_vtx_frame];
    const int ntm = (M + BM - 1) / BM;
    TileMap tm;
    tm.ntn = tp.ntn; tm.ntm = (ntm + 1) / 2; tm.total = tp.ntn * tm.ntm * tp.frames;
    tm.first = blockIdx.x >> 1; tm.step = gridDim.x >> 1; tm.mult = 2; tm.rank = (int)rank;
    float a_scale, inv;
    operand_scales<KIND>(p, AMODE == AMODE_GATHER, a_scale, inv);

    if (warp >= CONTROL_WARP) {
        reg_dec<REGS_CONTROL>();
        if (warp == CONTROL_WARP) {
            const bool leader = lane == 0;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob);
            const uint32_t chunk_bytes = 2u * BN2 * 128;           // full hi|lo image of one k-chunk in global memory
            const int my_tiles = tm.my_tiles();
            const int f_ntile = (my_tiles > 0) ? tm.decode(tm.first).n_tile : 0;
            // this CTA's 128 (64) rows of the hi and of the lo image of chunk (n_tile, kc) -> dst
            auto fetch_chunk = [&](uint32_t dst, int n_tile, int kc, uint32_t bar) {
                const uint8_t *src = gB + ((size_t)n_tile * nK + kc) * chunk_bytes + rank * HALF_B;
                bulk_g2s(dst, src, HALF_B, bar);
                bulk_g2s(dst + HALF_B, src + BN2 * 128, HALF_B, bar);
            };
            if (resb) {
                if (my_tiles > 0) {
                    if (leader) {
                        mbar_arrive_expect_tx(bar_b(0), 2 * HALF_B * (uint32_t)nK);
                        for (int kc = 0; kc < nK; ++kc) fetch_chunk(base + b_region + kc * b_stride, f_ntile, kc, bar_b(0));
                    }
                    mbar_wait(bar_b(0), 0);
                    if (rank == 0) mbar_wait<true>(bar_bp(0), 0);          // the peer's half image landed too
                    else if (leader) mbar_arrive_cluster(bar_bp(0), 0);
                }
            }
            // (streaming mode: the weight chunks are fetched by the dedicated warp below, which only follows the ring --
            //  with the fetch in this loop a late chunk delayed the NEXT request too: the issuer cannot reach "stage free,
            //  request chunk i + 2" before it has issued chunk i's MMAs, so the requests trailed the MMAs at L / 2 per stage)
            int s = 0;
            uint32_t ph = 0;
            const uint32_t idesc = make_idesc<KIND>(BN2, 2 * BM);     // M = 256 channels, N = 256 rows over the pair
            Tracer tr{(tp.trace && blockIdx.x == 0 && leader) ? tp.trace + 2048 : nullptr, 0};
            if (!(resb && rank != 0)) {                            // resident mode: the peer's control warp is done
                for (int li = 0; li < my_tiles; ++li) {
                    const int buf = li & 1;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN2);
                    if (rank == 0) {
                        tr(20);
                        mbar_wait<true>(bar_acce(buf), ((li >> 1) & 1) ^ 1);   // both CTAs drained this accumulator
                        tr(21);
                        tc_fence_after();
                    }
                    for (int kc = 0; kc < nK; ++kc) {
                        if (!resb) mbar_wait(bar_b(s), ph);                    // own half chunk landed
                        if (rank == 0) {
                            if (!resb) mbar_wait<true>(bar_bp(s), ph);         // peer's half chunk landed
                            tr(23);
                            mbar_wait<true>(bar_a(s), ph);                     // A images of both CTAs written
                            tr(22);
                            tc_fence_after();
                            const uint32_t a_hi = base + s * a_stride;
                            const uint32_t b_hi = base + b_region + (resb ? kc : s) * b_stride;
                            const uint32_t lah = desc_lo(a_hi), lal = desc_lo(a_hi + A_HALF_BYTES);
                            const uint32_t lbh = desc_lo(b_hi), lbl = desc_lo(b_hi + HALF_B);
                            if (leader) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {                  // weights = A operand, activations = B
                                    umma<KIND, 2>(tmem_d, desc64(lbh + 2 * k), desc64(lal + 2 * k), idesc, (kc | k) != 0);
                                    umma<KIND, 2>(tmem_d, desc64(lbl + 2 * k), desc64(lah + 2 * k), idesc, 1);
                                    umma<KIND, 2>(tmem_d, desc64(lbh + 2 * k), desc64(lah + 2 * k), idesc, 1);
                                }
                                umma_commit<2>(bar_m(s));                      // both CTAs: stage s free when retired
                            }
                            tr(24);
                        } else if (leader) {
                            mbar_arrive_cluster(bar_bp(s), 0);                 // tell the leader CTA
                        }
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                    if (rank == 0 && leader) umma_commit<2>(bar_accf(buf));    // both CTAs: accumulator complete
                }
            }
        } else if (warp == CONTROL_WARP + 1 && !resb) {
            // ================= weight-chunk fetcher (streaming mode) =================
            // Walks the same (tile, k-chunk) sequence as the MMA issuer and requests chunk j into ring slot j % S as soon as
            // the MMAs that read the slot one ring turn ago have retired (multicast commit on bar_m) -- never later.
            const bool leader = lane == 0;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob);
            const uint32_t chunk_bytes = 2u * BN2 * 128;
            const int my_tiles = tm.my_tiles();
            int s = 0;
            uint32_t free_ph = 1;                                  // passes on the fresh barriers of the first ring turn
            for (int li = 0; li < my_tiles; ++li) {
                const int n_tile = tm.decode(tm.first + li * tm.step).n_tile;
                for (int kc = 0; kc < nK; ++kc) {
                    mbar_wait(bar_m(s), free_ph);
                    if (leader) {
                        const uint32_t dst = base + b_region + s * b_stride;
                        const uint8_t *src = gB + ((size_t)n_tile * nK + kc) * chunk_bytes + rank * HALF_B;
                        mbar_arrive_expect_tx(bar_b(s), 2 * HALF_B);
                        bulk_g2s(dst, src, HALF_B, bar_b(s));
                        bulk_g2s(dst + HALF_B, src + BN2 * 128, HALF_B, bar_b(s));
                    }
                    if (++s == S) { s = 0; free_ph ^= 1; }
                }
            }
        }
    } else if (warp < PRODUCER_WARPS) {
        reg_inc<RoleCfg<AMODE>::REGS_PRODUCER>();
        auto arrive = [&](int s) {
            if (rank == 0) mbar_arrive(bar_a(s));
            else mbar_arrive_cluster(bar_a(s), 0);
        };
        if constexpr (AMODE == AMODE_GATHER)             // K = H = 256: 8 units per tile
            producer_gather_role<KIND, 8>(p, a_scale, smem, a_stride, aux_addr, S, M, tm, tid, lane, arrive, tp.trace);
        else
            producer_role<KIND, AMODE>(p, a_scale, smem, a_stride, aux_addr, S, nK, M, tm, tid, lane, arrive, tp.trace);
    } else {
        if constexpr (RoleCfg<AMODE>::REGS_EPILOGUE >= 96) reg_inc<RoleCfg<AMODE>::REGS_EPILOGUE>();
        else reg_dec<RoleCfg<AMODE>::REGS_EPILOGUE>();
        auto release = [&](int b) {
            if (rank == 0) mbar_arrive(bar_acce(b));
            else mbar_arrive_cluster(bar_acce(b), 0);
        };
        if constexpr (EPI == EPI_SEGMAX)
            epilogue_segmax_role<2, BN2>(p, inv, aux_addr, reinterpret_cast<int *>(aux + AUX_BYTES), tmem_base, M, tm, (int)rank,
                                         warp, lane, release, tp.trace);
        else
            epilogue_role<2, BN2, EPI>(p, inv, aux_addr, reinterpret_cast<int *>(aux + AUX_BYTES), tmem_base, M, tm, (int)rank,
                                       warp, lane, release, tp.trace);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // no CTA may exit while its pair can still signal it
    if (warp == CONTROL_WARP) tmem_dealloc<2>(tmem_base, 2 * BN2);
}

}  // namespace tc
}  // namespace morig
