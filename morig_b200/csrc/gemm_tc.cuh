// tcgen05 tile engine (sm_100a): 3xTF32 error-compensated GEMM with fp32 accumulation in TMEM.
//
//   D[128 x BN] (+)= A_hi*B_hi + A_hi*B_lo + A_lo*B_hi        (kind::tf32, cta_group::1, M=128, N=BN, K=8)
//
// fp32 inputs are split x = hi + lo with hi = rna_tf32(x); dropping lo*lo leaves ~2^-21 relative error per
// product, i.e. fp32-class results (the 1e-4 absolute tolerance of the path rules out plain TF32/BF16).
//
//   A operand  produced by 8 warps straight into 128B-swizzled K-major shared memory, 32 k-columns per stage:
//              plain activation rows, or relu(P[tgt[e]] + Q[col[e]]) gathered per CSR slot (fused EdgeConv)
//   B operand  weights pre-split (hi|lo) and pre-swizzled on the host into per-(n-tile, k-chunk) blobs that are
//              byte images of the shared-memory stage; one cp.async.bulk (TMA engine, mbarrier complete_tx) each
//   MMA        one elected thread of warp 8 issues 12 tcgen05.mma per stage and tcgen05.commit's to mbarriers
//   epilogue   the 8 producer warps read the accumulator with tcgen05.ld (32 lanes x 32 columns per warp):
//              bias (+ per-graph bias) -> ReLU -> BatchNorm affine, then store / per-graph column max /
//              segmented max over the CSR target (tile staged in shared memory, column-parallel walk)
#pragma once
#include "gemm_simt.cuh"

namespace morig {
namespace tc {

constexpr int BM = 128;
constexpr int KC = 32;                       // fp32 k-columns per stage = one 128-byte swizzle row
constexpr int A_HALF_BYTES = BM * 128;       // 16 KB: hi (then lo) image of the A stage
constexpr int PRODUCER_THREADS = 256;
constexpr int THREADS = PRODUCER_THREADS + 32;

template <int BN> struct Cfg {
    static constexpr int B_HALF_BYTES = BN * 128;
    static constexpr int STAGE_BYTES = 2 * A_HALF_BYTES + 2 * B_HALF_BYTES;
    static constexpr int STAGES = (BN == 256) ? 2 : (BN == 128 ? 3 : 4);
    static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int CS_LD = BN + 1;                         // epilogue staging tile row stride (floats)
    static constexpr int CS_BYTES = BM * CS_LD * 4;
    static constexpr int MAIN_BYTES = PIPE_BYTES > CS_BYTES ? PIPE_BYTES : CS_BYTES;
    static constexpr int AUX_BYTES = 1024;                       // barriers, tmem pointer, row targets
    static constexpr int SMEM_BYTES = MAIN_BYTES + AUX_BYTES + 1024;   // + slack for 1024B alignment
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
};

template <int BN, int AMODE, int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc_gemm_kernel(const TcP tp) {
    using C = Cfg<BN>;
    const GemmP &p = tp.g;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base = (raw_addr + 1023u) & ~1023u;          // swizzle atoms need 1024 B alignment
    uint8_t *smem = smem_raw + (base - raw_addr);
    uint8_t *aux = smem + C::MAIN_BYTES;
    const uint32_t aux_addr = base + C::MAIN_BYTES;
    // aux: [0,8S) a_full, [64,64+8S) b_full, [128,..) mma_done, 192 acc_full, 200 tmem ptr, 256.. row targets
    auto bar_a = [&](int s) { return aux_addr + 8u * s; };
    auto bar_b = [&](int s) { return aux_addr + 64u + 8u * s; };
    auto bar_m = [&](int s) { return aux_addr + 128u + 8u * s; };
    const uint32_t bar_acc = aux_addr + 192u;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(aux + 200);
    int32_t *s_tgt = reinterpret_cast<int32_t *>(aux + 256);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tile = blockIdx.x;                   // n-tiles fastest: CTAs sharing A rows run together
    const int m0 = blockIdx.y * BM;
    const int n0 = n_tile * BN;
    const int frame = (AMODE == AMODE_GATHER) ? blockIdx.z : 0;
    int M = p.M;
    if (AMODE == AMODE_GATHER) {
        M = p.rowptr[p.n_vtx_frame];
        if (m0 >= M) return;
    }
    const int nK = tp.nK;

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < C::STAGES; ++s) {
                mbar_init(bar_a(s), PRODUCER_THREADS);
                mbar_init(bar_b(s), 1);
                mbar_init(bar_m(s), 1);
            }
            mbar_init(bar_acc, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(aux_addr + 200u, BN);
    }
    if (EPI == EPI_SEGMAX && tid < BM) s_tgt[tid] = (m0 + tid < M) ? p.tgt[m0 + tid] : -1;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ================= control warp: B bulk copies + MMA issue (one elected lane) =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc<BN>();
            const uint32_t b_bytes = 2u * C::B_HALF_BYTES;
            const uint8_t *gB = reinterpret_cast<const uint8_t *>(tp.Bblob) + (size_t)n_tile * nK * b_bytes;
            auto issue_b = [&](int kc) {
                const int s = kc % C::STAGES;
                const uint32_t dst = base + s * C::STAGE_BYTES + 2 * A_HALF_BYTES;
                mbar_arrive_expect_tx(bar_b(s), b_bytes);
                bulk_g2s(dst, gB + (size_t)kc * b_bytes, b_bytes, bar_b(s));
            };
            for (int kc = 0; kc < C::STAGES - 1 && kc < nK; ++kc) issue_b(kc);
            for (int kc = 0; kc < nK; ++kc) {
                const int s = kc % C::STAGES;
                const uint32_t ph = (kc / C::STAGES) & 1;
                mbar_wait(bar_a(s), ph);
                mbar_wait(bar_b(s), ph);
                tc_fence_after();
                const uint32_t a_hi = base + s * C::STAGE_BYTES, a_lo = a_hi + A_HALF_BYTES;
                const uint32_t b_hi = a_hi + 2 * A_HALF_BYTES, b_lo = b_hi + C::B_HALF_BYTES;
#pragma unroll
                for (int k = 0; k < KC / 8; ++k) {
                    const uint32_t ko = k * 32;                  // 8 tf32 = 32 bytes along the swizzled row
                    const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
                    const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
                    umma_tf32(tmem_base, dal, dbh, idesc, (kc | k) != 0);
                    umma_tf32(tmem_base, dah, dbl, idesc, 1);
                    umma_tf32(tmem_base, dah, dbh, idesc, 1);
                }
                umma_commit(bar_m(s));                           // frees stage s when these MMAs retire
                const int nxt = kc + C::STAGES - 1;
                if (nxt < nK) {
                    if (kc >= 1) mbar_wait(bar_m((kc - 1) % C::STAGES), ((kc - 1) / C::STAGES) & 1);
                    issue_b(nxt);
                }
            }
            umma_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ================= producer warps: A stage images =================
        // thread -> 16-byte chunk c of rows (tid>>3) + 32*pass
        const int c = tid & 7;
        const int row0 = tid >> 3;
        const float *src0[4];
        const float *src1[4];
        bool ok[4];
#pragma unroll
        for (int ps = 0; ps < 4; ++ps) {
            const int r = m0 + row0 + 32 * ps;
            ok[ps] = r < M;
            if (AMODE == AMODE_GATHER) {
                int i = 0, j = 0;
                if (ok[ps]) { i = p.tgt[r]; j = p.col[r]; }
                const size_t fb = (size_t)frame * p.n_vtx_frame;
                src0[ps] = p.P + (fb + i) * (size_t)p.ldpq + 4 * c;
                src1[ps] = p.Q + (fb + j) * (size_t)p.ldpq + 4 * c;
            } else {
                src0[ps] = p.A + (size_t)(ok[ps] ? r : 0) * p.lda + 4 * c;
                src1[ps] = nullptr;
            }
        }
        float4 cur[4], nxt[4];
        auto load_chunk = [&](int kc, float4 (&dst)[4]) {
            const int k = kc * KC + 4 * c;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok[ps] && k < p.K) {
                    v = *reinterpret_cast<const float4 *>(src0[ps] + kc * KC);
                    if (AMODE == AMODE_GATHER) {
                        const float4 q = *reinterpret_cast<const float4 *>(src1[ps] + kc * KC);
                        v.x = fmaxf(v.x + q.x, 0.f); v.y = fmaxf(v.y + q.y, 0.f);
                        v.z = fmaxf(v.z + q.z, 0.f); v.w = fmaxf(v.w + q.w, 0.f);
                    }
                }
                dst[ps] = v;
            }
        };
        load_chunk(0, cur);
        for (int kc = 0; kc < nK; ++kc) {
            if (kc + 1 < nK) load_chunk(kc + 1, nxt);
            const int s = kc % C::STAGES;
            if (kc >= C::STAGES) mbar_wait(bar_m(s), ((kc / C::STAGES) - 1) & 1);
            uint8_t *a_hi = smem + s * C::STAGE_BYTES;
            uint8_t *a_lo = a_hi + A_HALF_BYTES;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
                const int row = row0 + 32 * ps;
                const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
                const float4 v = cur[ps];
                float4 h, l;
                h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
                *reinterpret_cast<float4 *>(a_hi + off) = h;
                *reinterpret_cast<float4 *>(a_lo + off) = l;
            }
            fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(bar_a(s));
            if (kc + 1 < nK) {
#pragma unroll
                for (int ps = 0; ps < 4; ++ps) cur[ps] = nxt[ps];
            }
        }

        // ================= epilogue =================
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int q = warp & 3;                      // TMEM lane quarter of this warp
        const int half = warp >> 2;                  // column half handled by this warp
        const int row = q * 32 + lane;
        const int r = m0 + row;
        const bool row_ok = r < M;
        int g = 0;
        if (EPI == EPI_STORE && p.batch && row_ok) g = (r / p.n_vtx) * p.n_graphs + p.batch[r % p.n_vtx];
        float *Cs = reinterpret_cast<float *>(smem);
#pragma unroll 1
        for (int cb = 0; cb < BN / 64; ++cb) {
            const int col0 = half * (BN / 2) + cb * 32;
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + col0 + j;
                const bool n_ok = n < p.N;
                float x = v[j] + ((n_ok && p.bias) ? p.bias[n] : 0.f);
                if (EPI == EPI_STORE && p.rowbias && n_ok && row_ok) x += p.rowbias[(size_t)g * p.ldrb + n];
                if (EPI == EPI_SEGMAX || p.relu) x = fmaxf(x, 0.f);
                x = fmaf(x, (n_ok && p.scale) ? p.scale[n] : 1.f, (n_ok && p.shift) ? p.shift[n] : 0.f);
                v[j] = x;
            }
            if (EPI == EPI_SEGMAX) {
#pragma unroll
                for (int j = 0; j < 32; ++j) Cs[row * C::CS_LD + col0 + j] = v[j];
            } else {
                if (p.C && row_ok) {
                    float *dst = p.C + (size_t)r * p.ldc + n0 + col0;
                    if (p.c_vec && n0 + col0 + 31 < p.N) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + col0 + j < p.N) dst[j] = v[j];
                    }
                }
                if (p.pool) {
                    const int g0 = __shfl_sync(0xffffffffu, g, 0);
                    const bool uniform = __all_sync(0xffffffffu, row_ok && g == g0);
                    if (uniform) {
                        float mine = neg_inf();
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float m = v[j];
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                            if (lane == j) mine = m;
                        }
                        const int n = n0 + col0 + lane;
                        if (n < p.N) atomic_max_f32(p.pool + (size_t)g0 * p.ldpool + n, mine);
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + col0 + j < p.N) atomic_max_f32(p.pool + (size_t)g * p.ldpool + n0 + col0 + j, v[j]);
                    }
                }
            }
        }
        if (EPI == EPI_SEGMAX) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            constexpr int PARTS = PRODUCER_THREADS / BN;
            constexpr int ROWS_PER = BM / PARTS;
            const int cc = tid % BN, part = tid / BN;
            const int n = n0 + cc;
            if (n < p.N) {
                const int ra = part * ROWS_PER, rb = ra + ROWS_PER;
                const size_t fb = (size_t)frame * p.n_vtx_frame;
                int cur_t = -1;
                float m = neg_inf();
                auto flush = [&]() {
                    const int lo = p.rowptr[cur_t], hi = p.rowptr[cur_t + 1];
                    float *dst = p.C + (fb + cur_t) * (size_t)p.ldc + n;
                    if (lo >= m0 + ra && hi <= m0 + rb) *dst = m;
                    else atomic_max_f32(dst, m);
                };
                for (int rr = ra; rr < rb; ++rr) {
                    const int t = s_tgt[rr];
                    if (t < 0) break;
                    if (t != cur_t) {
                        if (cur_t >= 0) flush();
                        cur_t = t;
                        m = neg_inf();
                    }
                    m = fmaxf(m, Cs[rr * C::CS_LD + cc]);
                }
                if (cur_t >= 0) flush();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, BN);
}

}  // namespace tc
}  // namespace morig
