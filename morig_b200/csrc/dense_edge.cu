// Host launchers for the FP32 tile engine (vertex dense layers, fused EdgeConv branch) and the
// narrow-channel EdgeConv branch kernel (H <= 32: warp per target vertex, lanes = channels).
#include <stdlib.h>
#include "gemm_tc.cuh"

namespace morig {

// ---- narrow EdgeConv branch --------------------------------------------------------------------
// One warp per (key-frame, target vertex).  H lanes hold the H output channels; 32/H edges of the
// segment are processed side by side.  h = relu(P[i] + Q[j]) lives one channel per lane and is
// broadcast with shuffles for the H x H second layer, whose column sits in registers.
// No atomics: the warp owns the whole segment.
template <int H>
__global__ void __launch_bounds__(256) edge_small_kernel(const morig_edge_desc d) {
    constexpr int G = 32 / H;                       // edges in flight per warp
    const int lane = threadIdx.x & 31;
    const int c = lane % H, sub = lane / H;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (int64_t)d.N * d.n_frames) return;      // whole warps leave together
    const int f = (int)(w / d.N), i = (int)(w % d.N);
    const size_t fb = (size_t)f * d.N;

    float wcol[H];
#pragma unroll
    for (int k = 0; k < H; ++k) wcol[k] = d.W1[(size_t)k * d.ldw + c];
    const float b1 = d.b1[c], sc = d.scale[c], sh = d.shift[c];

    const float pi = d.PQ[(fb + i) * (size_t)d.ldpq + d.p_off + c];
    const int lo = d.rowptr[i], hi = d.rowptr[i + 1];
    float m = neg_inf();
    for (int e0 = lo; e0 < hi; e0 += G) {
        const int e = e0 + sub;
        const bool act = e < hi;
        const int j = act ? d.col[e] : i;
        const float h0 = fmaxf(pi + d.PQ[(fb + j) * (size_t)d.ldpq + d.q_off + c], 0.f);
        float acc = b1;
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(__shfl_sync(0xffffffffu, h0, k, H), wcol[k], acc);
        const float z = fmaf(fmaxf(acc, 0.f), sc, sh);
        if (act) m = fmaxf(m, z);
    }
#pragma unroll
    for (int off = 16; off >= H; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (sub == 0) {
        for (int r = 0; r < d.out_repeat; ++r)
            d.out[((size_t)(f + r) * d.N + i) * (size_t)d.ldo + d.out_off + c] = m;
    }
    amax_commit(d.out_amax, fabsf(m));
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
    kern<<<grid, GEMM_THREADS, smem, stream>>>(p);
    MORIG_LAUNCH_CHECK(name);
    return 0;
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MORIG_CUDA(cudaLaunchKernelEx(&cfg, kern, tp));
    MORIG_LAUNCH_CHECK(name);
    return 0;
}

template <int KIND, int BN, int AMODE, int EPI>
static int launch_tc(const GemmP &p, const void *blob, int frames, cudaStream_t stream, const char *name) {
    if constexpr (BN >= 128) {
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
    kern<<<grid, tc::THREADS, smem, stream>>>(tp);
    MORIG_LAUNCH_CHECK(name);
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

}  // namespace morig

using namespace morig;

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
    if (d->Wtc && p.a_vec && !tc_disabled()) {
        MORIG_CHECK_ARG(aligned16(d->Wtc), "dense_fwd: Wtc must be 16B aligned");
        MORIG_CHECK_ARG(ceil_div(d->M, 128) <= 65535, "dense_fwd: M=%d too large for one launch", d->M);
        MORIG_CHECK_ARG((uint64_t)d->M * (uint64_t)d->lda < (1ull << 32), "dense_fwd: M*lda exceeds 32-bit element offsets");
        MORIG_CHECK_ARG(d->tc_kind == tc::KIND_TF32 || d->tc_kind == tc::KIND_F16, "dense_fwd: tc_kind=%d", d->tc_kind);
        if (d->tc_kind == tc::KIND_F16) {
            MORIG_CHECK_ARG(d->a_amax && d->tc_w_inv > 0.f, "dense_fwd: the fp16 kind needs a_amax and tc_w_inv");
            p.w_inv = d->tc_w_inv;
        }
        switch (d->tc_bn) {
            case 64:  return launch_tc_kind<64, AMODE_PLAIN, EPI_STORE>(d->tc_kind, p, d->Wtc, 1, stream, "tc_dense<64>");
            case 128: return launch_tc_kind<128, AMODE_PLAIN, EPI_STORE>(d->tc_kind, p, d->Wtc, 1, stream, "tc_dense<128>");
            case 256: return launch_tc_kind<256, AMODE_PLAIN, EPI_STORE>(d->tc_kind, p, d->Wtc, 1, stream, "tc_dense<256>");
            default:  MORIG_CHECK_ARG(false, "dense_fwd: tc_bn=%d unsupported (64,128,256)", d->tc_bn);
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
        const int64_t warps = (int64_t)d->N * d->n_frames;
        const unsigned blocks = (unsigned)ceil_div64(warps * 32, 256);
        if (H == 16) edge_small_kernel<16><<<blocks, 256, 0, stream>>>(*d);
        else edge_small_kernel<32><<<blocks, 256, 0, stream>>>(*d);
        MORIG_LAUNCH_CHECK("edge_small_kernel");
        return 0;
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
        MORIG_CHECK_ARG(d->tc_kind == tc::KIND_TF32 || d->tc_kind == tc::KIND_F16, "edgeconv_fwd: tc_kind=%d", d->tc_kind);
        if (d->tc_kind == tc::KIND_F16) {
            MORIG_CHECK_ARG(d->pq_amax && d->tc_w_inv > 0.f, "edgeconv_fwd: the fp16 kind needs pq_amax and tc_w_inv");
            p.w_inv = d->tc_w_inv;
        }
        const int kind = d->tc_kind;
        if (H == 64)  return launch_tc_kind<64, AMODE_GATHER, EPI_SEGMAX>(kind, p, d->W1tc, d->n_frames, stream, "tc_edge<64>");
        if (H == 128) return launch_tc_kind<128, AMODE_GATHER, EPI_SEGMAX>(kind, p, d->W1tc, d->n_frames, stream, "tc_edge<128>");
        return launch_tc_kind<256, AMODE_GATHER, EPI_SEGMAX>(kind, p, d->W1tc, d->n_frames, stream, "tc_edge<256>");
    }
    if (H == 64) {
        dim3 grid(ceil_div(d->E_max, 128), 1, d->n_frames);
        return launch_gemm<128, 64, AMODE_GATHER, EPI_SEGMAX>(p, grid, stream, "edge<128,64>");
    }
    dim3 grid(ceil_div(d->E_max, 128), H / 128, d->n_frames);
    return launch_gemm<128, 128, AMODE_GATHER, EPI_SEGMAX>(p, grid, stream, "edge<128,128>");
}
