// Weighted mean-shift iteration of the joint-extraction post-process (utils/cluster_utils.py:14-35 of the reference,
// called on the shifted vertices right after the jointnet / masknet forward, evaluate/eval_rigging.py:91):
//
//   K[i][j]   = max(bw^2 - |p_i - p_j|^2, 0) * w_i
//   p'_j      = p_j + 0.3 * ( sum_i K[i][j] p_i / (sum_i K[i][j] + 1e-10) - p_j )
//   diff      = sqrt( sum_j |p'_j - p_j|^2 )
//
// The reference materialises the N x N kernel matrix in fp64 numpy every iteration; here a block owns 32 target
// points j, streams the source points i through shared memory in tiles and keeps the four running sums of each
// (j, slice) in fp64 registers.  All reductions run in a fixed order, so results are deterministic; they differ from
// numpy's BLAS summation order by rounding only (tests: 1e-9 absolute, same iteration count).
#include "common.cuh"

namespace morig {

constexpr int MS_JB = 32;            // target points per block
constexpr int MS_SLICES = 8;         // source-point slices per target point
constexpr int MS_THREADS = MS_JB * MS_SLICES;
constexpr int MS_TILE = 256;         // source points per shared-memory tile

__global__ void __launch_bounds__(MS_THREADS) meanshift_step_kernel(const double *__restrict__ pts,
                                                                    const double *__restrict__ w, double bw2, int N,
                                                                    double *__restrict__ out, double *__restrict__ d2) {
    __shared__ double s_p[MS_TILE][4];                       // x, y, z, weight
    __shared__ double s_acc[MS_SLICES][MS_JB][4];
    const int jl = threadIdx.x % MS_JB, sl = threadIdx.x / MS_JB;
    const int j = blockIdx.x * MS_JB + jl;
    const bool j_ok = j < N;
    const double px = j_ok ? pts[3 * j] : 0.0, py = j_ok ? pts[3 * j + 1] : 0.0, pz = j_ok ? pts[3 * j + 2] : 0.0;
    double ax = 0.0, ay = 0.0, az = 0.0, den = 0.0;
    for (int i0 = 0; i0 < N; i0 += MS_TILE) {
        __syncthreads();
        for (int t = threadIdx.x; t < MS_TILE; t += MS_THREADS) {
            const int i = i0 + t;
            const bool ok = i < N;
            s_p[t][0] = ok ? pts[3 * i] : 0.0;
            s_p[t][1] = ok ? pts[3 * i + 1] : 0.0;
            s_p[t][2] = ok ? pts[3 * i + 2] : 0.0;
            s_p[t][3] = ok ? (w ? w[i] : 1.0) : 0.0;        // padding points carry zero weight
        }
        __syncthreads();
#pragma unroll 4
        for (int t = sl; t < MS_TILE; t += MS_SLICES) {
            const double dx = px - s_p[t][0], dy = py - s_p[t][1], dz = pz - s_p[t][2];
            const double y = (dx * dx + dy * dy) + dz * dz;   // numpy's order over the last axis
            const double k = fmax(bw2 - y, 0.0) * s_p[t][3];
            ax += k * s_p[t][0];
            ay += k * s_p[t][1];
            az += k * s_p[t][2];
            den += k;
        }
    }
    s_acc[sl][jl][0] = ax; s_acc[sl][jl][1] = ay; s_acc[sl][jl][2] = az; s_acc[sl][jl][3] = den;
    __syncthreads();
    if (sl == 0 && j_ok) {
        double sx = 0.0, sy = 0.0, sz = 0.0, sd = 0.0;
        for (int s = 0; s < MS_SLICES; ++s) {                // fixed order
            sx += s_acc[s][jl][0]; sy += s_acc[s][jl][1]; sz += s_acc[s][jl][2]; sd += s_acc[s][jl][3];
        }
        const double r = 1.0 / (sd + 1e-10);
        const double nx = 0.3 * (sx * r - px) + px, ny = 0.3 * (sy * r - py) + py, nz = 0.3 * (sz * r - pz) + pz;
        out[3 * j] = nx; out[3 * j + 1] = ny; out[3 * j + 2] = nz;
        const double ex = nx - px, ey = ny - py, ez = nz - pz;
        d2[j] = (ex * ex + ey * ey) + ez * ez;
    }
}

// sum of n doubles in a fixed order (one block): strided partial sums, then a tree
__global__ void __launch_bounds__(1024) sum_f64_kernel(const double *__restrict__ x, int n, double *out) {
    __shared__ double s[1024];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) a += x[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int off = 512; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

}  // namespace morig

using namespace morig;

extern "C" MORIG_API int morig_meanshift_step(const double *pts, const double *weights, double bandwidth, int32_t N,
                                              double *pts_out, double *d2_scratch, double *diff_sq, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    MORIG_CHECK_ARG(pts && pts_out && d2_scratch && diff_sq && N > 0, "meanshift_step: bad argument");
    MORIG_CHECK_ARG(pts != pts_out, "meanshift_step: pts_out must not alias pts");
    meanshift_step_kernel<<<ceil_div(N, MS_JB), MS_THREADS, 0, stream>>>(pts, weights, bandwidth * bandwidth, N, pts_out,
                                                                         d2_scratch);
    MORIG_LAUNCH_CHECK("meanshift_step_kernel");
    sum_f64_kernel<<<1, 1024, 0, stream>>>(d2_scratch, N, diff_sq);
    MORIG_LAUNCH_CHECK("sum_f64_kernel");
    return 0;
}
