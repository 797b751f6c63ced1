/*
 * morig_b200 — C ABI of the B200-native (sm_100a) forward path of MoRig's rigging networks.
 *
 * The reference (zhan-xu/MoRig) is pure Python: it has no FFI of its own.  What it binds instead
 * are third-party CUDA wheels (PyG 2.0.4, torch_scatter 2.0.9, cuBLAS via torch 1.11) reached from
 * models/basic_modules.py and models/rignet.py.  Each entry point below replaces the group of those
 * calls named in its comment (paths relative to the reference root).  The only caller is the
 * Python host layer `morig_b200/` (ctypes), which mirrors the reference's nn.Module interface.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous memory owned by the caller (PyTorch);
 *     the library never allocates, frees or retains device memory;
 *   - all floating point is IEEE fp32, all indices int32 except the API edge lists (int64 as
 *     delivered by the reference's datasets, datasets/dataset_rig.py:119-120);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises;
 *   - return value 0 = success, otherwise a cudaError_t or MORIG_E_*; `morig_last_error()` gives
 *     the thread-local message.  No exceptions, no exit(), nothing printed.
 */
#ifndef MORIG_B200_H
#define MORIG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MORIG_ABI_VERSION 7

#if defined(__GNUC__)
#define MORIG_API __attribute__((visibility("default")))
#else
#define MORIG_API
#endif

#define MORIG_E_BADARG   1001   /* shape / alignment / flag combination not supported */
#define MORIG_E_WORKSPACE 1002  /* workspace too small */

MORIG_API int         morig_version(void);
MORIG_API const char *morig_last_error(void);
/* number of SMs of the current device (grid sizing by the host layer) */
MORIG_API int         morig_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Graph preparation.  Replaces remove_self_loops + add_self_loops, which the reference re-runs in
 * every EdgeConvMotion.forward (models/basic_modules.py:188-189; 36x per forward on the same two
 * edge lists), and produces the target-sorted CSR the fused kernels consume.
 *
 *   edge_index  int64 [2, E]   row 0 = source j, row 1 = target i (PyG flow source_to_target)
 *   rowptr      int32 [N+1]    out: segment start of every target vertex
 *   col         int32 [E+N]    out: source vertex of every edge, grouped by target; inside a
 *                              target the original edge order is kept and the self loop is last
 *   tgt         int32 [E+N]    out: target vertex of every CSR slot (expanded rowptr)
 *   The normalised edge count E' = (#edges with j != i) + N is rowptr[N] (device side only).
 *   ws          scratch, at least morig_graph_prep_workspace(E, N) bytes
 * ------------------------------------------------------------------------------------------- */
MORIG_API size_t morig_graph_prep_workspace(int64_t E, int32_t N);
MORIG_API int    morig_graph_prep(const int64_t *edge_index, int64_t E, int32_t N,
                        int32_t *rowptr, int32_t *col, int32_t *tgt,
                        void *ws, size_t ws_bytes, void *stream);

/* Brute-force k-nearest-neighbour graph (fp32 squared distance ((a-b)^2).sum(), self excluded, ties
 * to the lower index) inside each graph of a batch; the synthetic stand-in for the reference's
 * offline geodesic-ball builder (data_proc/common_ops.py:214-226).  gptr int32 [B+1] = vertex range
 * of each graph.  Output edge_index int64 [2, k*N]: row 0 = i, row 1 = neighbour. */
MORIG_API int morig_knn_graph(const float *pos, const int32_t *gptr, int32_t n_graphs, int32_t N, int32_t k,
                    int64_t *edge_index, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Dense row-wise layer with fused epilogue ("vertex MLP").  Replaces torch.cat + Linear + ReLU +
 * BatchNorm1d chains on N rows (MLP, models/basic_modules.py:31-36; GCUMotion.mlp :217-218;
 * GCNRig.mlp_glb / mlp_transform, models/rignet.py:62-66), the per-graph scatter_max +
 * repeat_interleave (models/rignet.py:63-64) and, with layer-0 factorisation, the first Linear of
 * the edge MLPs (models/basic_modules.py:193-194) evaluated once per vertex instead of per edge.
 *
 *   out[r, n] = post( sum_k A[r, k] * W[k, n] + bias[n] + rowbias[group(r), n] )
 *   post(v)   = (relu ? max(v, 0) : v) * scale[n] + shift[n]        (scale/shift optional)
 *   group(r)  = (r / n_vtx) * n_graphs + batch[r % n_vtx]           (key-frame block, graph)
 *   pool[group(r), n] = max over rows of out[r, n]                  (optional; ordered atomics,
 *                        pool must be pre-filled with -inf)
 * ------------------------------------------------------------------------------------------- */
typedef struct morig_dense_desc {
    const float   *A;        int32_t lda;     /* [M, K] activations, row stride lda            */
    const float   *W;        int32_t ldw;     /* [K, ldw] packed weights (transposed Linear)   */
    const float   *bias;                       /* [N] or NULL                                   */
    const float   *scale;                      /* [N] or NULL  (eval BatchNorm scale)           */
    const float   *shift;                      /* [N] or NULL  (eval BatchNorm shift)           */
    const float   *rowbias;  int32_t ldrb;    /* [G, ldrb] or NULL; needs batch                */
    const int32_t *batch;                      /* [n_vtx] graph id per vertex, sorted; or NULL  */
    int32_t        n_vtx, n_graphs;
    float         *C;        int32_t ldc;     /* [M, N] output or NULL (pool only)             */
    float         *pool;     int32_t ldpool;  /* [G, N] or NULL                                */
    int32_t        M, N, K;
    int32_t        relu;
    /* optional tensor-core operand: W pre-split into hi|lo halves and pre-swizzled into the shared-memory
     * image of each (n-tile of tc_bn rows, k-chunk) -- see morig_b200/packing.py:pack_tc_blob.
     *   tc_kind 0: tf32 halves, k-chunks of 32;   tc_kind 1: fp16 halves of W * 2^j, k-chunks of 64,
     *   tc_w_inv = 2^-j, and a_amax must point to a device float >= max|A| (see morig_absmax_f32).
     * When non-NULL (and A is 16-byte aligned with lda % 4 == 0, K % 4 == 0) the layer runs on the tcgen05
     * split-precision engine (3 MMAs per product, fp32-class results), otherwise on the fp32 CUDA-core
     * engine using W. */
    const void    *Wtc;      int32_t tc_bn;   /* tc_bn in {128, 256} (channels per n-tile)      */
    int32_t        tc_kind;  float   tc_w_inv;
    const float   *a_amax;                     /* device scalar, required for tc_kind 1          */
    /* optional device scalar: atomically raised to max |C[r, n]| over everything this call stores
     * (bit pattern of a non-negative float; zero it before the first producer of a buffer)          */
    float         *c_amax;
    /* optional: device scalar holding tc_w_inv (weight image packed on the device by morig_pack_tc_f16); overrides the
     * by-value field when non-NULL */
    const float   *tc_w_inv_dev;
} morig_dense_desc;

MORIG_API int morig_dense_fwd(const morig_dense_desc *d, void *stream);

/* fp16-split tensor-core image of a Linear weight, built on the DEVICE (training: the weights change every step).
 * Logical operand w_out_in [N, K]: element (n, k) = src[n * lds + k], or src[k * lds + n] when `transposed` (the input-
 * gradient GEMM dX = dY W uses W^T).  Same byte layout as packing.pack_tc_blob(kind = F16): blob[n_tile][k_chunk of 64]
 * [hi | lo][bn rows][16-byte chunk c ^ (row % 8)][8 fp16], values scaled by 2^j with max|W| 2^j in [2^14, 2^15);
 * *w_inv_dev = 2^-j.  blob bytes = ceil(N / bn) * ceil(K / 64) * 2 * bn * 128;  amax_scratch: one float. */
MORIG_API size_t morig_pack_tc_f16_bytes(int32_t N, int32_t K, int32_t bn);
MORIG_API int    morig_pack_tc_f16(const float *src, int32_t lds, int32_t N, int32_t K, int32_t transposed, int32_t bn, void *blob,
                                   float *w_inv_dev, float *amax_scratch, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused EdgeConv branch.  Replaces, for one of the two MLPs of EdgeConvMotion.message and for one
 * edge set: the x_i/x_j index_select gathers, cat[x_i, x_j - x_i], Linear/ReLU/BN x2 on E rows and
 * the scatter-max of PyG's aggregate (models/basic_modules.py:190-199, MLP :31-36).
 *
 *   PQ      [R, ldpq]  per-vertex halves of the factorised first Linear:
 *           P[v] = (W0a - W0b) x_v + b0   at columns [p_off, p_off+H)
 *           Q[v] =  W0b x_v               at columns [q_off, q_off+H)
 *   edge e = (j -> i):   h = relu(P[i] + Q[j])                     (BN #1 folded into W1/b1)
 *                        z = relu(h W1 + b1) * scale + shift       (BN #2 applied before max)
 *   out[i, out_off + c] = max over in-edges of z[c]
 *   R = n_frames * N rows: key-frame f uses rows [f*N, (f+1)*N) of PQ and out with the same CSR.
 *   out_repeat > 1 (pos branch shared by all key-frames) stores the result to that many row
 *   blocks of `out` while reading PQ rows [0, N).
 *   out must be pre-filled with -inf where segments may straddle 128-edge tiles (use
 *   morig_fill_f32); tiles combine with ordered atomics, so results are deterministic.
 *   (out_amax bounds the per-edge values z, hence also the stored maxima.)
 * ------------------------------------------------------------------------------------------- */
typedef struct morig_edge_desc {
    const float   *PQ;     int32_t ldpq, p_off, q_off;
    const int32_t *rowptr; const int32_t *col; const int32_t *tgt;
    int32_t        N;            /* vertices per key-frame block                    */
    int32_t        E_max;        /* allocated CSR slots (E + N); E' read on device  */
    int32_t        n_frames;     /* key-frame blocks sharing the CSR                */
    int32_t        out_repeat;   /* >=1                                             */
    const float   *W1;     int32_t ldw;      /* [H, ldw] packed second Linear       */
    const float   *b1, *scale, *shift;        /* [H]                                 */
    float         *out;    int32_t ldo, out_off;
    int32_t        H;
    const void    *W1tc;                      /* optional tcgen05 image of W1 (n-tile = H <= 256) */
    int32_t        tc_kind;  float   tc_w_inv;/* as in morig_dense_desc                           */
    const float   *pq_amax;                   /* device scalar >= max|PQ|: required for tc_kind 1 and for H <= 32 */
    float         *out_amax;                  /* optional, raised to max |edge value| before the max */
} morig_edge_desc;

MORIG_API int morig_edgeconv_fwd(const morig_edge_desc *d, void *stream);

/* `count` (1..4) narrow branches (H in {16, 32}, the same for all) on the same graph and key-frame count in ONE launch:
 * the pos branches of the three GCUs of a GCNRig (models/rignet.py:59-61 -> basic_modules.py:215-216) are independent
 * of the GCU chain -- they only read pos -- and each is too small to fill the machine. */
MORIG_API int morig_edgeconv_fwd_batch(const morig_edge_desc *d, int32_t count, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Per-vertex cls-token attention over key-frames.  Replaces TemporalAttn.forward up to and
 * including w_o (models/rignet.py:36-44); only the cls query row is evaluated because only
 * res[:, 0, :] is used (:45).  Constants are parameter-only products prepared by the host layer:
 *   u      [heads, C]      = Wk_h^T (Wq_h cls) / sqrt(d)       logit of token x is u_h . x
 *   l0     [heads]         = (Wq_h cls) . (Wk_h cls) / sqrt(d) logit of the cls key
 *   Mv     [heads, D, C]   = Wo[:, h] Wv_h                     value+output projection
 *   c0     [heads, D]      = Wo[:, h] (Wv_h cls)
 *   x      [N, T, C] -> out [N, D]
 * ------------------------------------------------------------------------------------------- */
MORIG_API int morig_temporal_attn_fwd(const float *x, int32_t N, int32_t T, int32_t C, int32_t heads, int32_t D,
                            const float *u, const float *l0, const float *Mv, const float *c0,
                            float *out, int32_t ldo, float *out_amax, void *stream);

/* x[r, 0:C] /= max(||x[r, 0:C]||_2, 1e-12)  — F.normalize(dim=1), models/rignet.py:87,98,120,131,199,203.
 * If dst2 != NULL the normalised row r = f*N + v is also written to dst2[v, f, 0:C]
 * (the torch.stack of key-frames, models/rignet.py:89). */
MORIG_API int morig_row_normalize(float *x, int32_t ldx, int32_t R, int32_t C,
                        float *dst2, int32_t N, int32_t n_frames, void *stream);

/* Key-frame reductions of the non-attention aggregators (models/rignet.py:92-95):
 * mode 0 = mean, 1 = max over T of x[N, T, C] -> out[N, C] (row stride ldo). */
MORIG_API int morig_frame_reduce(const float *x, int32_t N, int32_t T, int32_t C, int32_t mode,
                       float *out, int32_t ldo, void *stream);

/* Scatter source columns into a strided destination, replacing the torch.cat / column slicing on the
 * path (models/rignet.py:65,86,159-173).  For r in [0, n_frames*N), c in [0, C):
 *   dst[r, dst_off + c] = src[(r % N), src_off + (r / N) * frame_stride + cols ? cols[c] : c]   */
MORIG_API int morig_gather_cols(const float *src, int32_t lds, int32_t src_off, int32_t frame_stride,
                      const int32_t *cols, int32_t C, int32_t N, int32_t n_frames,
                      float *dst, int32_t ldd, int32_t dst_off, float *dst_amax, void *stream);

MORIG_API int morig_fill_f32(float *dst, int64_t n, float value, void *stream);
/* up to 8 (16-byte aligned) buffers in one launch: the -inf initialisation of all EdgeConv outputs of a GCNRig */
MORIG_API int morig_fill_many_f32(float *const *dst, const int64_t *n, int32_t count, float value, void *stream);
/* Start values only where the fused EdgeConv kernels merge with the atomic max: for every entry k (<= 8), every vertex v
 * whose CSR segment [rowptr[k][v], rowptr[k][v+1]) straddles a multiple of 32 slots and every key-frame f, writes `value`
 * to out[k][(f*N + v)*ld[k] + col0[k] .. + ncols[k]).  All other vertices are written by one plain store of the kernel
 * that owns their segment.  Replaces the reference's zero-initialised scatter output (torch_scatter `scatter(...,
 * reduce='max')` behind models/basic_modules.py:180-181) for the EdgeConv outputs of one GCNRig in a single launch. */
MORIG_API int morig_fill_cut_f32(const int32_t *const *rowptr, float *const *out, const int32_t *ld, const int32_t *col0,
                                 const int32_t *ncols, int32_t count, int32_t N, int32_t frames, float value, void *stream);

/* *amax = max(*amax, max |x[r, c]|) over r < R, c < C (row stride ldx): the operand range the fp16-split
 * tensor-core layers need for inputs that were not produced by this library (out_amax / dst_amax /
 * c_amax arguments above keep it up to date for everything that was). */
MORIG_API int morig_absmax_f32(const float *x, int32_t ldx, int32_t R, int32_t C, float *amax, void *stream);

/* ---------------------------------------------------------------------------------------------
 * One weighted mean-shift iteration of the joint-extraction post-process (SURVEY.md section 8(f) #4:
 * utils/cluster_utils.py:14-35 `meanshift_cluster`, called on the shifted vertices right after the
 * jointnet / masknet forward, evaluate/eval_rigging.py:91).  fp64 like the reference's numpy code.
 *   pts, pts_out [N,3] (must not alias), weights [N] or NULL, d2_scratch [N],
 *   diff_sq: device scalar receiving sum_j |p'_j - p_j|^2 (the host loop stops when its sqrt <= 1e-3).
 * ------------------------------------------------------------------------------------------- */
MORIG_API int morig_meanshift_step(const double *pts, const double *weights, double bandwidth, int32_t N,
                                   double *pts_out, double *d2_scratch, double *diff_sq, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Surface-geodesic graph build from surface samples + normals:
 * data_proc/common_ops.py:182-208 (`calc_surface_geodesic` after open3d's sampling) and :214-226
 * (`get_geo_edges`).  fp64; the all-pairs distances equal scipy's Dijkstra bit for bit.
 *   pts, normals [S,3]; verts [V,3]; out [V,V] = surface_geodesic; ws from ..._workspace(S, V).
 *   geo_ball_edges: edges [V, max_nn, 2] int64 rows (i, j) (first degree[i] rows of vertex i valid):
 *   ball members in ascending index order when they fit, else the max_nn nearest.
 * ------------------------------------------------------------------------------------------- */
MORIG_API size_t morig_surface_geodesic_workspace(int32_t S, int32_t V);
MORIG_API int    morig_surface_geodesic(const double *pts, const double *normals, int32_t S, const double *verts,
                                        int32_t V, double *out, void *ws, size_t ws_bytes, void *stream);
MORIG_API int    morig_geo_ball_edges(const double *geodesic, int32_t V, double radius, int32_t max_nn,
                                      int64_t *edges, int32_t *degree, void *stream);

/* Topological edges from the triangle list: data_proc/common_ops.py:15-32 (`get_tpl_edges`).
 *   faces [F,3] int64 -> edges [<= 6F, 2] int64 rows (v, n): per vertex its distinct face neighbours, ascending
 *   (the reference lists them in python-set order; the edge set is identical).  *count = rows written. */
MORIG_API size_t morig_tpl_edges_workspace(int64_t F);
MORIG_API int    morig_tpl_edges(const int64_t *faces, int64_t F, int64_t *edges, int64_t *count, void *ws,
                                 size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Rest of the joint-extraction post-process and the losses next to it.
 *
 * morig_nms_meanshift: utils/cluster_utils.py:38-63 (`nms_meanshift`, run after the mean-shift at
 *   evaluate/eval_rigging.py:94).  pts [N,3], attn [N] fp64; keep [N] receives 1 for the surviving modes (the caller
 *   compacts: pts[keep]).  Visiting order: decreasing neighbour count, equal counts from the higher index down
 *   (= np.argsort(kind="stable")[::-1]; the reference's default argsort leaves ties unspecified).
 * morig_nn_dist_*: nearest-neighbour distance (and index, first minimum) of every row of A [N,D] among B [M,D], D <= 4:
 *   both halves of `chamfer_distance_with_average` (models/customized_losses.py:231-251, fp32) and of `chamfer_dist`
 *   (utils/mst_utils.py:316-321, fp64).  morig_chamfer_bwd_f32: gradient with respect to A of
 *   g1 * sum_i d1[i] + g2 * sum_j d2[j]  (d1 / n1 = A against B, d2 / n2 = B against A).
 * morig_info_nce_*: loss[r] = logsumexp_m(A[r] . K[m] / tau) - A[r] . K[label[r]] / tau  (cross entropy of the similarity
 *   rows: `infoNCE`, models/customized_losses.py:107-135, per direction; `multi_pos_infoNCE` :137-158 after its
 *   sampling); lse [R] is saved for the backward, which writes dA and accumulates dK (optional).  With sel [R, S] != NULL
 *   row r only sees its candidate keys K[sel[r, s]] and label[r] is a position in that list (multi_pos_infoNCE: one
 *   positive + 200 negatives per anchor).
 * ------------------------------------------------------------------------------------------- */
MORIG_API size_t morig_nms_meanshift_workspace(int32_t N);
MORIG_API int    morig_nms_meanshift(const double *pts, const double *attn, int32_t N, double bandwidth, double thrd_density,
                                     double thrd_attn, uint8_t *keep, void *ws, size_t ws_bytes, void *stream);
MORIG_API int    morig_nn_dist_f32(const float *A, int32_t N, const float *B, int32_t M, int32_t D, float *dist, int32_t *arg,
                                   void *stream);
MORIG_API int    morig_nn_dist_f64(const double *A, int32_t N, const double *B, int32_t M, int32_t D, double *dist, int32_t *arg,
                                   void *stream);
MORIG_API int    morig_chamfer_bwd_f32(const float *A, int32_t N, const float *B, int32_t M, int32_t D, const float *d1,
                                       const int32_t *n1, const float *d2, const int32_t *n2, float g1, float g2, float *dA,
                                       void *stream);
MORIG_API int    morig_info_nce_fwd(const float *A, int32_t lda, const float *K, int32_t ldk, const int64_t *label,
                                    const int64_t *sel, int32_t S, int32_t R, int32_t M, int32_t C, float tau, float *loss,
                                    float *lse, void *stream);
MORIG_API int    morig_info_nce_bwd(const float *A, int32_t lda, const float *K, int32_t ldk, const int64_t *label,
                                    const int64_t *sel, int32_t S, const float *lse, const float *g, int32_t R, int32_t M,
                                    int32_t C, float tau, float *dA, int32_t ldda, float *dK, int32_t lddk, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Point-cloud primitives of the upstream flow producer (CorrNet / DeformNet,
 * models/corrnet.py:10-77, models/deformnet.py:34-99, PointNet++ modules models/basic_modules.py:66-138) and of the
 * surface-sampling front-end of the geodesic graph build (section 8(f) #2, data_proc/common_ops.py:175-181).
 * They replace torch_cluster.fps / radius / knn and PyG's knn_interpolate / PointConv (semantics restated in
 * oracle/pointops_port.py).  Segments: points of sample b are rows ptr[b] .. ptr[b+1] (sorted `batch` vectors).
 *
 * morig_fps            farthest point sampling per segment: out[out_ptr[b] ..] = global indices in selection order, first
 *                      pick = ptr[b] + start[b] (start == NULL: 0), ties -> lower index; segments <= 16384 points
 * morig_ball_query     nbr [M, K] = the first K points (index order) of the centre's segment with |x - y|^2 < radius^2,
 *                      count [M] = number of valid entries (torch_cluster.radius with max_num_neighbors = K)
 * morig_knn_topk       nbr [M, k] (k <= 8) nearest points of the query's segment, nearest first; metric 0 = squared
 *                      Euclidean distance, 1 = cosine similarity (torch_cluster.knn(cosine=True)); score optional
 * morig_knn_interpolate  out[i] = sum_k w_k f[nbr[i,k]] / sum_k w_k, w = 1 / max(|pos_x - pos_y|^2, 1e-16)
 * morig_sample_surface   S area-weighted uniform samples on a triangle mesh + unit face normals (fp64; splitmix64 hash of
 *                      (seed, sample) -- reproducible on any device); ws: 2 F doubles
 * morig_edge_mlp_layer   Linear -> ReLU -> affine on E edge rows of a bipartite neighbourhood graph (PointConv's local MLP,
 *                      PyG PointNetConv): rows = A [E, K] or relu(P[tgt[e]] + Q[col[e]]) (factorised first Linear on
 *                      cat[x_j, pos_j - pos_i]); result stored per edge (C) or max-reduced per target into `out`
 *                      (pre-filled with -inf), rowptr [n_targets + 1] / tgt [E] describing the target-sorted rows
 * ------------------------------------------------------------------------------------------- */
MORIG_API int morig_fps(const float *pos, const int32_t *ptr, const int32_t *out_ptr, const int32_t *start, int32_t B,
                        int32_t max_segment, int32_t *out, void *stream);
MORIG_API int morig_ball_query(const float *x, const int32_t *x_ptr, const float *y, const int32_t *y_batch, int32_t M, float radius,
                               int32_t K, int32_t *nbr, int32_t *count, void *stream);
MORIG_API int morig_knn_topk(const float *x, int32_t ldx, const int32_t *x_ptr, const float *y, int32_t ldy, const int32_t *y_batch,
                             int32_t M, int32_t D, int32_t k, int32_t metric, int32_t *nbr, float *score, void *stream);
MORIG_API int morig_knn_interpolate(const float *f, int32_t ldf, const float *pos_x, const float *pos_y, const int32_t *nbr, int32_t M,
                                    int32_t k, int32_t C, float *out, int32_t ldo, void *stream);
MORIG_API int morig_sample_surface(const double *verts, const int64_t *faces, int32_t F, int32_t S, uint64_t seed, double *pts,
                                   double *normals, double *ws, void *stream);

typedef struct morig_edge_layer_desc {
    const float   *A;      int32_t lda;              /* plain edge rows [E, K], or NULL                    */
    const float   *P, *Q;  int32_t ldpq;             /* gathered rows relu(P[tgt] + Q[col]), or NULL       */
    const int32_t *rowptr; const int32_t *col; const int32_t *tgt;
    int32_t        n_targets, E;
    const float   *W;      int32_t ldw;              /* [K, ldw] packed (transposed) Linear weight         */
    const float   *bias, *scale, *shift;             /* [N] (scale / shift optional)                       */
    float         *C;      int32_t ldc;              /* per-edge output [E, N], or NULL                    */
    float         *out;    int32_t ldo;              /* per-target max [n_targets, N], or NULL             */
    int32_t        N, K;
} morig_edge_layer_desc;

MORIG_API int morig_edge_mlp_layer(const morig_edge_layer_desc *d, void *stream);

/* =============================================================================================
 * Training path.  The reference trains these networks with torch autograd
 * (training/train_rig.py:136-195, training/train_skin.py:139-183): train-mode BatchNorm1d inside every MLP block
 * (models/basic_modules.py:33), scatter-max whose gradient goes to the FIRST maximal edge (torch_scatter), cuBLAS
 * GEMMs for the Linear gradients.  The entry points below are what the autograd functions of
 * morig_b200/autograd_ops.py bind; forward GEMMs and the input-gradient GEMMs (dX = dY W) go through
 * morig_dense_fwd.  fp32 data, fp64 accumulation for every reduction over rows, fixed summation order except the
 * scatter-add of morig_edge_gather_relu_bwd's dQ (fp32 atomics).  Matrices are row-major with explicit row strides.
 * ============================================================================================= */

/* dst [cols, ldd] = src [rows, cols]^T, columns rows..ldd-1 zeroed: a Linear weight [out, in] -> the packed [K, ldw]
 * operand of morig_dense_fwd (weights change every optimisation step) */
MORIG_API int morig_transpose_pad_f32(const float *src, int32_t rows, int32_t cols, int32_t lds, float *dst, int32_t ldd,
                                      void *stream);

/* Linear weight / bias gradient:  dW [N, K] (+)= dY^T (X * x_scale + x_shift),  dbias [N] (+)= column sums of dY
 * (x_scale / x_shift [K] optional; dbias optional; accumulate != 0 adds to the existing contents).
 * ws: morig_wgrad_workspace(M, N, K) bytes. */
MORIG_API size_t morig_wgrad_workspace(int32_t M, int32_t N, int32_t K);
MORIG_API int    morig_wgrad_f32(const float *dY, int32_t lddy, const float *X, int32_t ldx, int32_t M, int32_t N, int32_t K,
                                 const float *x_scale, const float *x_shift, float *dW, int32_t lddw, float *dbias,
                                 int32_t accumulate, void *ws, size_t ws_bytes, void *stream);

/* train-mode BatchNorm1d over the R rows of x [R, C] (torch semantics, models/basic_modules.py:33): batch mean and
 * biased variance normalise, running_mean / running_var (optional) move by `momentum` towards the batch mean / unbiased
 * variance.  Writes mean, invstd, scale = gamma * invstd, shift = beta - mean * scale [C] and, if y != NULL,
 * y = x * scale + shift.  y_amax (optional): device scalar, zeroed by the caller, that receives max |y| (ordered-int atomic max
 * per thread block) -- the operand range the fp16-split tensor-core GEMM consuming y needs; the same for dz_amax of
 * morig_bn_relu_bwd and h_amax of morig_edge_gather_relu.  ws: morig_colstats_workspace(R, C) bytes. */
MORIG_API size_t morig_colstats_workspace(int32_t R, int32_t C);
MORIG_API int    morig_bn_train_fwd(const float *x, int32_t ldx, int32_t R, int32_t C, const float *gamma, const float *beta,
                                    float eps, float momentum, float *running_mean, float *running_var, float *mean,
                                    float *invstd, float *scale, float *shift, float *y, int32_t ldy, float *y_amax, void *ws,
                                    size_t ws_bytes, void *stream);

/* y = x * scale[c] + shift[c] (per column): the apply step of the BatchNorm above on its own; also the 1 / T of the
 * key-frame mean (models/rignet.py:92-93) in the training path */
MORIG_API int    morig_col_affine(const float *x, int32_t ldx, int32_t R, int32_t C, const float *scale, const float *shift, float *y,
                                  int32_t ldy, void *stream);

/* backward of [ReLU ->] BatchNorm(train) given dy and the saved BatchNorm input x (= ReLU output):
 *   dgamma = sum dy * xhat, dbeta = sum dy, dz = [x > 0 or !relu] * gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat))
 * coef: scratch [3 C].  dz may alias dy. */
MORIG_API int    morig_bn_relu_bwd(const float *dy, int32_t lddy, const float *x, int32_t ldx, int32_t R, int32_t C,
                                   const float *gamma, const float *mean, const float *invstd, int32_t relu, float *dz,
                                   int32_t lddz, float *dgamma, float *dbeta, float *coef, float *dz_amax, void *ws,
                                   size_t ws_bytes, void *stream);
/* dz = [y > 0] * dy */
MORIG_API int    morig_relu_bwd(const float *dy, int32_t lddy, const float *y, int32_t ldy, int32_t R, int32_t C, float *dz,
                                int32_t lddz, void *stream);

/* first edge layer after factorisation, materialised for training: h [E, C] = relu(P[tgt[e]] + Q[col[e]]) over the E
 * valid CSR slots (models/basic_modules.py:193-194); backward: dP[v] = sum over v's in-edges, dQ[u] = sum over u's
 * out-edges of [h > 0] * dh (both overwritten) */
MORIG_API int    morig_edge_gather_relu(const float *P, int32_t ldp, const float *Q, int32_t ldq, const int32_t *tgt,
                                        const int32_t *col, int32_t E, int32_t C, float *h, int32_t ldh, float *h_amax,
                                        void *stream);
MORIG_API int    morig_edge_gather_relu_bwd(const float *dh, int32_t lddh, const float *h, int32_t ldh, const int32_t *rowptr,
                                            const int32_t *col, int32_t N, int32_t E, int32_t C, float *dP, int32_t ldp,
                                            float *dQ, int32_t ldq, void *stream);

/* segmented max over contiguous row segments ptr[s]..ptr[s+1] of y [R, C] with argmax = the FIRST maximal row
 * (torch_scatter.scatter_max; PyG aggr='max', models/basic_modules.py:180-181; models/rignet.py:63,176); empty
 * segment -> 0 / arg -1.  Backward: dy [R, C] overwritten with zeros and dy[arg[s, c], c] = dout[s, c]. */
MORIG_API int    morig_segmax_fwd(const float *y, int32_t ldy, const int32_t *ptr, int32_t S, int32_t C, float *out, int32_t ldo,
                                  int32_t *arg, int32_t lda, void *stream);
MORIG_API int    morig_segmax_bwd(const float *dout, int32_t lddo, const int32_t *arg, int32_t lda, int32_t S, int32_t C,
                                  float *dy, int32_t lddy, int32_t R, void *stream);
/* ptr [S + 1] of a sorted key vector (PyG `batch`) */
MORIG_API int    morig_seg_ptr(const int32_t *keys, int32_t N, int32_t S, int32_t *ptr, void *stream);
/* out[r] = src[idx[r]] (repeat_interleave of the pooled feature, models/rignet.py:64) and its gradient, the per-segment sum */
MORIG_API int    morig_row_gather(const float *src, int32_t lds, const int32_t *idx, int32_t R, int32_t C, float *out,
                                  int32_t ldo, void *stream);
MORIG_API int    morig_seg_sum(const float *x, int32_t ldx, const int32_t *ptr, int32_t S, int32_t C, float *out, int32_t ldo,
                               void *stream);

/* F.normalize(dim=1) out of place, and its gradient */
MORIG_API int    morig_normalize_fwd(const float *x, int32_t ldx, int32_t R, int32_t C, float *y, int32_t ldy, void *stream);
MORIG_API int    morig_normalize_bwd(const float *x, int32_t ldx, const float *dy, int32_t lddy, int32_t R, int32_t C, float *dx,
                                     int32_t lddx, void *stream);

/* cls-query attention over the key-frames with explicit projections (TemporalAttn, models/rignet.py:36-45, training
 * form): q0 / kc / vc [HD] = projected cls token, Kx / Vx [N, T, HD] = projected key-frame tokens, HD = heads * d.
 * out [N, HD]; att [N, heads, T + 1] saved for the backward (slot 0 = cls).  Backward writes dq0 / dkc / dvc [HD]
 * (sums over vertices, fixed order) and dKx / dVx.  ws: morig_attn_cls_bwd_workspace(N, HD) bytes. */
MORIG_API int    morig_attn_cls_fwd(const float *q0, const float *kc, const float *vc, const float *Kx, const float *Vx,
                                    int32_t N, int32_t T, int32_t HD, int32_t d, float *out, float *att, void *stream);
MORIG_API size_t morig_attn_cls_bwd_workspace(int32_t N, int32_t HD);
MORIG_API int    morig_attn_cls_bwd(const float *q0, const float *kc, const float *vc, const float *Kx, const float *Vx,
                                    const float *att, const float *dout, int32_t N, int32_t T, int32_t HD, int32_t d,
                                    float *dq0, float *dkc, float *dvc, float *dKx, float *dVx, void *ws, size_t ws_bytes,
                                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MORIG_B200_H */
