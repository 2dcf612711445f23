/*
 * dvcp_b200.h -- C ABI of the B200-native DeepVCP registration hot path.
 *
 * This is the drop-in boundary: a plain-C shared library (libdvcp_b200.so,
 * built from deepvcp-pointcloud-registration_b200/csrc for sm_100a) that a
 * binding in any host language can load. No torch / ATen / pybind types cross
 * it. The reference (vccheng2001/DeepVCP-Pointcloud-Registration) is pure
 * PyTorch; its "FFI for this path" is the Python call surface of the modules
 * cited next to each entry point below (paths relative to the reference root)
 * plus the third-party `knn_cuda.KNN` call convention. INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller allocates all outputs and workspaces; kernels never allocate;
 *   - inputs are never written;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = launched; < 0 = argument error (DVCP_E_*);
 *     > 0 = a cudaError_t raised by the launch;
 *   - point clouds are addressed with element strides
 *     value(b, n, c) = base[b * bstride + n * pstride + c * cstride]
 *     so that both the reference's channel-major model input [B, C, N]
 *     (pstride 1, cstride N) and its point-major utility input [B, N, 3]
 *     (pstride 3, cstride 1) are read in place;
 *   - B > 1 always means "B independent single-cloud problems" (the reference
 *     itself only runs B = 1).
 */
#ifndef DVCP_B200_H
#define DVCP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVCP_ABI_VERSION 2

#if defined(__GNUC__)
#define DVCP_API __attribute__((visibility("default")))
#else
#define DVCP_API
#endif

#define DVCP_E_ARG        (-1) /* null pointer / non-positive size            */
#define DVCP_E_UNSUPPORTED (-2) /* size outside what the kernels are built for */
#define DVCP_E_WORKSPACE  (-3) /* workspace too small                          */

/* Reference-mode quirks. A set bit REPLICATES what the reference does; a clear bit
 * selects the semantics its code intends. DVCP_QUIRKS_REFERENCE is the drop-in mode. */
#define DVCP_QUIRK_KEYPOINT_VIEW       1  /* Q3  deepVCP.py:46        gather result re-read row-major as [K, C]        */
#define DVCP_QUIRK_PER_FEATURE_WEIGHT  2  /* Q7  get_cat_feat_tgt.py:65,92  distance weight indexed by feature         */
#define DVCP_QUIRK_COST_VOLUME_RESHAPE 4  /* Q4  deepVCP.py:106 + cpg.py:34 (feature, candidate) axes scrambled;
                                             host-side: selects `layout` of dvcp_cpg                                  */
#define DVCP_QUIRK_IGNORE_T_INIT       8  /* Q6  deepVCP.py:86-91     t_init never added to the transformed key-points */
#define DVCP_QUIRK_NO_REFLECTION_FIX  16  /* Q10 deepVCP_loss.py:36-40 R = V U^T may have det = -1                     */
#define DVCP_QUIRK_FPS_ORDER_MISMATCH 32  /* Q5  deepVCP.py:35,46,61; get_cat_feat_tgt.py:85  feature rows (FPS order) addressed with
                                              original-order / key-point-local indices. Wiring above the kernels: the caller
                                              passes the tables / indices of the mode it wants; no entry point reads this bit */
#define DVCP_QUIRKS_REFERENCE         63

typedef void *dvcp_stream_t;

typedef struct {
    const void *base;  /* float* or double* (see dtype) */
    int64_t bstride, pstride, cstride; /* in elements */
} dvcp_cloud_t;

/* Folded parameters of one shared-MLP layer (1x1 conv + eval BatchNorm + ReLU,
 * pointnet2_utils.py:171-172,196-198): y = relu((W x + b) * alpha + beta) with
 * alpha = gamma / sqrt(var + eps), beta = bn_bias - mean * alpha. Row-major
 * W[out][in]. */
typedef struct {
    const float *W, *b, *alpha, *beta;
    int in_ch, out_ch;
} dvcp_mlp_layer_t;

/* Spatial index of a cloud (64 <= N <= 16384; up to 131072 through dvcp_build_index_ws): the points in space-filling-curve (Hilbert) order, cut
 * into buckets of 32 consecutive points with their bounding boxes. Built by
 * dvcp_build_index, or as a by-product of dvcp_fps (which sorts the cloud
 * anyway); consumed by the pruned ball-query / SA / KNN kernels. Pruning never
 * changes results: members are always decided by the exact arithmetic. */
typedef struct {
    float *sorted_pt;    /* [B, cap, 4]  x, y, z, original index (int32 bits) in curve order;   
                            unused slots +inf, +inf, +inf, -1: one 16-byte load per point       */
    float *bucket_box;   /* [B, cap/32, 8] minx,miny,minz,maxx,maxy,maxz,count,0                */
    int cap;             /* dvcp_index_capacity(N)                                              */
} dvcp_cloud_index_t;

DVCP_API int dvcp_abi_version(void);
DVCP_API const char *dvcp_error_string(int code);

/* ---- a1  farthest_point_sample(xyz, npoint)        pointnet2_utils.py:63-84
 * start[B] is the random first index the caller drew (pointnet2_utils.py:75).
 * dtype 0: float32 cloud, 1: float64 cloud. out64 [B,npoint] (int64) and/or
 * out32 [B,npoint] (int32) may be null. index_out (all-null = not wanted) also
 * publishes the spatial index of the cloud (float32, 64 <= N <= 16384 only). */
DVCP_API int dvcp_fps(dvcp_cloud_t xyz, int dtype, int B, int N, int npoint, const int64_t *start,
             int64_t *out64, int32_t *out32, dvcp_cloud_index_t index_out, dvcp_stream_t stream);

/* Same as dvcp_fps for float32 clouds whose spatial index is ALREADY built (dvcp_build_index on
 * the same stream or ordered before this call): the index is consumed, not rewritten, so other
 * kernels may read it concurrently (DeepVCP.forward runs the SA layer beside the sampling).
 * concurrent: how the sampling shares the GPU with other work (a stream of batches); same results in every mode.
 *   0  few large clouds are spread over clusters of 8 CTAs each (shortest kernel: 2.3 ms for 16 clouds of 16384);
 *   1  the same with half-size CTAs;
 *   2  ONE CTA per cloud (3.8 ms for any number of clouds up to the SM count, but under half the SM time of mode 0):
 *      for throughput, with the feature halves of several batches in flight (GraphedRegistration, depth >= 3). */
DVCP_API int dvcp_fps_indexed(dvcp_cloud_t xyz, int B, int N, int npoint, const int64_t *start, int64_t *out64,
                     int32_t *out32, dvcp_cloud_index_t index, int concurrent, dvcp_stream_t stream);

/* Capacity (slots) of the spatial index of an N-point cloud; 0 = N not indexable. */
DVCP_API int dvcp_index_capacity(int N);
DVCP_API int dvcp_build_index(dvcp_cloud_t xyz, int B, int N, dvcp_cloud_index_t index_out, dvcp_stream_t stream);

/* The index of larger clouds (16384 < N <= 131072), built by several CTAs per cloud (csrc/index_big.cu:
 * per-run block sorts + rank merges). dvcp_index_capacity_any covers 64 <= N <= 131072 (equal to
 * dvcp_index_capacity up to 16384); dvcp_build_index_ws builds either kind and needs
 * dvcp_build_index_workspace_bytes(B, N) bytes of workspace (0 up to 16384 points). Consumed by
 * dvcp_knn_indexed (SURVEY 8(d) SWEEP: KNN at N = 32k..128k); the sampling and SA kernels take the
 * single-CTA index only. */
DVCP_API int dvcp_index_capacity_any(int N);
DVCP_API int64_t dvcp_build_index_workspace_bytes(int B, int N);
DVCP_API int dvcp_build_index_ws(dvcp_cloud_t xyz, int B, int N, dvcp_cloud_index_t index_out, void *workspace,
                        int64_t workspace_bytes, dvcp_stream_t stream);

/* Test hook: the plain O(N * npoint) kernel for float32 clouds (the spatially
 * pruned kernel dvcp_fps normally dispatches to must give identical indices). */
DVCP_API int dvcp_fps_plain(dvcp_cloud_t xyz, int B, int N, int npoint, const int64_t *start, int64_t *out64,
                   dvcp_stream_t stream);

/* ---- a2  square_distance(src, dst)                  pointnet2_utils.py:19-40
 * out [B,S,N] float32, expanded form. Utility only: the hot path never
 * materialises it. */
DVCP_API int dvcp_square_distance(dvcp_cloud_t src, dvcp_cloud_t dst, int B, int S, int N, float *out,
                         dvcp_stream_t stream);

/* ---- a3  query_ball_point(radius, nsample, xyz, new_xyz)  pointnet2_utils.py:87-107
 * radius2 = float32(radius**2). out [B,S,nsample] int64; an empty ball yields
 * N in every slot (the reference then raises in index_points). */
DVCP_API int dvcp_ball_query(dvcp_cloud_t xyz, dvcp_cloud_t new_xyz, int B, int N, int S, float radius2,
                    int nsample, int64_t *out, dvcp_stream_t stream);

/* ---- a4  index_points(points, idx)                  pointnet2_utils.py:43-60
 * points [B,N,C] contiguous float32, idx [B,M] int64 -> out [B,M,C]. */
DVCP_API int dvcp_index_points(const float *points, const int64_t *idx, int B, int N, int C, int64_t M,
                      float *out, dvcp_stream_t stream);

/* index_points with int32 indices (the FPS order as the kernels keep it): out[b,m,:] = points[b,idx[b,m],:]. */
DVCP_API int dvcp_index_points_i32(const float *points, const int32_t *idx, int B, int N, int C, int64_t M,
                          float *out, dvcp_stream_t stream);

/* ---- a5+a6  PointNetSetAbstraction.forward (group_all=False), eval mode
 *             pointnet2_utils.py:110-138,176-202
 * Fused: centroid gather, ball query, grouping [xyz - centre, feats], shared
 * MLP, max over the ball. centroid_idx [B,S] int32 are the FPS indices.
 * feats: D extra channels addressed like a cloud (null when D == 0).
 * out_feat [B,S,out_ch_last] float32, out_xyz [B,S,3] float32 (may be null).
 * index (all-null = none): spatial index of xyz; when given, only buckets that
 * can hold a member are visited and overflow_ws ([B*S] bytes) is required. */
DVCP_API int dvcp_sa_layer(dvcp_cloud_t xyz, dvcp_cloud_t feats, int D, const int32_t *centroid_idx, int B,
                  int N, int S, float radius2, int nsample, const dvcp_mlp_layer_t *layers_host,
                  int n_layers, dvcp_cloud_index_t index, unsigned char *overflow_ws, float *out_feat,
                  float *out_xyz, dvcp_stream_t stream);

/* The same layer with EVERY point of the cloud as a centroid, in original point order
 * (out_feat [B,N,out_ch_last]); DeepVCP uses npoint == N, and the features of a point do not
 * depend on its FPS rank, so this runs beside the sampling and the rows are gathered into FPS
 * order afterwards. identity_idx [B,N] int32 = 0..N-1 per cloud (caller-provided constant);
 * index and overflow_ws ([B*N + 32] bytes) are required. */
DVCP_API int dvcp_sa_layer_all(dvcp_cloud_t xyz, dvcp_cloud_t feats, int D, const int32_t *identity_idx, int B, int N,
                      float radius2, int nsample, const dvcp_mlp_layer_t *layers_host, int n_layers,
                      dvcp_cloud_index_t index, unsigned char *overflow_ws, float *out_feat,
                      dvcp_stream_t stream);

/* Linear layer over rows: Y [rows,out] = X [rows,in] W^T + b (W [out,in] row-major; in <= 128, out <= 32).
 * deep_feat_extraction.py:15 (`fc`): only used by the repaired three-layer feature extraction
 * (SURVEY 8f rank 1); the reference never calls it. */
DVCP_API int dvcp_linear_rows(const float *X, int64_t rows, int in, int out, const float *W, const float *b, float *Y,
                     dvcp_stream_t stream);

/* ---- a8  weighting_layer.forward(X, K)              weighting_layer.py:26-33
 * X [B,S,32]; W1[16,32] b1 W2[8,16] b2 W3[1,8] b3; scores [B,S] (softplus
 * output, may be null if only indices are wanted -> then `scores` is still
 * required as workspace); topk [B,K] int64 ordered (score desc, index asc). */
DVCP_API int dvcp_weighting_scores(const float *X, int B, int S, const float *W1, const float *b1,
                          const float *W2, const float *b2, const float *W3, const float *b3,
                          float *scores, dvcp_stream_t stream);
DVCP_API int dvcp_topk(const float *scores, int B, int S, int K, int64_t *topk, dvcp_stream_t stream);

/* ---- a9..a11 + a15(src)  key-point stage            deepVCP.py:44-67,86-91,101
 * src_pts [B,C_in,N] channel-major float32; topk [B,Kp] int64; kp_start [B]
 * int64 (FPS start among the key-points, deepVCP.py:54 -> pointnet2_utils.py:75);
 * src_feat [B,S,32] (FPS order); R_init [B,3,3] float64; t_init float64 (may be null), pair b reads
 * t_init[b * t_bstride + 0..2] (t_bstride 0 = one translation for all pairs, like the reference's [1,3]);
 * dfe = {W1[32,35],b1,W2[32,32],b2,W3[32,32],b3}. quirks: DVCP_QUIRK_KEYPOINT_VIEW (Q3),
 * DVCP_QUIRK_IGNORE_T_INIT (Q6: centres = R_init kp, t_init unused; clear: + t_init).
 * Outputs (any may be null): keypts [B,Kp,C_in], picked [B,Kp,nsample] int64,
 * src_cat [B,Kp,nsample,35], src_dfe [B,Kp,32], centres [B,Kp,3] float64. */
typedef struct {
    const float *W1, *b1, *W2, *b2, *W3, *b3;
} dvcp_dfe_params_t;
DVCP_API int dvcp_keypoint_stage(const float *src_pts, int C_in, int B, int N, const int64_t *topk, int Kp,
                        const int64_t *kp_start, const float *src_feat, int S, const double *R_init,
                        const double *t_init, int64_t t_bstride, float radius2, int nsample,
                        dvcp_dfe_params_t dfe, int quirks, float *keypts,
                        int64_t *picked, float *src_cat, float *src_dfe, double *centres,
                        dvcp_stream_t stream);

/* ---- a12 voxelize(point_clouds, r, s)               voxelize.py:19-83
 * centres [M,3] float64 -> out [M,G^3,3] float32; G from dvcp_grid_size. */
DVCP_API int dvcp_grid_size(double r, double s);
DVCP_API int dvcp_candidates(const double *centres, int64_t M, double r, double s, int G, float *out,
                    dvcp_stream_t stream);

/* ---- a13 knn_cuda.KNN(k, transpose_mode=True)(ref, query)
 *          call sites get_cat_feat_tgt.py:45,52; deepVCP_loss.py:70,72
 * ref addressed as a cloud [B,N,3]; query [B,Q,3] contiguous float32.
 * dist [B,Q,K] float32 (sqrt), idx64 [B,Q,K] and/or idx32 (either may be null).
 * Order: (squared distance, index) ascending. 1 <= K <= 32, K <= N. */
DVCP_API int dvcp_knn(dvcp_cloud_t ref, const float *query, int B, int N, int64_t Q, int K, float *dist,
             int64_t *idx64, int32_t *idx32, dvcp_stream_t stream);

/* Same contract and results as dvcp_knn, with the reference cloud given through
 * its spatial index (exact pruning, see csrc/knn.cu). `chain` consecutive queries
 * are processed by one warp, each seeded by the result of the previous one, so
 * they should be spatially close; within a chain every other run of `zline`
 * queries is walked backwards (candidate lattice: chain = G*G, zline = G walks an
 * x-slab in boustrophedon order). Any chain >= 1, 1 <= zline <= chain is correct. */
DVCP_API int dvcp_knn_indexed(dvcp_cloud_index_t ref_index, const float *query, int B, int N, int64_t Q, int K,
                     int chain, int zline, float *dist, int64_t *idx64, int32_t *idx32,
                     dvcp_stream_t stream);

/* Same contract for GROUPS of nearby queries: queries [g * group, (g + 1) * group) of every batch item form
 * one group (the G^3 candidate lattice of one key-point: group = G^3, zline = G, cell = the lattice step s
 * of voxelize.py). One CTA per group copies the target points around the group's bounding box from the
 * index into shared memory, answers the queries there, and certifies every answer (an uncertified query
 * goes through the index search of dvcp_knn_indexed). Exact for ANY queries; results are bit-identical to
 * dvcp_knn. pool_cap: shared-memory pool capacity in points (0 = default; 64..8192). Clouds of up to
 * 65536 points. stats (nullable): 8 device counters ADDED to, for profiling -- queries certified by the pool,
 * uncertified, list overflows, cold starts, queries of groups without a pool, pool points, points admitted
 * by the bound, points scanned. workspace: dvcp_knn_groups_workspace_bytes(B, Q, group) bytes (the pools, their
 * cell tables, and the list of queries the pools could not certify: a last kernel answers those through the
 * index, one warp per query). */
DVCP_API int64_t dvcp_knn_groups_workspace_bytes(int B, int64_t Q, int group);
DVCP_API int dvcp_knn_groups(dvcp_cloud_index_t index, const float *query, int B, int N, int64_t Q, int K,
                    int group, int zline, float cell, int pool_cap, float *dist, int64_t *idx64,
                    int32_t *idx32, uint64_t *stats, void *workspace, int64_t workspace_bytes,
                    dvcp_stream_t stream);

/* ---- a14+a15 Get_Cat_Feat_Tgt + feat_embedding_layer(src=False), fused
 *          get_cat_feat_tgt.py:53-96, deep_feat_embedding.py:46-60
 * cand [B,Q,3]; tgt_xyz cloud [B,N,3]; tgt_feat [B,N,32]; knn_dist [B,Q,32];
 * knn_idx [B,Q,32] int32 -> out [B,Q,32]. quirks bit 1 = per-feature weight (Q7). */
DVCP_API int dvcp_dfe_tgt_fused(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat,
                       const float *knn_dist, const int32_t *knn_idx, int B, int N, int64_t Q,
                       dvcp_dfe_params_t dfe, int quirks, float *out, dvcp_stream_t stream);

/* Tensor-core form of dvcp_dfe_tgt_fused (tcgen05 / TMEM, 3xTF32). The three
 * un-activated Linear layers are collapsed by the caller into one affine map
 * (Wc = W3 W2 W1, bc = W3 (W2 b1 + b2) + b3, float64 on the host); b_hi / b_lo are
 * the TF32 high and low parts of the 32 x 40 operand image
 *     Bm[n][k] = Wc[n][3 + k] (k < 32), Wc[n][k - 32] (32 <= k < 35), bc[n] (k = 35), 0
 * stored at float offset dvcp_dfe_tc_b_offset(n, k) (UMMA K-major core-matrix
 * layout), dvcp_dfe_tc_b_floats() floats each. Same inputs / output as the fused
 * entry point; results agree with it to FP32 round-off of the collapsed map.
 * feature_major_c: 0 = out [B,Q,32] (candidate, feature); C > 0 (Q % C == 0) = every block of C candidates
 * (one key-point) is written feature-major, out [B,Q/C,32,C] -- the logical [32,C] order of deepVCP.py:106
 * that dvcp_cpg reads with layout 0. */
DVCP_API int dvcp_dfe_tc_b_floats(void);
DVCP_API int dvcp_dfe_tc_b_offset(int n, int k);
DVCP_API int dvcp_dfe_tgt_tc(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat, const float *knn_dist,
                    const int32_t *knn_idx, int B, int N, int64_t Q, const float *b_hi, const float *b_lo,
                    int quirks, int feature_major_c, float *out, dvcp_stream_t stream);
DVCP_API int64_t dvcp_cpg_tc_image_bytes(void);

/* ---- a15 feat_embedding_layer.forward on a materialised input
 *          deep_feat_embedding.py:23-61
 * X [rows,K,35] (dtype 0 float32 / 1 float64, cast like X.float()) -> out [rows,32]. */
DVCP_API int dvcp_dfe_dense(const void *X, int dtype, int64_t rows, int K, dvcp_dfe_params_t dfe, float *out,
                   dvcp_stream_t stream);

/* ---- a16 cpg.forward                                 cpg.py:27-60
 * src_dfe [M,32]; tgt_dfe [M,32*C] holding, per key-point, the LOGICAL
 * row-major order of the reference's [32,C] argument when layout == 0, or the
 * DFE's own [C,32] order when layout == 1 (the kernel then applies the
 * permute of deepVCP.py:106 itself); cand [M,C,3]; conv weights as in the
 * state_dict (cpg.conv{1,2,3}.weight/bias). vcp [M,3]; logits [M,C] may be null.
 * workspace: dvcp_cpg_workspace_bytes(M, G). */
typedef struct {
    const float *w1, *b1, *w2, *b2, *w3, *b3;
} dvcp_cpg_params_t;
DVCP_API int64_t dvcp_cpg_workspace_bytes(int64_t M, int G);
DVCP_API int dvcp_cpg(const float *src_dfe, const float *tgt_dfe, int layout, const float *cand, int64_t M,
             int G, dvcp_cpg_params_t p, float *vcp, float *logits, void *workspace,
             int64_t workspace_bytes, dvcp_stream_t stream);
/* Same, with the kernel family chosen by the caller (parity tests compare the families with each other
 * and with the oracle at every grid size): AUTO = what dvcp_cpg picks (TCZ for layout 0 up to 11^3); FUSED = whole chain in one kernel
 * with the volume in shared memory (G <= 11); LAYERED = one kernel per layer through the workspace. */
#define DVCP_CPG_AUTO    0
#define DVCP_CPG_FUSED   1
#define DVCP_CPG_LAYERED 2
#define DVCP_CPG_TC      3   /* conv1 as an implicit GEMM on tcgen05 (3xTF32, TMEM); layout 0, 2 <= G <= 11 */
#define DVCP_CPG_TCZ     4   /* same, the three z taps of a column as the N dimension (9 operand shifts instead of 27) */
DVCP_API int dvcp_cpg_path(const float *src_dfe, const float *tgt_dfe, int layout, const float *cand, int64_t M,
             int G, dvcp_cpg_params_t p, float *vcp, float *logits, void *workspace,
             int64_t workspace_bytes, int path, dvcp_stream_t stream);

/* ---- float64 clouds -- what the reference's own loaders produce (ModelNet40Dataset.py:38,92: float64 clouds;
 * KITTIDataset.py:84,97: float32 scan, float64 target). torch then evaluates these functions in double and casts
 * to float right before the shared MLP (pointnet2_utils.py:198) / the embedding (deep_feat_embedding.py:29):
 *   square_distance / query_ball_point   pointnet2_utils.py:35-40,100-102 in double (fma-chain dot, Python double
 *                                        radius**2 as the bound)
 *   sa_layer_f64                         sample_and_group + PointNetSetAbstraction with float64 xyz / features
 *                                        (clouds of doubles; relative coordinates formed in double, then .float())
 *   keypoint_stage_f64                   deepVCP.py:44-67,86-91 with a float64 source cloud: key-points, their
 *                                        grouping and Get_Cat_Feat_Src in double; keypts out [B,Kp,C_in] double.
 * dvcp_fps takes dtype 1 for float64 clouds. These entry points walk the whole cloud (exact, not tuned). */
DVCP_API int dvcp_square_distance_f64(dvcp_cloud_t src, dvcp_cloud_t dst, int B, int S, int N, double *out,
                             dvcp_stream_t stream);
DVCP_API int dvcp_ball_query_f64(dvcp_cloud_t xyz, dvcp_cloud_t new_xyz, int B, int N, int S, double radius2,
                        int nsample, int64_t *out, dvcp_stream_t stream);
DVCP_API int dvcp_sa_layer_f64(dvcp_cloud_t xyz, dvcp_cloud_t feats, int D, const int32_t *centroid_idx, int B,
                      int N, int S, double radius2, int nsample, const dvcp_mlp_layer_t *layers,
                      int n_layers, float *out_feat, double *out_xyz, dvcp_stream_t stream);
DVCP_API int dvcp_keypoint_stage_f64(const double *src_pts, int C_in, int B, int N, const int64_t *topk, int Kp,
                            const int64_t *kp_start, const float *src_feat, int S, const double *R_init,
                            const double *t_init, int64_t t_bstride, double radius2, int nsample,
                            dvcp_dfe_params_t dfe, int quirks, double *keypts,
                            int64_t *picked, float *src_cat, float *src_dfe, double *centres,
                            dvcp_stream_t stream);

/* ---- a17 get_rigid_transform(x, y)                   deepVCP_loss.py:13-44
 * x, y [B,3,n] (dtype 0 float32 / 1 float64); R [B,3,3], t [B,3] float64.
 * quirks & DVCP_QUIRK_NO_REFLECTION_FIX: R = V U^T as the reference computes it (Q10); clear: the
 * reflection is corrected (det R = +1). weights [B,n] float64 (null = all 1): per-correspondence
 * weights of the weighted solve (centroids and covariance weighted). */
DVCP_API int dvcp_kabsch(const void *x, const void *y, int dtype, const double *weights, int B, int n, int quirks,
                double *R, double *t, dvcp_stream_t stream);

/* ---- a18 svd_optimization(x, y_pred, R_true, t_true)  deepVCP_loss.py:57-90
 * x, y_pred [B,3,n] float64; R_true [B,3,3], t_true [B,3] float64; keep =
 * int(0.8 n). Outputs R2 [B,3,3], t2 [B,3] float64; R1/t1 (first solve) may be
 * null. n <= 1024. quirks: DVCP_QUIRK_NO_REFLECTION_FIX as in dvcp_kabsch (both solves). inliers (nullable)
 * [B,keep] int64: the kept point indices in the order of the reference's topk(largest=False, sorted=True)
 * (:77), from which the caller forms x1 and y_pred2 (:82-88). */
DVCP_API int dvcp_kabsch_refine(const double *x, const double *y_pred, const double *R_true,
                       const double *t_true, int B, int n, int keep, int quirks, double *R2, double *t2,
                       double *R1, double *t1, int64_t *inliers, dvcp_stream_t stream);

/* train.py:110 -> deepVCP_loss.py:105-107,121: svd_optimization on the forward's own outputs. src_keypts,
 * tgt_vcp [B,n,3] float32 (cast to float64 like .double(), read in place: no permute / cast copies);
 * R_true [B,3,3], t_true [B,3] float64 -> R2 [B,3,3], t2 [B,3] float64. Same arithmetic as
 * dvcp_kabsch_refine on the permuted float64 copies. */
DVCP_API int dvcp_pose_from_forward(const float *src_keypts, const float *tgt_vcp, const double *R_true,
                           const double *t_true, int B, int n, int keep, int quirks, double *R2, double *t2,
                           dvcp_stream_t stream);

/* Layout helper of the forward: out [B,N,4] float32 = (x, y, z, 0) per point of a cloud given in any layout
 * (dvcp_dfe_tgt_tc gathers neighbours with one 16-byte load from it). */
DVCP_API int dvcp_pack_xyz4(dvcp_cloud_t xyz, int B, int N, float *out, dvcp_stream_t stream);

/* ---- data ingest   KITTIDataset.py:11-16,39-46,67-84
 * raw: the B scans concatenated, [sum M_b, 4] float32 (x, y, z, reflectance; 16-byte aligned); scan b is
 * rows scan_offset[b] .. scan_offset[b+1]-1 (scan_offset [B+1] int64, device). idx [B,N] int64: the rows
 * `downsample` keeps, relative to the scan (null = rows 0..N-1). src [B,3,N] float32 (the model's layout);
 * tgt [B,3,N] (may be null) = float32(R_b @ double(src) + t_b) with R [B,9], t [B,3] float64;
 * reflectance [B,N] (may be null). A row index outside its scan yields NaN coordinates. */
DVCP_API int dvcp_ingest_kitti(const float *raw, const int64_t *scan_offset, const int64_t *idx, const double *R,
                      const double *t, int B, int N, float *src, float *tgt, float *reflectance,
                      dvcp_stream_t stream);

/* Training: gradient of dvcp_dfe_tgt_fused. Inputs as the forward's; w_collapsed [32,35] = W3 W2 W1
 * (row-major); grad_out [B,Q,32]. The arg-max neighbour of every (candidate, channel) is recomputed with the forward's
 * arithmetic. ACCUMULATES into grad_w [32,35] (gradient w.r.t. the collapsed map), grad_b [32] and grad_feat [B,N,32]
 * (gradient w.r.t. tgt_feat): the caller zeroes them. Distances, coordinates and candidates receive no gradient
 * (get_cat_feat_tgt.py:45-58: knn_cuda runs under no_grad). */
DVCP_API int dvcp_dfe_tgt_backward(const float *cand, dvcp_cloud_t tgt_xyz, const float *tgt_feat, const float *knn_dist,
                          const int32_t *knn_idx, int B, int N, int64_t Q, dvcp_dfe_params_t dfe,
                          const float *w_collapsed, int quirks, const float *grad_out, float *grad_w, float *grad_b,
                          float *grad_feat, dvcp_stream_t stream);

/* ModelNet40Dataset.py:38-41,62-92: B clouds of M rows (x, y, z, nx, ny, nz) float64 as np.loadtxt returns them
 * (raw [B,M,6]); the first N rows of each -> src [B,6,N] and (tgt non-null) tgt [B,6,N] = (R_b xyz + t_b,
 * R_b normals), channel-major, float64 (out_f64 != 0: what the reference's loader yields) or float32.
 * R [B,9], t [B,3] float64. */
DVCP_API int dvcp_ingest_modelnet(const double *raw, const double *R, const double *t, int B, int M, int N, int out_f64,
                         void *src, void *tgt, dvcp_stream_t stream);

/* Voxel-grid filter (the preprocessing of the paper's KITTI pipeline -- the reference's loader
 * only has the random down-sample of KITTIDataset.py:11-16): one output point per occupied cell of the cubic
 * lattice of edge `cell` anchored at (ox, oy, oz). pts: M rows of `stride` floats, the first `channels` (3 or 4:
 * x, y, z[, reflectance]) are reduced. Cell of a point = floor((p - o) / cell) per axis in float32. mode 0: centroid
 * of the cell's points (float64 sums in ascending point index, rounded to float32); mode 1: the cell's first point.
 * Cells come out in ascending (ix, iy, iz) order: out [capacity, channels], out_count [capacity] (nullable: points
 * per cell), n_out (device int64): occupied cells (may exceed capacity; only the first `capacity` are written).
 * workspace: dvcp_voxel_filter_workspace_bytes(M) bytes, 256-byte aligned. */
DVCP_API int64_t dvcp_voxel_filter_workspace_bytes(int64_t M);
DVCP_API int dvcp_voxel_grid_filter(const float *pts, int stride, int channels, int64_t M, float ox, float oy, float oz,
                           float cell, int mode, void *workspace, int64_t capacity, float *out, int32_t *out_count,
                           int64_t *n_out, dvcp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DVCP_B200_H */
