/*
 * pgs_b200.h -- C ABI of the B200-native hot path (sm_100a) that replaces the three
 * un-vendored native dependencies of prs-eth/PanopticSegForLargeScalePointCloud:
 *
 *   MinkowskiEngine          coordinate hash / kernel maps / sparse conv fwd+bwd
 *   torch-points-kernels     ball_query(PARTIAL_DENSE) + region_grow
 *   hdbscan                  core distances, mutual-reachability MST, tree -> labels
 *
 * The reference has no FFI of its own for this path: its boundary is Python attribute access on
 * those modules (reference call sites are cited per entry point below, paths relative to the
 * reference root).  This header is the boundary underneath the Python mirrors in
 * panopticsegforlargescalepointcloud_b200/{me,tpk,hdbscan}.py.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless its name ends in _host;
 *   - no hidden device allocation: scratch space is passed in (sizes from the *_bytes queries);
 *   - every launch goes to `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = ok, non-zero = error, text from pgs_last_error() (thread local);
 *   - handles: none.  All state lives in caller-owned buffers, so calls are thread-compatible.
 *   - row indices are int32 (N < 2^31), feature rows are fp32 row-major [N, C].
 */
#ifndef PGS_B200_H_
#define PGS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGS_OK 0
#define PGS_ERR_INVALID 1
#define PGS_ERR_CUDA 2
#define PGS_ERR_RANGE 3

/* status word bits written by device code (see pgs_cmap_build) */
#define PGS_STATUS_COORD_RANGE 1u /* a coordinate did not fit the 16|16|16|16 key packing */
#define PGS_STATUS_TABLE_FULL 2u  /* hash table capacity exhausted (caller bug)          */

int pgs_version(void);
const char* pgs_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t pgs_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Coordinate map  (replaces ME.SparseTensor's CoordinateManager insert + stride;
 *                  reference: torch_points3d/applications/minkowski.py:121-122,
 *                  modules/MinkowskiEngine/api_modules.py:244-285 (stride-2 conv_in))
 *
 * Key  = b:16 | x:16 | y:16 | z:16, spatial fields biased by +32768.
 * Table = open addressing, linear probing, capacity a power of two >= 2n.
 * Row ids of the new map follow FIRST OCCURRENCE in input order (deterministic; for unique
 * input coordinates and tensor_stride_out == tensor stride of the input the map is the identity).
 * ------------------------------------------------------------------------------------------ */

/* slots needed for n rows (power of two, >= 2n, >= 1024) */
int64_t pgs_cmap_capacity(int64_t n);
/* scratch bytes for pgs_cmap_build */
size_t pgs_cmap_build_scratch_bytes(int64_t n);

/* Build the map of floor(c / tensor_stride_out) * tensor_stride_out over n input rows.
 *   coords        int32 [n,4]  (batch, x, y, z)
 *   tkeys/tvals   table storage, capacity `cap` (from pgs_cmap_capacity); overwritten
 *   out_coords    int32 [n,4]  first *n_out rows valid
 *   in2out        int32 [n]    row of the new map that input row i falls into
 *   n_out         int32 [1]    number of rows of the new map
 *   status        uint32 [1]   OR-ed PGS_STATUS_* bits (caller zeroes it)
 */
int pgs_cmap_build(const int32_t* coords, int64_t n, int32_t tensor_stride_out,
                   uint64_t* tkeys, int32_t* tvals, int64_t cap,
                   int32_t* out_coords, int32_t* in2out, int32_t* n_out,
                   uint32_t* status, void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Kernel map / rulebook  (replaces ME kernel_map for MinkowskiConvolution(Transpose) k=3;
 *                         reference: api_modules.py:26-55,244-270,293)
 *
 * Gather-table form (output stationary):  nbr[k * n_q + q] = row r of the probed map with
 *     c_r == c_q + sign * delta_k * step        (else -1)
 * delta_k enumerates {-1,0,1}^3 with x fastest: k = (dx+1) + 3(dy+1) + 9(dz+1).
 *   conv k3 s1 on a map        : q = rows of that map, probe the same map,  sign=+1, step=t
 *   conv k3 s2 fine->coarse    : q = coarse rows, probe the fine map,       sign=+1, step=t_fine
 *   transposed s2 coarse->fine : q = fine rows,   probe the coarse map,     sign=-1, step=t_fine
 * ------------------------------------------------------------------------------------------ */
int pgs_kmap_build(const int32_t* q_coords, int64_t n_q,
                   const uint64_t* tkeys, const int32_t* tvals, int64_t cap,
                   int32_t step, int32_t sign, int32_t ksize,
                   int32_t* nbr, void* stream);

/* ME-style rulebook (pair lists grouped by kernel offset, ascending output row inside a group):
 *   pair p in [offs[k], offs[k+1]):  in_idx[p] -> out_idx[p] through weight k.
 * in_idx/out_idx must hold K*n_q entries (upper bound); offs holds K+1 int32. */
size_t pgs_kmap_pairs_scratch_bytes(int64_t n_q, int32_t K);
int pgs_kmap_pairs(const int32_t* nbr, int64_t n_q, int32_t K,
                   int32_t* in_idx, int32_t* out_idx, int32_t* offs,
                   void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sparse convolution  (replaces ME ConvolutionForward/BackwardKernelGPU;
 *                      reference: api_modules.py:26-55 (ResBlock), 244-270 (ResNetDown), 293)
 *
 *   fwd        : Y[q] = sum_k X[nbr[tk(k)][q]] * W[k]           W fp32 [K, Cin, Cout]
 *   bwd input  : the same kernel with the sibling table and w_transposed=1
 *                (dX[q] = sum_k dY[nbr'[tk(k)][q]] * W[k]^T)
 *   bwd weight : dW[k] += sum_{pairs of k} X[in]^T dY[out]
 * tk(k) = mirror ? K-1-k : k.  nbr == NULL means K == 1 with the identity map.
 * ------------------------------------------------------------------------------------------ */
int pgs_conv_fwd(const float* X, const float* W, const int32_t* nbr, int64_t n_q,
                 int32_t K, int32_t c_in, int32_t c_out, int32_t mirror, int32_t w_transposed,
                 float* Y, void* stream);

/* Occupancy-sorted gather table for the tensor-core conv kernels.  Those kernels skip every (16- or 128-row tile,
 * kernel offset) pair without a neighbour; rows in arbitrary order leave almost none (a 16-row union has ~26 of 27
 * offsets although a row has 2..13).  pgs_kmap_row_masks writes one K-bit occupancy mask per row (K <= 31; bit order:
 * centre, faces, edges, corners = most significant); the caller sorts the masks (stable) to get `order`, and
 * pgs_kmap_permute writes nbr_sorted[k][r] = nbr[k][order[r]].  Passing (nbr_sorted, order) to a tensor-core entry
 * point gives the same rows as (nbr, NULL): tile row r is output row order[r]. */
int pgs_kmap_row_masks(const int32_t* nbr, int64_t n_q, int32_t K, int32_t* masks, void* stream);

int pgs_kmap_permute(const int32_t* nbr, int64_t n_q, int32_t K, const int32_t* order, int32_t* nbr_sorted,
                     void* stream);

/* tcgen05 (5th-generation tensor core) variant of pgs_conv_fwd: same result contract, fp32 in / fp32 out,
 * 3-pass tf32 hi/lo split with fp32 accumulation in tensor memory.  Supported when
 * pgs_conv_tc_supported(c_in, c_out) (c_in % 16 == 0, c_out % 16 == 0, 16 <= c_out <= 192).
 * scratch holds the re-arranged weights (pgs_conv_tc_scratch_bytes); W == NULL: scratch was filled by
 * pgs_conv_prep_weights_batch (also accepted by pgs_conv_fwd_mma / pgs_conv_fwd_mma_split). */
int pgs_conv_tc_supported(int32_t c_in, int32_t c_out);
size_t pgs_conv_tc_scratch_bytes(int32_t K, int32_t c_in, int32_t c_out);
int pgs_conv_fwd_tc(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q,
                    int32_t K, int32_t c_in, int32_t c_out, int32_t mirror, int32_t w_transposed,
                    float* Y, void* scratch, size_t scratch_bytes, void* stream);

/* Register-operand tensor-core variant for the narrow layers (mma.sync m16n8k8 tf32, A fragments gathered
 * straight from global memory, no shared-memory staging of the feature rows): same result contract and the same
 * 3-pass tf32 hi/lo split as pgs_conv_fwd_tc.  Supported when pgs_conv_mma_supported(c_in, c_out)
 * (both multiples of 16 in 16..64); nbr must not be NULL.  scratch holds the fragment-ordered hi/lo weights
 * (pgs_conv_mma_scratch_bytes). */
int pgs_conv_mma_supported(int32_t c_in, int32_t c_out);
size_t pgs_conv_mma_scratch_bytes(int32_t K, int32_t c_in, int32_t c_out);
int pgs_conv_fwd_mma(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q,
                     int32_t K, int32_t c_in, int32_t c_out, int32_t mirror, int32_t w_transposed,
                     float* Y, void* scratch, size_t scratch_bytes, void* stream);

/* Few-row variant of pgs_conv_fwd_mma for the coarse U-Net levels (latency bound: a few MFLOP, up to 4 MB of
 * weights): one warp per (16 rows, 16 output channels, part of the kernel offsets), partial sums meet in Y by
 * atomicAdd.  Any channel counts that are multiples of 16.  Same scratch as pgs_conv_fwd_mma. */
int pgs_conv_mma_split_supported(int32_t c_in, int32_t c_out);
int pgs_conv_fwd_mma_split(const float* X, const float* W, const int32_t* nbr, const int32_t* order, int64_t n_q,
                           int32_t K, int32_t c_in, int32_t c_out, int32_t mirror, int32_t w_transposed,
                           float* Y, void* scratch, size_t scratch_bytes, void* stream);

/* All weight re-arrangements of a pass in one launch (the fused executor prepares the forward and the backward
 * layout of every convolution at the start of a step).  desc: device array of n_desc records of 8 int64:
 * { W, dst, K, c_in, c_out, w_transposed, layout (0 = pgs_conv_fwd_tc, 1 = pgs_conv_fwd_mma / _split), 0 } with c_in /
 * c_out as the conv entry point will see them; dst = the `scratch` later passed to that entry point together with
 * W == NULL ("weights already arranged").  max_elems = largest K * c_in * c_out of the batch (sizes the grid). */
int pgs_conv_prep_weights_batch(const int64_t* desc, int32_t n_desc, int64_t max_elems, void* stream);

/* dW must be zeroed by the caller (accumulates).  in_idx/out_idx/offs (device) from pgs_kmap_pairs
 * of the FORWARD table; max_pairs >= max_k (offs[k+1]-offs[k]) (host value, sizes the grid; callers pass the row count
 * of the output side, which bounds it without a device read-back).
 * in_idx == out_idx == offs == NULL: K == 1 identity pairs 0..max_pairs-1. */
int pgs_conv_bwd_weight(const float* X, const float* dY,
                        const int32_t* in_idx, const int32_t* out_idx, const int32_t* offs,
                        int64_t max_pairs, int32_t K, int32_t c_in, int32_t c_out, int32_t mirror,
                        float* dW, void* stream);

/* Tensor-core weight gradient (mma.sync tf32 + bf16 correction, pair rows loaded straight into the fragments with one
 * vector load per lane and row piece);
 * pgs_conv_bwd_weight forwards to it when both channel counts are multiples of 16 (PGS_DW_IMPL=ffma disables). */
int pgs_conv_dw_mma_supported(int32_t c_in, int32_t c_out);
int pgs_conv_bwd_weight_mma(const float* X, const float* dY,
                            const int32_t* in_idx, const int32_t* out_idx, const int32_t* offs,
                            int64_t max_pairs, int32_t K, int32_t c_in, int32_t c_out, int32_t mirror,
                            float* dW, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ball query + region growing  (replaces torch_points_kernels.ball_query(mode="PARTIAL_DENSE") and
 *                               torch_points_kernels.region_grow -- un-vendored dependency, tpk 0.7.0;
 *                               reference: torch_points3d/models/panoptic/PointGroup3heads.py:166-174,
 *                               185-202,296-304,340-357; core/spatial_ops/neighbour_finder.py:35-37,164)
 *
 * gid[i] >= 0 is the group (scene, or semantic class x scene) of support/query point i; points only see
 * points of their own group; gid < 0 = not a support point / no neighbours.  gid < 32767.
 * Neighbour list of a query = the FIRST nsample support points of its group, in ascending row index,
 * with fp32 d2 = fma(dz,dz,fma(dy,dy,dx*dx)) <= radius*radius  (the upstream kernel's scan order).
 *
 * Grid: cells of edge `cell` >= radius; rows sorted by (gid, cell) then row index.
 *   spos        float4 [n]     (x, y, z, bit-cast row index) in sorted order
 *   skeys       uint64 [n]     sorted cell keys (all-ones = not a support point)
 *   tkeys/tvals hash cell key -> cell ordinal, capacity `cap` (pgs_cmap_capacity(n))
 *   cell_start  int32 [n+1]    rows of cell c are [cell_start[c], cell_start[c+1])
 *   meta        int32 [2]      {#support rows, #cells}
 *   status      uint32 [1]     PGS_STATUS_COORD_RANGE if a point does not fit +-32766 cells (caller zeroes)
 * ------------------------------------------------------------------------------------------ */
size_t pgs_bq_grid_scratch_bytes(int64_t n);
int pgs_bq_grid_build(const float* pos, const int32_t* gid, int64_t n, float cell,
                      uint64_t* tkeys, int32_t* tvals, int64_t cap,
                      float* spos, uint64_t* skeys, int32_t* cell_start, int32_t* meta,
                      uint32_t* status, void* scratch, size_t scratch_bytes, void* stream);

/* queries that are not the support set: qpos float4 [n] (x,y,z,row), qkeys uint64 [n], input order */
int pgs_bq_pack_queries(const float* pos, const int32_t* gid, int64_t n, float cell,
                        float* qpos, uint64_t* qkeys, void* stream);

/* One warp per query.  For self-queries pass qpos = spos, qkeys = skeys.
 *   nbr  int32 [n_rows, nsample]  row = query's row index; first cnt[row] entries valid (unordered unless
 *                                 the row is full, in which case it holds the nsample smallest indices)
 *   cnt  int32 [n_rows]           zeroed here for the first n_q rows
 *   dist fp32  [n_rows, nsample]  squared distances, same layout (may be NULL) */
int pgs_bq_query(const float* spos, const float* qpos, const uint64_t* qkeys, int64_t n_q,
                 const uint64_t* tkeys, const int32_t* tvals, int64_t cap, const int32_t* cell_start,
                 float radius, int32_t nsample, int32_t* nbr, int32_t* cnt, float* dist, void* stream);

/* reference layout: idx int64 [n, nsample] ascending, -1 padded; dist2 fp32 [n, nsample], -1 padded */
int pgs_bq_export(const int32_t* nbr, const float* dist, const int32_t* cnt, int64_t n, int32_t nsample,
                  int64_t* idx_out, float* dist_out, void* stream);

/* Region growing == min-ancestor labelling over the directed edges q -> nbr[q][*]:
 * label[v] = smallest row index that reaches v.  pgs_rg_propagate runs `rounds` (push + pointer-jump)
 * sweeps and leaves *changed != 0 if any label moved; call until it reads 0.
 * pushed (nullable): int32 [n] frontier state, filled with -1 by the caller before the first call -- a row only pushes
 * again when its label moved since its last push (exact: atomicMin is monotone); NULL = every row pushes in every sweep. */
int pgs_rg_init(const int32_t* gid, int64_t n, int32_t* label, void* stream);
int pgs_rg_propagate(const int32_t* nbr, const int32_t* cnt, const int32_t* gid, int64_t n, int32_t nsample,
                     int32_t rounds, int32_t* label, int32_t* pushed, int32_t* changed, void* stream);

/* Nearest support point of every query (k = 1) on the grid built by pgs_bq_grid_build (cell >= the typical point spacing);
 * replaces torch_geometric.nn.knn(x, y, k=1) of the eval-time back-projection (reference:
 * torch_points3d/metrics/panoptic_tracker_pointgroup_npm3d.py:384 (block -> original points), :592 (subsampled -> full
 * cloud)).  qpos / qkeys from pgs_bq_pack_queries with the SAME cell; meta = the grid's {#support rows, #cells}.
 *   idx_out int32 [n_q]  support row with the smallest fp32 d2 = fma(dz,dz,fma(dy,dy,dx*dx)), ties to the smaller index
 *                        (-1: no support point in the query's group);   d2_out fp32 [n_q] that distance (inf if none).
 * Rings of cells are searched outwards up to max_ring; queries that still found nothing scan all support rows. */
int pgs_nn1_query(const float* spos, const float* qpos, const uint64_t* qkeys, int64_t n_q,
                  const uint64_t* tkeys, const int32_t* tvals, int64_t cap, const int32_t* cell_start,
                  const int32_t* meta, float cell, int32_t max_ring, int32_t* idx_out, float* d2_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Proposal bookkeeping after clustering  (replaces torch_points_kernels.instance_iou -- reference call sites
 *     torch_points3d/core/losses/panoptic_losses.py:37,126, every panoptic tracker -- and the dense-mask cross IoU + numpy
 *     greedy NMS of torch_points3d/models/panoptic/structure_3heads.py:6-17,40-61)
 * Proposals in CSR form: flat int64 [n_flat] point ids, offs int32 [n_prop + 1].
 *   pgs_prop_gt_iou   : gt_id int32 [N] = global ground-truth instance of a point (-1 none), gt_size / gt_scene int32
 *                       [total_gt], prop_scene int32 [n_prop];  inter int32 [n_prop, total_gt] (work), iou fp32 same shape:
 *                       |P & G| / |P u G| for instances of the proposal's own scene, 0 elsewhere.
 *   pgs_prop_cross_nms: inter int32 [n_prop, n_prop] receives |P_i & P_j| (i != j); rank_order int32 [n_prop] = proposals in
 *                       descending score order; keep uint8 [n_prop] = 1 for the proposals greedy NMS picks
 *                       (a picked proposal removes every later one with cross IoU > threshold).
 * ------------------------------------------------------------------------------------------ */
int pgs_prop_gt_iou(const int64_t* flat, const int32_t* offs, int32_t n_prop, int64_t n_flat, const int32_t* gt_id,
                    int32_t total_gt, const int32_t* gt_size, const int32_t* gt_scene, const int32_t* prop_scene,
                    int32_t* inter, float* iou, void* stream);
size_t pgs_prop_nms_scratch_bytes(int64_t n_flat);
int pgs_prop_cross_nms(const int64_t* flat, const int32_t* offs, int32_t n_prop, int64_t n_flat,
                       const int32_t* rank_order, float threshold, int32_t* inter, uint8_t* keep, void* scratch,
                       size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * HDBSCAN  (replaces hdbscan.HDBSCAN(min_cluster_size, min_samples, cluster_selection_epsilon).fit_predict --
 *           un-vendored dependency hdbscan 0.8.27; reference: torch_points3d/utils/hdbscan_cluster.py:8-13,
 *           117-167; models/panoptic/pointgroupembed.py:240-245,704; models/panoptic/pointgroup.py:208-212)
 *
 * pgs_hdb_mst (device): float64 core distances (distance to the min_samples-th nearest sample counting the
 * sample itself) and the exact minimum spanning tree of the mutual-reachability graph
 * w(a,b) = max(core_a, core_b, d(a,b)/alpha) under the strict order (w, min(a,b), max(a,b)).
 *   X     fp32 [n, D] row-major, 1 <= D <= 8            core  fp64 [n]
 *   u, v  int32 [n-1] (u < v, original row ids)         w     fp64 [n-1]   sorted by the strict order
 * Synchronises `stream` once per Boruvka round (the round count is data dependent); *rounds_host receives it.
 *
 * pgs_hdb_labels_host (host, sequential O(n)): sorted MST -> single linkage -> condensed tree -> excess of
 * mass + epsilon selection -> labels (-1 = noise; clusters numbered by ascending condensed-tree id).
 * All pointers of this function are HOST pointers.
 * ------------------------------------------------------------------------------------------ */
size_t pgs_hdb_scratch_bytes(int64_t n, int32_t D);
/* diagnostics of the last pgs_hdb_mst call: out_host int64 [max_rounds][4] = per Boruvka round {box-test steps (32 group or
 * leaf boxes each), candidate blocks evaluated point by point, point pairs whose distance was computed, warps} */
int pgs_hdb_search_stats(int64_t* out_host, int32_t max_rounds);
int pgs_hdb_mst(const float* X, int64_t n, int32_t D, int32_t min_samples, double alpha,
                double* core, int32_t* u, int32_t* v, double* w, int32_t* rounds_host,
                void* scratch, size_t scratch_bytes, void* stream);
/* rank_out int32 [n] (device): position of every input row in the Morton order pgs_hdb_mst sorted the points by, read
 * from the `scratch` of that call (same n, D).  Relabelling the MST's endpoints with it (keeping each edge's (u, v) order)
 * gives the host tree stage an isomorphic tree whose leaves sit next to their spatial neighbours -- the union-find and
 * the condensed-tree walk then stay in cache; the labels come back per rank and are permuted back by the caller. */
int pgs_hdb_morton_rank(const void* scratch, int64_t n, int32_t D, int32_t* rank_out, void* stream);
int pgs_hdb_labels_host(const int32_t* u_host, const int32_t* v_host, const double* w_host, int64_t n,
                        int32_t min_cluster_size, double cluster_selection_epsilon,
                        int32_t* labels_host, int32_t* n_clusters_host);

/* ------------------------------------------------------------------------------------------
 * Fused BatchNorm (+ ReLU) on a feature matrix  (replaces ME.MinkowskiBatchNorm [+ ME.MinkowskiReLU];
 *                    reference: modules/MinkowskiEngine/api_modules.py:40-41,53-54,269-270)
 * X, Y, dY, dX fp32 [n, C] row-major, C % 4 == 0.  nn.BatchNorm1d semantics: batch statistics over all rows
 * (biased variance) in training, running estimates in eval; running_var gets the unbiased variance.
 * sums: fp64 [2C] scratch.  save_mean / save_invstd: fp32 [C], written by forward, read by backward.
 * relu != 0 fuses max(., 0) into the forward and its mask (Y > 0) into the backward.
 * dweight / dbias receive sum(g * xhat) / sum(g) (overwritten, not accumulated).
 * ------------------------------------------------------------------------------------------ */
int pgs_bn_forward(const float* X, int64_t n, int32_t C, const float* weight, const float* bias,
                   float* running_mean, float* running_var, int32_t training, float momentum, float eps,
                   int32_t relu, double* sums, float* save_mean, float* save_invstd, float* Y, void* stream);
int pgs_bn_backward(const float* X, const float* Y, const float* dY, int64_t n, int32_t C, const float* weight,
                    const float* save_mean, const float* save_invstd, int32_t training, int32_t relu, double* sums,
                    float* dX, float* dweight, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------
 * Flat-kernel mean shift on embeddings  (replaces sklearn.cluster.MeanShift(bandwidth, bin_seeding=True).fit as called
 *                                        by torch_points3d/utils/meanshift_cluster.py:9-18,72-123; call sites
 *                                        models/panoptic/PointGroup3heads.py:235,276,323,376, pointgroupembed.py:491,532)
 *   pgs_ms_iterate : every seed climbs to its mode (mean of the points within `bandwidth`, until it moves less than
 *                    1e-3 * bandwidth or max_iter); centers fp32 [n_seeds, D], counts = points within the bandwidth at the
 *                    last query (0: nothing near the seed, drop it), iters = completed iterations
 *   pgs_ms_assign  : label = index of the nearest centre (fp64 distances, ties to the lower index); dist nullable
 * D in 1..8.  Seeding (grid bins), the greedy removal of near-duplicate centres and the label order live in
 * meanshift.py (a few thousand centres: host work, like the reference).
 * ------------------------------------------------------------------------------------------ */
int pgs_ms_iterate(const float* X, int64_t n, int32_t D, const float* seeds, int64_t n_seeds, float bandwidth,
                   int32_t max_iter, float* centers, int32_t* counts, int32_t* iters, void* stream);
int pgs_ms_assign(const float* X, int64_t n, int32_t D, const float* centers, int32_t n_centers, int32_t* labels,
                  double* dist, void* stream);

/* flags of the _ex variants (used by the fused U-Net executor, fastpath.py) */
#define PGS_BN_ACCUMULATE_PARAM_GRADS 1 /* dweight / dbias += instead of = (write straight into param.grad) */
#define PGS_BN_SUMS_ZEROED 2            /* the caller zeroed `sums` (one memset for all layers of a pass) */
int pgs_bn_forward_ex(const float* X, int64_t n, int32_t C, const float* weight, const float* bias,
                      float* running_mean, float* running_var, int32_t training, float momentum, float eps,
                      int32_t relu, int32_t flags, double* sums, float* save_mean, float* save_invstd, float* Y,
                      void* stream);
/* backward_ex: with relu != 0 and Y == NULL the mask [y > 0] is recomputed from X (y = fma(x, invstd * w, bias - mean *
 * invstd * w) with the forward's own roundings), which saves reading Y; `bias` is the forward's bias (NULL = 0) and is
 * only used for that. */
int pgs_bn_backward_ex(const float* X, const float* Y, const float* dY, int64_t n, int32_t C, const float* weight,
                       const float* bias, const float* save_mean, const float* save_invstd, int32_t training,
                       int32_t relu, int32_t flags, double* sums, float* dX, float* dweight, float* dbias,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Feature-matrix glue of the U-Net (replaces `a + b` and ME.cat on SparseTensors sharing a coordinate map;
 * reference: api_modules.py:76-82 (residual sum), 306-311 (skip concatenation)).
 *   pgs_add2 : y = a + b over n_elems floats (multiple of 4)
 *   pgs_cat2 : split == 0: y[r] = [a[r] | b[r]]   split == 1: a[r], b[r] = the two parts of y[r]  (backward of cat)
 * ------------------------------------------------------------------------------------------ */
int pgs_add2(const float* a, const float* b, float* y, int64_t n_elems, void* stream);
int pgs_cat2(float* a, int32_t ca, float* b, int32_t cb, float* y, int64_t n, int32_t split, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-network executor  (replaces the per-module Python dispatch of MinkowskiUnet.forward and of the autograd
 *                          backward through it; reference: torch_points3d/applications/minkowski.py:160-196,
 *                          modules/MinkowskiEngine/api_modules.py:76-82,281-285,306-311)
 *
 * The U-Net's structure is static: a tape of conv / bn(+relu) / add / cat ops over numbered feature slots (slot 0 =
 * network input).  The host fills one record per convolution / batch norm with this step's device pointers (weights,
 * kernel-map tables, arranged-weight scratch) and calls ONE function per direction; the functions only walk the tape
 * and launch the kernels declared above on `stream` (no allocation, no synchronisation, nothing data dependent).
 * All arrays of these two calls are HOST arrays; the pointers inside the records are device pointers.
 *
 * kind_f / kind_b select the conv entry point: 0 pgs_conv_fwd (FFMA), 1 pgs_conv_fwd_tc, 2 pgs_conv_fwd_mma,
 * 3 pgs_conv_fwd_mma_split.  For 1..3 the arranged weights must already be in wprep_f / wprep_b
 * (pgs_conv_prep_weights_batch).  nbr_* == NULL: K == 1 identity map.
 * ------------------------------------------------------------------------------------------ */
#define PGS_OP_CONV 0
#define PGS_OP_BN 1
#define PGS_OP_ADD 2
#define PGS_OP_CAT 3

typedef struct {
  int32_t kind, a, b, dst; /* b = -1 unless add / cat */
  int32_t idx;             /* index into convs / bns */
  int32_t relu;            /* bn: fuse max(., 0) */
} pgs_unet_op;

typedef struct {
  const float* W;          /* fp32 [K, c_in, c_out] */
  float* dW;               /* accumulated into (caller zeroes); NULL: no weight gradient */
  const int32_t* nbr_f;    /* forward gather table (rows of the output map) and its occupancy order (or NULL) */
  const int32_t* order_f;
  const int32_t* nbr_b;    /* sibling table for the input gradient (rows of the input map) */
  const int32_t* order_b;
  const int32_t* pair_in;  /* pgs_kmap_pairs of the forward table (weight gradient); NULL with K == 1 */
  const int32_t* pair_out;
  const int32_t* pair_offs;
  void* wprep_f;           /* arranged weights, forward / backward layout */
  void* wprep_b;
  uint64_t wprep_bytes;
  int64_t max_pairs;
  int32_t K, c_in, c_out;
  int32_t kind_f, kind_b;
  int32_t mirror_f, mirror_b;
  int32_t need_dx;         /* 0: skip the input gradient (network input that needs none) */
} pgs_unet_conv;

typedef struct {
  const float* weight;
  const float* bias;
  float* running_mean;
  float* running_var;
  float* dweight;          /* NULL: not needed */
  float* dbias;
  float momentum, eps;
  int32_t training;
  int32_t accumulate;      /* dweight / dbias += (they alias param.grad) instead of = */
} pgs_unet_bn;

/* sizeof(pgs_unet_op), sizeof(pgs_unet_conv), sizeof(pgs_unet_bn): lets a foreign-language binding verify its mirrors */
void pgs_unet_record_bytes(int32_t* out3);

/* slot_ptr[s]: fp32 [slot_n[s], slot_c[s]] activation of slot s (slot 0 = input, the rest caller-allocated).
 * sums: fp64, zeroed by the caller; stats: fp32; both indexed by stat_off[bn] (2*C entries per batch norm). */
int pgs_unet_forward(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots,
                     float* const* slot_ptr, const int64_t* slot_n, const int32_t* slot_c,
                     const pgs_unet_conv* convs, const pgs_unet_bn* bns,
                     double* sums, float* stats, const int64_t* stat_off, void* stream);

/* Tape in reverse.  d_out: gradient of slot `out_slot`.  garena: scratch for every intermediate gradient
 * (pgs_unet_backward_scratch_elems floats).  Weight gradients run on `side_stream` when it is non-NULL (they only
 * depend on a layer's input activation and output gradient); the function makes `stream` wait for them before it
 * returns.  grad_in receives the gradient(s) of slot 0 (device pointers into garena, to be summed; at most 8) and
 * *n_grad_in their number (0 when no convolution reading slot 0 has need_dx). */
int64_t pgs_unet_backward_scratch_elems(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots,
                                        const int64_t* slot_n, const int32_t* slot_c);
int pgs_unet_backward(const pgs_unet_op* ops, int32_t n_ops, int32_t n_slots, int32_t out_slot,
                      float* const* slot_ptr, const int64_t* slot_n, const int32_t* slot_c,
                      const pgs_unet_conv* convs, const pgs_unet_bn* bns,
                      double* sums, const float* stats, const int64_t* stat_off,
                      const float* d_out, float* garena, int64_t garena_elems,
                      float** grad_in, int32_t* n_grad_in, void* stream, void* side_stream);

#ifdef __cplusplus
}
#endif
#endif /* PGS_B200_H_ */
