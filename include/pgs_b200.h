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

/* dW must be zeroed by the caller (accumulates).  in_idx/out_idx/offs (device) from pgs_kmap_pairs
 * of the FORWARD table; max_pairs = max_k (offs[k+1]-offs[k]) (host value, sizes the grid).
 * in_idx == out_idx == offs == NULL: K == 1 identity pairs 0..max_pairs-1. */
int pgs_conv_bwd_weight(const float* X, const float* dY,
                        const int32_t* in_idx, const int32_t* out_idx, const int32_t* offs,
                        int64_t max_pairs, int32_t K, int32_t c_in, int32_t c_out, int32_t mirror,
                        float* dW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PGS_B200_H_ */
