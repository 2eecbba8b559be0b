/* rpgp.h -- C ABI of librpgp.so: the B200-native (sm_100a) matrix-free kernel-matrix multiply of
 * Randomly-Projected-Additive-GPs and its quadratic-form derivative.
 *
 * This is the drop-in boundary of SURVEY.md §8(b).  The reference has no FFI of its own (it reaches the arithmetic
 * through GPyTorch LazyTensors and PyKeOps); each entry point below names the reference interface it replaces
 * (paths relative to the reference repository):
 *
 *   rpgp_mvm_fwd_f32      K.V as called by linear_cg: `KeOpsLazyTensor._matmul` / `SumLazyTensor._matmul`, built at
 *                         gp_models/kernels/imq_kernel.py:32-58 (KeOps pattern), driven from
 *                         training_routines.py:515-517,536-537 and fitting/optimizing.py:65-74
 *   rpgp_quad_bwd_f32     `LazyTensor._quad_form_derivative(L, R)`; explicit-formula analogue
 *                         gp_models/kernels/memory_efficient_gam_kernel.py:33-59 (GAMFunction.backward)
 *   rpgp_project_f32      gp_models/kernels/scaled_projection_kernel.py:21-37 (ScaledProjectionKernel.forward) and
 *                         gp_models/kernels/polynomial_projection_kernels.py:115-137 (_project)
 *   rpgp_kernel_rows_f32  dense rows / blocks of K: gp_models/kernels/memory_efficient_gam_kernel.py:12-30
 *                         (GAMFunction.forward), the pivoted-Cholesky row fetch and `diag`
 *   rpgp_*_f64            the `--double` path (gp_experiment_runner.py:252,313-315; training_routines.py:481)
 *   rpgp_kmv_host_f32     the whole path for host buffers: projection + K.V, as one call (what a ctypes/cffi
 *                         binding inside the reference's kernel classes would use)
 *
 * Conventions
 *   - all device pointers are borrowed for the duration of the (asynchronous) launch; nothing is retained.
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises except the *_host ones.
 *   - every function returns 0 on success, otherwise an rpgp_status and rpgp_last_error() describes the failure
 *     (the Python shim turns it into RuntimeError, mirroring the ValueError at memory_efficient_gam_kernel.py:15-16).
 *   - "packed" coordinates: the n x (J*K) scaled projections Z^ are stored, pre-multiplied by sqrt(log2(e)/2), as
 *     [nchunks][n][CP] float planes; chunk c holds projection groups c*G .. c*G+G-1, group g at columns
 *     g*KP .. g*KP+K-1, zero padded.  rpgp_plan_layout() chooses (CP, nchunks, KP, G) for a given (J, K);
 *     rpgp_pack_coords_f32 / rpgp_project_f32 write the layout.  Right-hand sides are [n][TP] float, zero padded,
 *     TP = rpgp_padded_rhs(layout, t, backward).  `neg_log2c` is [nchunks*G] float: -log2(c_j), +inf for padding groups.
 */
#ifndef RPGP_H
#define RPGP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RPGP_OK = 0,
    RPGP_ERR_INVALID = 1,
    RPGP_ERR_CUDA = 2,
    RPGP_ERR_UNSUPPORTED = 3,
    RPGP_ERR_WORKSPACE = 4
} rpgp_status;

typedef struct {
    int J, K;      /* projections, coordinates per projection */
    int CP;        /* packed coordinates per chunk (multiple of 4, <= 32) */
    int nchunks;   /* coordinate chunks */
    int KP;        /* padded coordinates per group (1 when K == 1) */
    int G;         /* groups per chunk */
    int base;      /* base kernel of every group (rpgp_base_kernel); rpgp_plan_layout sets RBF */
} rpgp_layout;

/* base kernels k(d^2) of a projection group (training_routines.py:57-83 `_map_to_kernel`; gp_models/kernels/imq_kernel.py:8-9,47):
 *   RBF exp(-d^2/2);  Matern nu=1.5 (1 + sqrt3 d) exp(-sqrt3 d);  inverse multiquadric (d^2 + 1)^-1/2;  cosine cos(d) (gpytorch's
 *   CosineKernel, training_routines.py:76-81: the caller folds pi / period_length into the coordinates).
 * Every base kernel runs in the SIMT forward / gradient kernels and in the symmetric tensor-core kernel with the same layouts (K = 1:
 * one coordinate and one MUFU per projection -- the distance is |d|, no square root); the distance-on-tensor-core variant of the
 * symmetric kernel is RBF only (its exponent IS the squared distance), and the cosine kernel with K > 1 takes the rectangular kernel
 * (rpgp_mvm_sym_supported says so).  Non-RBF kernels take right-hand-side widths 4 and 16 only. */
typedef enum { RPGP_BASE_RBF = 0, RPGP_BASE_MATERN15 = 1, RPGP_BASE_INVERSE_MQ = 2, RPGP_BASE_COSINE = 3 } rpgp_base_kernel;

int rpgp_version(void);
const char* rpgp_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py `gpu_launches`) */
unsigned long long rpgp_launch_count(void);

/* layout planning ------------------------------------------------------------------------------------------------ */
int rpgp_plan_layout(int J, int K, rpgp_layout* out);
int rpgp_plan_layout_base(int J, int K, int base, rpgp_layout* out);
/* padded right-hand-side width TP the kernels are compiled for; 0 when t exceeds rpgp_max_rhs (chunk the columns) */
int rpgp_padded_rhs(const rpgp_layout* lay, int t, int backward);
int rpgp_max_rhs(const rpgp_layout* lay, int backward);
double rpgp_coord_scale(void);            /* sqrt(log2(e)/2) */

/* natural (n x J*K, row stride ld) float coordinates -> packed planes, multiplied by `scale` */
int rpgp_pack_coords_f32(const float* Z, int64_t n, int64_t ld, const rpgp_layout* lay, float scale, float* Zp,
                         void* stream);
/* c (J floats, device) -> neg_log2c (nchunks*G floats, device) */
int rpgp_pack_log2c_f32(const float* c, const rpgp_layout* lay, float* neg_log2c, void* stream);

/* projection: Zp = pack( scale * post_inv[q] * sum_k X[i,k] * pre_inv[k] * W[q,k] ), FP64 accumulation.
 * X: n x d (row stride ldx), W: (J*K) x d row-major (the torch.nn.Linear weight), pre_inv: d or NULL, post_inv: J*K
 * or NULL (reciprocal lengthscales; prescale / postscale of scaled_projection_kernel.py:23-27). */
int rpgp_project_f32(const float* X, int64_t n, int d, int64_t ldx, const float* W, const float* pre_inv,
                     const float* post_inv, const rpgp_layout* lay, float scale, float* Zp, void* stream);

/* rpgp_project_f32 runs on the tensor cores (tcgen05 kind::tf32, 3xTF32 split, FP32 accumulation in tensor memory; csrc/project_tc.cu)
 * when d <= 128 and J*K <= 112 (rpgp_project_tc_supported), otherwise on the FP64-accumulating SIMT kernel.
 * rpgp_project2_f32: the same product with two optional outputs -- the packed planes Zp and / or the natural n x (J*K) matrix Zn (row
 * stride ldz, NOT multiplied by `scale`), which is what the kernel classes' autograd graph carries.
 * rpgp_project_bwd_f32: the vector-Jacobian product of the projection, dW[q][k] = sum_i dZ[i][q] X[i][k] (J*K x d, row-major): with
 * Z[i][q] = post_inv[q] sum_k X[i][k] pre_inv[k] W[q][k] the gradients are dW = post_inv pre_inv^T * dW', d pre_inv[k] = sum_q post_inv[q]
 * W[q][k] dW'[q][k], d post_inv[q] = sum_k pre_inv[k] W[q][k] dW'[q][k]  (what torch autograd derives from nn.Linear + div in
 * scaled_projection_kernel.py:21-27).  Deterministic (per-CTA partials, FP64 reduction in a fixed order). */
int rpgp_project_tc_supported(int d, const rpgp_layout* lay);
int rpgp_project2_f32(const float* X, int64_t n, int d, int64_t ldx, const float* W, const float* pre_inv, const float* post_inv,
                      const rpgp_layout* lay, float scale, float* Zp, float* Zn, int64_t ldz, void* stream);
size_t rpgp_project_bwd_workspace_bytes(int64_t n, int d, int JK);
int rpgp_project_bwd_f32(const float* X, int64_t n, int d, int64_t ldx, const float* dZ, int64_t ldz, int JK, float* dW, void* workspace,
                         size_t workspace_bytes, void* stream);

/* forward K.V ---------------------------------------------------------------------------------------------------- */
size_t rpgp_mvm_workspace_bytes(int64_t m, int64_t n, const rpgp_layout* lay, int t);
/* out[i, 0..t) = sum_i' K[i,i'] V[i',:], i < m.  z1p: [nchunks][m][CP] with plane stride z1_stride (elements);
 * z2p likewise with n.  Pass a row block of the packed planes (pointer + r0*CP, m = r1-r0, same plane stride) to
 * compute the rows owned by one rank.  out: m x ldo. */
int rpgp_mvm_fwd_f32(const float* z1p, int64_t m, int64_t z1_stride, const float* z2p, int64_t n, int64_t z2_stride,
                     const rpgp_layout* lay, const float* neg_log2c, const float* Vp, int t, float* out, int ldo,
                     void* workspace, size_t workspace_bytes, void* stream);

/* symmetric forward K(Z,Z).V on the tensor cores ---------------------------------------------------------------------
 * Every kernel value is evaluated once and used for both out[i] and out[i'] (tcgen05, 3xTF32 split, FP64 global
 * accumulation); any layout of rpgp_plan_layout (K >= 1; coordinate chunks are summed), t <= 16 per call
 * (rpgp_mvm_sym_supported).  zp: [nchunks][n][CP].  Vp16: [n][16] zero-padded right-hand sides.
 * Row blocks are 128 rows; a launch handles the unique block pairs owned by row blocks [row_block_begin, row_block_end)
 * and writes their contributions to ALL n rows of out -- with the full range [0, ceil(n/128)) out is K.V, with a
 * sub-range (one rank of a multi-GPU job) the outputs of the ranks must be summed (all-reduce).
 * Accumulation uses FP64 atomics, so results are reproducible to FP32 rounding but not bit-identical run to run. */
size_t rpgp_mvm_sym_workspace_bytes(int64_t n, const rpgp_layout* lay);
/* K > 1 (4 <= K <= 24): the squared distances |z|^2 + |z'|^2 - 2 z.z' are evaluated on tcgen05 as augmented inner products
 * (3xTF32, centred coordinates) while the centred, scaled squared group norms r2 = |z_ig - mean_g|^2 stay small enough for the
 * cancellation: sqrt(mean r2^2) <= rpgp_mvm_sym_distance_bound() and max r2 <= 10x that.  Statistics written by the operand
 * pre-pass decide per call on the device; beyond the bound the direct-difference kernel runs (no host synchronisation either way).  rpgp_mvm_sym_distance_plan reports the chunking of that path:
 * plan = {supported, groups per chunk, k-steps of 8 per group, 128-byte operand lines per row, chunks}. */
float rpgp_mvm_sym_distance_bound(void);
int rpgp_mvm_sym_distance_plan(const rpgp_layout* lay, int plan[5]);
int rpgp_mvm_sym_supported(const rpgp_layout* lay, int t);
int rpgp_mvm_sym_f32(const float* zp, int64_t n, const rpgp_layout* lay, const float* neg_log2c, const float* Vp16, int t,
                     float* out, int ldo, int row_block_begin, int row_block_end, void* workspace, size_t workspace_bytes,
                     void* stream);

/* quadratic-form derivative ------------------------------------------------------------------------------------------
 * G = sum_col sum_{i,i'} Lrow[i,col] K[i,i'] Rcol[i',col].
 *   dz1p[c][i][q] = dG / d z1p[c][i][q]   (packed, scaled coordinates)
 *   g[c*G+g]      = sum_{i,i'} S[i,i'] k_g[i,i']   ( = dG / d ln c_g ;  dG / d neg_log2c = -ln2 * g )
 * symmetric != 0: z1p is a row block of z2p itself (K(Z,Z)); Rrow / Lcol are then the R rows of the block and the
 * full L, the returned dz1p is the TOTAL derivative w.r.t. the rows' coordinates (both roles) and g is this block's
 * share of the full sum.  symmetric == 0: z1p / z2p independent, Rrow / Lcol ignored. */
size_t rpgp_quad_workspace_bytes(int64_t m, int64_t n, const rpgp_layout* lay, int t);
int rpgp_quad_bwd_f32(const float* z1p, int64_t m, int64_t z1_stride, const float* z2p, int64_t n, int64_t z2_stride,
                      const rpgp_layout* lay, const float* neg_log2c, const float* Lrow, const float* Rrow,
                      const float* Rcol, const float* Lcol, int t, int symmetric, float* dz1p, float* g,
                      void* workspace, size_t workspace_bytes, void* stream);

/* dense rows / blocks of K on natural (un-scaled) coordinates: out[p, i'] = K(Zr[p], Z2[i']); out: P x n */
int rpgp_kernel_rows_f32(const float* Zr, int64_t P, const float* Z2, int64_t n, int64_t ld, int J, int K,
                         const float* c, float* out, int64_t ldo, void* stream);

/* the natural-coordinate entry points (dense rows here, FP64 below) with an explicit base kernel; the plain names are RBF */
int rpgp_kernel_rows_base_f32(const float* Zr, int64_t P, const float* Z2, int64_t n, int64_t ld, int J, int K, int base,
                              const float* c, float* out, int64_t ldo, void* stream);
int rpgp_kernel_rows_base_f64(const double* Zr, int64_t P, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                              const double* c, double* out, int64_t ldo, void* stream);
int rpgp_mvm_fwd_base_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                          const double* c, const double* V, int t, double* out, void* stream);
int rpgp_quad_bwd_base_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K, int base,
                           const double* c, const double* L, const double* R, int t, double* dZ1, double* g, void* stream);

/* FP64 path (natural coordinates, small n): same semantics, un-tiled */
int rpgp_mvm_fwd_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K,
                     const double* c, const double* V, int t, double* out, void* stream);
/* dZ1 (m x J*K, dense) and g (J) must be zeroed by the caller; contributions are accumulated */
int rpgp_quad_bwd_f64(const double* Z1, int64_t m, const double* Z2, int64_t n, int64_t ld, int J, int K,
                      const double* c, const double* L, const double* R, int t, double* dZ1, double* g, void* stream);
int rpgp_kernel_rows_f64(const double* Zr, int64_t P, const double* Z2, int64_t n, int64_t ld, int J, int K,
                         const double* c, double* out, int64_t ldo, void* stream);

/* whole path on HOST buffers (allocates, copies H2D, projects, multiplies, copies D2H, synchronises):
 * out = K(X1, X2) V + diag_add * V   (diag_add only applied when X1 == X2, i.e. X2 == NULL)
 * X1: m x d, X2: n x d or NULL (=> X1), W: (J*K) x d, pre_inv: d or NULL, post_inv: J*K or NULL, c: J, V: n x t. */
int rpgp_kmv_host_f32(const float* X1, int64_t m, const float* X2, int64_t n, int d, const float* W, int J, int K,
                      const float* pre_inv, const float* post_inv, const float* c, const float* V, int t,
                      float diag_add, float* out, int device);

/* the same path as a persistent PLAN (what a binding inside the reference's kernel classes keeps per model): device buffers for
 * X, W, the scales, c, the packed Z^, V, the product and the workspace are allocated ONCE; every call only copies and computes.
 *   rpgp_plan_create        n x d inputs, J x K projections, at most tmax right-hand sides; `stream` = the caller's cudaStream_t
 *                           (e.g. torch's current stream, so that a collective on the product orders naturally) or NULL for a
 *                           stream owned by the plan.
 *   rpgp_plan_set_operator  H2D of X, W, pre_inv / post_inv (may be NULL), c and the projection -> packed Z^; once per MLL step
 *                           (scaled_projection_kernel.py:21-37 recomputes it per Kernel.forward; the CG iterations reuse it).
 *   rpgp_plan_kmv_begin     H2D of V (n x t) and the symmetric product K(X,X) V restricted to the unique block pairs of the 128-row
 *                           blocks [row_block_begin, row_block_end) -- the full range [0, ceil(n/128)) is the whole product, a
 *                           sub-range is one rank's share and rpgp_plan_device_out() (n x t floats, row stride t) must then be
 *                           summed over the ranks (NCCL all-reduce on `stream`) before rpgp_plan_kmv_end.  Asynchronous.
 *                           n < 1024 or a layout outside rpgp_mvm_sym_supported: rows [128*begin, 128*end) by the SIMT forward
 *                           kernel, the other rows of the product are zero (same contract: sum over ranks = K V).
 *   rpgp_plan_kmv_end       product += diag_add * V; D2H of rows [row_begin, row_end) into `out` ((row_end-row_begin) x t); synchronises.
 * Host pointers may be pageable or pinned (pinned makes the copies asynchronous and full-speed). */
typedef struct rpgp_plan rpgp_plan;
int rpgp_plan_create(int64_t n, int d, int J, int K, int tmax, int device, void* stream, rpgp_plan** out);
int rpgp_plan_destroy(rpgp_plan* plan);
int rpgp_plan_set_operator(rpgp_plan* plan, const float* X, const float* W, const float* pre_inv, const float* post_inv,
                           const float* c);
int rpgp_plan_kmv_begin(rpgp_plan* plan, const float* V, int t, int row_block_begin, int row_block_end);
void* rpgp_plan_device_out(rpgp_plan* plan);
int rpgp_plan_kmv_end(rpgp_plan* plan, float diag_add, float* out, int64_t row_begin, int64_t row_end);

/* issue-rate microbenchmarks (roofline denominators): out[4*i..] = {fp32 lane-ops/clk/SM, mufu/clk/SM, ms, MHz} */
int rpgp_measure_peaks(double* out, int max_ops, const char** names);

#ifdef __cplusplus
}
#endif
#endif /* RPGP_H */
