/* libb200gp -- C ABI of the B200-native candidate-kernel scoring path.
 *
 * Drop-in boundary for the arithmetic the reference (lschlessinger1/MS-project, src/autoks) delegates to
 * GPy / NumPy-LAPACK when it scores a candidate kernel:
 *   GPModel.score_model / fit              src/autoks/core/gp_model.py:31-82
 *     -> GPy GP.parameters_changed / ExactGaussianInference.inference (K, jitchol, pdinv, dpotrs, gradients)
 *   compute_kernel                         src/autoks/backend/kernel.py:317-322
 *   HellingerDistanceBuilder               src/autoks/distance/distance.py:126-194
 *
 * Conventions
 *   - one opaque context per GPU; a context is not thread-safe, distinct contexts are independent;
 *   - every pointer marked `d_` is a CALLER-OWNED DEVICE pointer (e.g. torch tensor.data_ptr()), FP64 / int32,
 *     row-major; the library allocates no device memory of its own: the caller hands it one workspace;
 *   - every function returns 0 on success, <0 on API misuse or a CUDA error (text in b200gp_last_error);
 *     NUMERICAL failures are per-item status codes, never a non-zero return;
 *   - all work is enqueued on the context's stream; functions that must read device results on the host
 *     (jitter ladder) synchronise that stream, the others return asynchronously.
 *
 * Kernel programs (compiled on the host from Covariance.postfix_tokens, src/autoks/core/covariance.py:32-45):
 *   int32 quadruples {op, a, b, 0}:  leaf   {B200GP_OP_SE..CONST, active_dim (0-based), theta offset, 0}
 *                                    binary {B200GP_OP_ADD|MUL, lhs token index, rhs token index, 0}
 *   in postfix order; the last token is the root.  prog_off[M+1] counts tokens.
 *   theta is the GPy param_array of each item's kernel followed by Gaussian_noise.variance (constrained
 *   values, i.e. after the Logexp transform); theta_off[M+1] counts doubles (P_i + 1 per item).
 *   Leaf parameter order: SE [variance, lengthscale]; RQ [variance, lengthscale, power];
 *   PER [variance, period, lengthscale]; LIN [variances, shifts]; CONST [variance].
 */
#ifndef B200GP_H
#define B200GP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200gp_ctx b200gp_ctx;

enum { B200GP_OP_SE = 0, B200GP_OP_RQ = 1, B200GP_OP_PER = 2, B200GP_OP_LIN = 3, B200GP_OP_CONST = 4,
       B200GP_OP_ADD = 5, B200GP_OP_MUL = 6,
       B200GP_OP_LOOKUP = 7 /* {op, index_dim, theta_offset}: RBF over a looked-up distance, see b200gp_set_lookup */ };

/* per-item status */
enum { B200GP_OK = 0, B200GP_JITTERED = 1, B200GP_NOT_PD = 2, B200GP_NONFINITE = 3 };

/* gradient quirk bits: reproduce the fork's notebook formulas instead of the exact derivative */
enum { B200GP_QUIRK_PER_LENGTHSCALE = 1, B200GP_QUIRK_LIN_SHIFT = 2 };

enum { B200GP_TILE = 128, B200GP_MAX_DIMS = 16, B200GP_MAX_TOKENS = 64, B200GP_MAX_PARAMS = 128 };

const char* b200gp_version(void);

/* context on `device`, enqueueing on `cuda_stream` (a cudaStream_t; NULL = legacy default stream) */
int b200gp_create(int device, void* cuda_stream, b200gp_ctx** out);
int b200gp_destroy(b200gp_ctx* ctx);
const char* b200gp_last_error(b200gp_ctx* ctx);
int b200gp_set_quirks(b200gp_ctx* ctx, int quirk_mask);

/* Distance table of the OP_LOOKUP leaf: the BOMS surrogate's RBFDistanceBuilderKernelKernel
 * (src/autoks/core/strategies/bayes_opt_gp_strategy.py:27-28; kernel-kernel.ipynb#2) evaluates
 * k(i, j) = variance * exp(-1/2 (D[i][j] / lengthscale)^2) over model indices stored in X.  d_table is n x n row-major
 * (DistanceBuilder.get_kernel(n)); indices outside [0, n) or NaN entries give NaN.  (NULL, 0) clears it.  The pointer is
 * borrowed: it must stay valid while programs with OP_LOOKUP leaves are evaluated. */
int b200gp_set_lookup(b200gp_ctx* ctx, const double* d_table, int n);

/* Workspace needed to score `slots` matrices of order N at a time (bytes). */
size_t b200gp_workspace_bytes(int N, int D, int slots, int max_params);

/* Bind the training data and the workspace.  d_X is N x D row-major, d_y has N entries (GPy's Y is N x 1;
 * ModelSelector._prepare_data standardises both, base.py:200-222).  The library keeps padded,
 * dimension-major copies inside the workspace. */
int b200gp_bind(b200gp_ctx* ctx, void* d_workspace, size_t workspace_bytes, int slots, int max_params,
                const double* d_X, int N, int D, const double* d_y);

/* kern.K(X) for the n items ids[0..n) over the bound X: d_K_out is n x N x N (symmetric, dense).
 * add_diag = 0: the bare kernel matrix (compute_kernel, backend/kernel.py:317-322);
 * add_diag = 1: K_y = K + (noise + 1e-8) I as ExactGaussianInference forms it. */
int b200gp_build_K(b200gp_ctx* ctx, const int32_t* d_prog, const int32_t* d_prog_off, const double* d_theta,
                   const int32_t* d_theta_off, const int32_t* d_ids, int n, int add_diag, double* d_K_out);

/* Log marginal likelihood and its gradient w.r.t. the constrained parameters [theta_kernel..., noise]
 * for the n items ids[0..n) (GP.log_likelihood() and the `.gradient` fields GP.parameters_changed fills).
 * Outputs are indexed by ITEM id: d_lml[item], d_dlml[theta_off[item] ...], d_status[item],
 * d_tries[item] (jitter-ladder tries, 0 = plain factorisation; may be NULL).
 * with_grad = 0 skips K^-1 and the gradient pass.  Synchronises the stream. */
int b200gp_lml_grad(b200gp_ctx* ctx, const int32_t* d_prog, const int32_t* d_prog_off, const double* d_theta,
                    const int32_t* d_theta_off, const int32_t* d_ids, int n, int with_grad, double* d_lml,
                    double* d_dlml, int32_t* d_status, int32_t* d_tries);

/* GP.predict for ONE item: posterior mean and marginal variance at the Ns test inputs d_Xs (Ns x D
 * row-major, same standardisation as the bound X), GPy semantics (Posterior._raw_predict + Gaussian.predictive_values;
 * callers: ModelSelector.predict, src/autoks/core/model_selection/base.py:231-237; KernelKernelGPModel._predict /
 * get_f_max, src/autoks/gp_regression_models.py:113-134):
 *     mean = k(x*, X) K_y^-1 y,   var = clip(k(x*, x*) - k(x*, X) K_y^-1 k(X, x*), 1e-15, inf) [+ noise variance]
 * d_id points at the item id on the device (ids[0..1) of b200gp_lml_grad), `item` is the same id on the host.  The
 * factorisation runs through the same pipeline and jitter ladder as b200gp_lml_grad; d_lml[item] / d_status[item]
 * receive its log marginal likelihood and status (NOT_PD: mean and var are NaN).  Synchronises the stream. */
int b200gp_posterior(b200gp_ctx* ctx, const int32_t* d_prog, const int32_t* d_prog_off, const double* d_theta,
                     const int32_t* d_theta_off, const int32_t* d_id, int item, const double* d_Xs, int Ns,
                     int include_likelihood, double* d_mean, double* d_var, double* d_lml, int32_t* d_status);

/* In-place blocked Cholesky of ONE dense SPD matrix (lower triangle; N a multiple of 128), the
 * `dpotrf` of GPy's jitchol at large N.  d_work needs b200gp_potrf_work_bytes(N) bytes.
 * logdet / status are HOST outputs (status 0 ok, >0 = 1-based index of the first bad pivot). */
size_t b200gp_potrf_work_bytes(int N);
int b200gp_potrf(b200gp_ctx* ctx, double* d_A, int N, void* d_work, double* logdet, int32_t* status);

/* ---- BOMS Hellinger kernel-kernel distance (src/autoks/distance/distance.py:126-194) ------------------------
 * precompute: for model m (program m, P_m kernel parameters) and sample s: G = K(theta[m][s]) + lambda[s] I over the
 *   n-point subset d_Xsub (n x D row-major, n <= 128), d_logdet[m*S+s] = log det G, optional d_G[m][s] = G (n x n dense).
 *   d_theta_samples is packed [model][sample][param]: model m starts at theta_off[m]*S doubles (theta_off counts
 *   kernel parameters only -- there is no noise entry here; lambda plays that role).  status 1 = not PD.
 *   (create_precomputed_info, distance.py:170-194)
 * pairs: d_dist[p] = Hellinger distance of models (pair_i[p], pair_j[p]); samples whose logdets differ by <= tol_logdet
 *   re-use logdet_i (distance.py:145-148).  status 1 = a 1/2(G_i+G_j) was not PD (the reference would call
 *   cov_nearest there); d_work needs b200gp_hellinger_pairs_work_bytes(npairs, S).
 *   (hellinger_distance + compute_distance, distance.py:135-168,110-118) */
int b200gp_hellinger_precompute(b200gp_ctx* ctx, const double* d_Xsub, int n, int D, const int32_t* d_prog,
                                const int32_t* d_prog_off, const double* d_theta_samples, const int32_t* d_theta_off,
                                const double* d_lambda, int M, int S, double* d_logdet, double* d_G, int32_t* d_status);
size_t b200gp_hellinger_pairs_work_bytes(int npairs, int S);
int b200gp_hellinger_pairs(b200gp_ctx* ctx, const double* d_logdet, const double* d_G, int M, int S, int n,
                           const int32_t* d_pair_i, const int32_t* d_pair_j, int npairs, double tol_logdet, void* d_work,
                           double* d_dist, int32_t* d_status);

/* Centred kernel alignment of every pair of the n items ids[0..n) over the bound X (SURVEY.md 8f row f2;
 * pairwise_centered_alignments -> centered_alignment -> center_kernel / alignment / inner_frob,
 * src/autoks/core/covariance.py:278-338): d_out[a * n + b] = <HK_aH, HK_bH>_F / sqrt(<HK_aH, HK_aH>_F <HK_bH, HK_bH>_F),
 * K = kern.K(X) without noise, H = I - 11^T / N.  Needs n <= bound slots (all n kernel matrices are resident);
 * theta has the b200gp_lml_grad layout (the noise entries are ignored). */
int b200gp_alignment_min_slots(int N, int n);      /* slots to bind for n kernels of order N (>= n: scratch re-uses slots) */
int b200gp_centered_alignments(b200gp_ctx* ctx, const int32_t* d_prog, const int32_t* d_prog_off, const double* d_theta,
                               const int32_t* d_theta_off, const int32_t* d_ids, int n, double* d_out);

/* Frobenius (metric 0) / correlation (metric 1) distance between the sampled prior Gram matrices of model pairs
 * (FrobeniusDistanceBuilder.frobenius_distance / CorrelationDistanceBuilder.correlation_distance,
 * src/autoks/distance/distance.py:208-212,242-265).  d_G is the [M][S][n][n] stack b200gp_hellinger_precompute wrote
 * (the reference's vec(K_s + lambda_s I) columns, which it stores in float32: entries are rounded to float32 here too). */
int b200gp_vector_pairs(b200gp_ctx* ctx, const double* d_G, int M, int S, int n, const int32_t* d_pair_i,
                        const int32_t* d_pair_j, int npairs, int metric, double* d_dist);

/* Gram matrices and their log-determinants for `n` items over the bound X, through the scoring path's workspace slots
 * (fused K-build + blocked DMMA Cholesky), for any N: what the distance builders need when the reference hands them
 * more than B200GP_TILE points (it passes the whole training set: smbo_model_selector.py:91-97).
 *   diag_mode 0: K;  1: K + (noise + 1e-8) I (exact inference);  2: K + noise I (Hellinger Gram matrices,
 *   distance.py:186-190; the item's last theta entry is lambda_s)
 *   d_K_out [n][N][N] symmetric copies or NULL;  d_logdet [items] / d_status [items] (1 = not positive definite) or NULL */
int b200gp_gram_logdet(b200gp_ctx* ctx, const int32_t* d_prog, const int32_t* d_prog_off, const double* d_theta,
                       const int32_t* d_theta_off, const int32_t* d_ids, int n, int diag_mode, double* d_K_out,
                       double* d_logdet, int32_t* d_status);
/* b200gp_hellinger_pairs for Gram matrices of any order n (= the N the context is bound to).  The caller lists the
 * (pair, sample) cells whose log-determinants differ by more than the tolerance (hellinger_distance refactorises only
 * those, distance.py:145-153) as d_work_pair[nwork] / d_work_s[nwork]; each costs one blocked Cholesky of
 * 1/2 (G_i,s + G_j,s) in a workspace slot.  d_work: b200gp_hellinger_pairs_large_work_bytes(npairs, S) bytes. */
size_t b200gp_hellinger_pairs_large_work_bytes(int npairs, int S);
int b200gp_hellinger_pairs_large(b200gp_ctx* ctx, const double* d_logdet, const double* d_G, int M, int S, int n,
                                 const int32_t* d_pair_i, const int32_t* d_pair_j, int npairs, const int32_t* d_work_pair,
                                 const int32_t* d_work_s, int nwork, void* d_work, double* d_dist, int32_t* d_status);

/* Phase timings (ms) of the most recent b200gp_lml_grad chunk, measured with CUDA events on the context
 * stream: [build_K, cholesky, inverse(V), solve, Kinv, grad+finalize].  Returns the number written. */
int b200gp_last_phase_ms(b200gp_ctx* ctx, float* out, int cap);
int b200gp_set_profiling(b200gp_ctx* ctx, int enabled);
/* panel width (in 128-column tiles) of the blocked Cholesky / L^-T schedules; the default is chosen per batch
 * (1-2 tile columns for latency-bound small batches, 4 for a single large matrix, 8 for large batches) */
int b200gp_set_panel_tiles(b200gp_ctx* ctx, int tiles);
/* CTA tile shape of the DMMA phases.  1 (default): 64 x 64 tiles, three 4-warp CTAs per SM, structural zeros skipped per
 * CTA; 0: 128 x 64 tiles, two 8-warp CTAs per SM, structural zeros skipped per warp.  The results are bit-identical
 * (tests/test_gpu_fullsize.py); the 64 x 64 form is 3-5 % faster on every phase. */
int b200gp_set_tile64(b200gp_ctx* ctx, int enabled);
/* Diagonal 128 x 128 tiles (Cholesky factor + triangular inverse, the latency-critical step of every blocked phase).
 * 1 (default): look-ahead register factorisation of [A | I], one column per block barrier (csrc/tile_potrf_la.cuh); 2: the
 * same with a split barrier (mbarrier arrive after the early publish, wait at the next step: bit-identical, measured 0.7 %
 * faster on a lone evaluation); 0: blocked over 16-column panels (csrc/tile_potrf_blocked.cuh).  Same results up to rounding
 * between 0 and 1 / 2 (tests/test_gpu_potrf.py runs all three). */
int b200gp_set_potrf_la(b200gp_ctx* ctx, int mode);
/* How the 64 x 64 tile GEMMs stage their operands.  3 (default): TMA (`cp.async.bulk.tensor.2d` into an mbarrier-guarded
 * ring, csrc/tile_gemm64_tma.cuh) with tensor maps swizzled SWIZZLE_128B_ATOM_32B; 2: TMA, SWIZZLE_128B; 1: TMA, no swizzle;
 * 0: cp.async (LDGSTS) with a block barrier per K chunk; 4: as 3 with a 3-stage ring and four CTAs per SM (measured equal).
 * Bit-identical results in every mode.  When the driver cannot encode the tensor maps the library stays on the cp.async
 * kernels. */
int b200gp_set_tma(b200gp_ctx* ctx, int mode);
/* 1: K^-1 and the gradient pass of the <= 4-leaf programs run as ONE kernel (K^-1 never reaches HBM); 0 (default): two
 * kernels.  Same results up to the order of the partial sums; measured +0.8 % throughput at N = 2048. */
int b200gp_set_fused_grad(b200gp_ctx* ctx, int enabled);
/* kernels launched by this context since creation (for bench.py's gpu_launches) */
long long b200gp_launch_count(b200gp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200GP_H */
