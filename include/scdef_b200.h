/* scdef_b200 — C ABI of the B200-native scDEF training hot path.
 *
 * The reference (cbg-ethz/scDEF) has no FFI: its seam is the pair of jitted Python
 * callables built in `_get_or_build_learn_grad_fns`
 * (src/scdef/models/_scdef.py:2808-2856) and consumed by `local_update` / `global_update`
 * (3087-3163).  This header is what a maintainer would bind (ctypes) to replace them:
 *
 *   scdef_elbo_grad   <->  local_loss_grad / global_loss_grad   (_scdef.py:2848-2849;
 *                          objective 2822-2846, batch_elbo 2148-2181, elbo 1823-2146,
 *                          jax_utils.py:6-46)
 *   scdef_adam_local  <->  optax.adam(local_lr) update + apply over local_params
 *                          (_scdef.py:3079, 3112-3116) + frozen-z snap-back (3248-3255)
 *   scdef_adam_global <->  optax.adam(lr) update + apply + freeze_w restore + clip_params
 *                          (_scdef.py:3080, 3149-3162, 3034-3042)
 *   scdef_posterior   <->  set_posterior_means / set_posterior_variances (3394-3476)
 *   scdef_row_stats   <->  the per-cell constants of load_adata (482-504) that the hot
 *                          path needs (library size, sum_g lgamma(x+1))
 *
 * Conventions: every pointer marked "device" is caller-owned contiguous CUDA memory
 * (PyTorch `data_ptr()`), 16-byte aligned.  All calls enqueue on the caller's stream and
 * return immediately; there are no hidden host syncs, no cudaMalloc, no CPU fallback.
 * Return value 0 = success, negative = SCDEF_E*.  Not thread-safe per ctx.
 * Variational blocks are stored as the reference stores them: a stacked pair
 * [mu, rho = log sigma] along axis 0, float32.
 */
#ifndef SCDEF_B200_H
#define SCDEF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCDEF_ABI_VERSION 8
#define SCDEF_MAX_LAYERS 8

#define SCDEF_OK 0
#define SCDEF_EINVAL -1     /* bad argument / shape */
#define SCDEF_EALIGN -2     /* pointer not 16-byte aligned */
#define SCDEF_EUNSUPPORTED -3 /* shape outside what the kernels support */
#define SCDEF_EWORKSPACE -4 /* workspace too small */
#define SCDEF_ECUDA -5      /* CUDA launch / runtime error (see scdef_last_error) */

#define SCDEF_PASS_LOCAL 1  /* gradients w.r.t. local_params  (argnums=2) */
#define SCDEF_PASS_GLOBAL 2 /* gradients w.r.t. global_params (argnums=3) */

#define SCDEF_NOISE_PHILOX 0   /* counter-based N(0,1), generated inside the kernels */
#define SCDEF_NOISE_EXPLICIT 1 /* caller-provided draws (parity tests) */

#define SCDEF_LIK_AUTO 0
#define SCDEF_LIK_SIMT 1 /* fp32 CUDA-core likelihood kernel (any shape)               */
#define SCDEF_LIK_TC 2   /* tcgen05 / TMEM split-bf16 likelihood kernel (K0 <= 112, G % 8 == 0, nb <= 8) */

typedef struct scdef_ctx scdef_ctx;

/* Static description of the model (what `elbo` reads from `self`). */
typedef struct scdef_model_desc {
  int32_t abi_version;      /* = SCDEF_ABI_VERSION */
  int32_t n_layers;         /* L */
  int32_t layer_sizes[SCDEF_MAX_LAYERS]; /* K_0 (bottom) ... K_{L-1} */
  int64_t n_cells;          /* rows of the local tensors held by THIS rank */
  int32_t n_genes;          /* G */
  int32_t n_batches;        /* nb >= 1 */
  int32_t use_brd;          /* _scdef.py:1936, 1982, 2014 */
  int32_t marginalize_alpha;/* _scdef.py:1972 */
  int32_t supervised_layer; /* sscDEF pinned layer (_sscdef.py:484) or -1 */
  float alpha_floor;        /* 0.1 (_scdef.py:2068) / 1.0 (_sscdef.py:506) */
  float top_alpha, brd, brd_mean, shrinkage_shape, shrinkage_rate;
  float cell_scale_shape, gene_scale_shape;
  /* device constants */
  const float* X;                 /* (N,G) row-major counts, resident; NULL if batches are streamed */
  const int32_t* batch_id;        /* (N) batch of each cell (argmax of batch_indices_onehot) */
  const float* batch_lib_ratio;   /* (N)      _scdef.py:500-504 */
  const float* cell_alpha_factor; /* (N)      _scdef.py:1279-1281 */
  const float* gene_ratio;        /* (nb,G)   _scdef.py:521-524, 568-573 (broadcast if scalar) */
  const float* x_lgamma_rowsum;   /* (N) sum_g lgamma(x+1) or NULL (computed per batch) */
  int32_t* row_slot;              /* (N) scratch owned by the library between calls; caller fills -1 once */
  /* W priors, w_priors[l] = [shape, rate] (_scdef.py:1258-1277; _iscdef.py:442-450):
     matrix (K_l, K_{l-1}) device pointers, or NULL -> the scalar is used */
  const float* w_prior_shape[SCDEF_MAX_LAYERS];
  const float* w_prior_rate[SCDEF_MAX_LAYERS];
  float w_prior_shape_scalar[SCDEF_MAX_LAYERS];
  float w_prior_rate_scalar[SCDEF_MAX_LAYERS];
  const float* fixed_top_z;       /* sscDEF (N, K_t) constant z of the pinned layer, or NULL */
} scdef_model_desc;

/* One set of per-block device pointers: parameters, gradients or Adam moments.
   Layouts: cell_scale (2,R,1), z (2,R,sumK), gene_scale (2,nb,G), brd (2,K0,1),
   W[l] (2,K_l,K_{l-1}) with K_{-1}=G, ard (2,K0,1), alpha (2,1,1).
   R = n_cells for parameters/moments; R = B (minibatch slot order) for local gradients. */
typedef struct scdef_blocks {
  float* cell_scale;
  float* z;
  float* gene_scale;
  float* brd;
  float* W[SCDEF_MAX_LAYERS];
  float* ard;
  float* alpha;
} scdef_blocks;

/* Noise of one `elbo_grad` call (= one `rng_input` of the reference, _scdef.py:3231). */
typedef struct scdef_noise {
  int32_t mode;     /* SCDEF_NOISE_* */
  int32_t coupling; /* 1: blocks of identical shape share their draws within an MC sample, as
                       the reference's single-key sampling does; 0: independent */
  uint64_t seed;    /* philox key */
  uint64_t step;    /* philox key, the per-iteration split */
  /* explicit draws, leading axis S (device): */
  const float* eps_cell_scale; /* (S,B,1)    */
  const float* eps_z;          /* (S,B,sumK) */
  const float* eps_gene_scale; /* (S,nb,G)   */
  const float* eps_brd;        /* (S,K0,1)   */
  const float* eps_W[SCDEF_MAX_LAYERS]; /* (S,K_l,K_{l-1}) */
  const float* eps_ard;        /* (S,K0,1)   */
  const float* eps_alpha;      /* (S,1,1)    */
  /* Data parallelism with position-based ownership (philox mode, per-cell blocks only): this rank's
     B rows are rows [local_row_offset, local_row_offset+B) of a global minibatch of local_rows_total
     rows, and draw exactly the normals the single-rank run draws for those rows.  0/0: B rows at 0. */
  int64_t local_row_offset;
  int64_t local_rows_total;
} scdef_noise;

/* Per-call arguments of the objective (the non-array arguments of `objective`). */
typedef struct scdef_call {
  int32_t pass;          /* SCDEF_PASS_LOCAL | SCDEF_PASS_GLOBAL */
  int32_t num_samples;   /* S */
  int32_t batch;         /* B = number of minibatch rows on this rank (may be 0) */
  int32_t lik_impl;      /* SCDEF_LIK_* */
  float annealing;       /* annealing_parameter */
  float stop_gradients[SCDEF_MAX_LAYERS]; /* 0/1 per layer (_scdef.py:3064-3066) */
  float stop_cell_budgets, stop_gene_budgets;
  float alpha;
  float scale;           /* N_total / B_total  (_scdef.py:2143); under data parallelism the
                            GLOBAL cell and minibatch counts */
  float global_weight;   /* weight of the cell-independent terms (priors/entropies of the global
                            blocks) in loss and global gradients: 1 on one GPU, 1/R on R ranks so
                            the all-reduce sum restores them */
  const int64_t* indices;/* device (B) rows of this rank's local tensors */
  const float* X_batch;  /* device (B,G) counts of the minibatch, or NULL -> gathered from desc.X */
  int32_t flags;         /* SCDEF_FLAG_* */
  int32_t reserved;
} scdef_call;

/* The previous scdef_elbo_grad on this ctx used the same workspace, noise, num_samples, lik_impl and
   W_0 / gene-scale parameters (the LOCAL pass of the same hot-loop iteration, `_scdef.py:3238,3260`:
   same rng_input, globals not yet updated): its sampled W_0 tile images may be reused. */
#define SCDEF_FLAG_REUSE_W0_SAMPLES 1
/* The caller ignores loss_out of this pass (the reference overwrites the local-pass loss with the
   global-pass one, `_scdef.py:3235-3271`): kernels may skip value-only work. */
#define SCDEF_FLAG_SKIP_LOSS 2
/* GLOBAL pass only: run the likelihood group before the hierarchy kernel (their outputs are disjoint), so that the W_0
   gradient — almost all of the buffer a data-parallel caller all-reduces — is final while the rest of the pass still
   runs.  Results are the same as without the flag up to the order of the float64 loss atomics. */
#define SCDEF_FLAG_LIK_FIRST 4

/* Data parallelism: `cuda_event` (a cudaEvent_t, or NULL to clear) is recorded on the call's stream by every
   scdef_elbo_grad(GLOBAL) at the point where grads->W[0] is final (tcgen05 path: before the hierarchy kernel when
   SCDEF_FLAG_LIK_FIRST is set and before k_finish otherwise; CUDA-core path: at the end of the pass), so the caller can
   start reducing it on another stream while the rest of the pass runs.  No counterpart in the reference (single device). */
int scdef_set_w0_grad_event(scdef_ctx* ctx, void* cuda_event);

int scdef_abi_version(void);
int scdef_ctx_create(const scdef_model_desc* desc, scdef_ctx** out);
void scdef_ctx_destroy(scdef_ctx* ctx);
const char* scdef_last_error(scdef_ctx* ctx);

/* Scratch needed by scdef_elbo_grad for minibatches up to B_max rows and S samples. */
size_t scdef_workspace_bytes(scdef_ctx* ctx, int B_max, int S, int lik_impl);

/* loss = -(1/S) sum_s ELBO_s and its gradient w.r.t. the local or the global blocks.
   `loss_out` (device double[2]): [0] += cell-dependent part, [1] += global part * global_weight;
   the caller zeroes it.  `grads`: for PASS_LOCAL only cell_scale (2,B,1) and z (2,B,sumK) are
   written (minibatch slot order; every other row of the reference's dense gradient is 0);
   for PASS_GLOBAL all global blocks are written (overwritten, not accumulated). */
int scdef_elbo_grad(scdef_ctx* ctx, const scdef_call* call, const scdef_blocks* params,
                    const scdef_noise* noise, double* loss_out, const scdef_blocks* grads,
                    void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Dense Adam over ALL n_cells rows of the local blocks (optax semantics: rows outside the
   minibatch see a zero gradient but their moments decay and they still move).  `grads` is the
   compact (2,B,.) output of scdef_elbo_grad, `indices` the same minibatch rows.
   `freeze_z_layers[l] != 0` keeps the z columns of layer l unchanged (_scdef.py:3248-3255).
   Locals are never clipped (clip_params skips a 2-element list, _scdef.py:3038). */
int scdef_adam_local(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* grads,
                     const scdef_blocks* m, const scdef_blocks* v, const int64_t* indices,
                     int B, int64_t step, float lr, float b1, float b2, float eps,
                     const int32_t* freeze_z_layers, void* cuda_stream);

/* Exact lazy form of scdef_adam_local (bit-identical results, see adam.cuh): rows outside the minibatch are
   not streamed every step; their zero-gradient Adam steps are replayed when the row is next needed.
     phase SCDEF_LAZY_CATCHUP     : replay steps row_step[n]+1 .. step-1 for the rows `indices`
                                    (call BEFORE the local pass of iteration `step`)
     phase SCDEF_LAZY_UPDATE      : apply iteration `step` with `grads` to the rows `indices` (call after it)
     phase SCDEF_LAZY_CATCHUP_ALL : replay up to and including `step` for every row (before the parameters are
                                    read, the optimiser is reset, or lr / betas change)
   `row_step` (device int32[n_cells], caller zero-fills when the optimiser state is reset) and `bc_table`
   (device float2[bc_len]: (1/(1-b1^k), 1/(1-b2^k)) for k = 1..bc_len, the last entry is used beyond) are
   caller-owned. */
/* Fills host_out[2*(k-1)..] = (1/(1-b1^k), 1/(1-b2^k)) for k = 1.. until both are exactly 1 (or max_len);
   returns the number of entries.  Same arithmetic as the dense kernel's per-step factors. */
int scdef_adam_bias_table(float b1, float b2, float* host_out, int max_len);
#define SCDEF_LAZY_CATCHUP 0
#define SCDEF_LAZY_UPDATE 1
#define SCDEF_LAZY_CATCHUP_ALL 2
int scdef_adam_local_lazy(scdef_ctx* ctx, int phase, const scdef_blocks* params, const scdef_blocks* grads,
                          const scdef_blocks* m, const scdef_blocks* v, const int64_t* indices, int B,
                          int64_t step, float lr, float b1, float b2, float eps,
                          const int32_t* freeze_z_layers, int32_t* row_step, const float* bc_table,
                          int bc_len, void* cuda_stream);

/* Adam over the global blocks, then freeze_w restore (W blocks keep their values, moments
   still update) and clip_params: mu <= 1e4, rho <= 10 on gene_scale, brd and every W
   (not ard, not alpha). */
int scdef_adam_global(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* grads,
                      const scdef_blocks* m, const scdef_blocks* v, int64_t step, float lr,
                      float b1, float b2, float eps, int freeze_w, void* cuda_stream);

/* E[v] = exp(mu + sigma^2/2) and Var[v] of every block, on the device (`means`/`vars` have the
   shape of one half of each block; either may be NULL). */
int scdef_posterior(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* means,
                    const scdef_blocks* vars, void* cuda_stream);

/* Hard assignment counts per factor (`filter_factors`, _scdef.py:3540-3549, and `get_effective_factors`,
   1378-1385): for every layer l, counts[off_l + k] = #{cells n : argmax_j E[z_l][n, j] = k} with
   E[z] = exp(mu + sigma^2/2) evaluated exactly as scdef_posterior does (first maximum on ties, like
   np.argmax).  counts: device int64[sum K], overwritten.  Reads the z block of `params` only. */
int scdef_assignment_counts(scdef_ctx* ctx, const scdef_blocks* params, int64_t* counts, void* cuda_stream);

/* Per-cell constants from a resident count matrix: lib[n] = sum_g x, lgam[n] = sum_g lgamma(x+1). */
int scdef_row_stats(const float* X, int64_t n_rows, int32_t n_genes, float* lib, float* lgam,
                    void* cuda_stream);

/* Counts kept as uint16 on the device (raw UMI counts fit; half the HBM of float32): the per-cell constants, and the
   minibatch rows converted to the dense float32 `X_batch` of scdef_elbo_grad. */
int scdef_u16_row_stats(const uint16_t* X, int64_t n_rows, int32_t n_genes, float* lib, float* lgam, void* cuda_stream);
int scdef_u16_gather_rows(const uint16_t* X, const int64_t* rows, int32_t n_rows, int32_t n_genes, float* out,
                          void* cuda_stream);

/* The count statistics behind the constants of `load_adata` (src/scdef/models/_scdef.py:498-573), on the device:
   batch_stats[(b, {n cells, sum lib, sum lib^2})] for b = 0..n_batches-1 and, in slot n_batches, over all cells;
   gene_sums[(b, g)] likewise ((n_batches + 1) x n_genes); lib[n] (optional) = library size of every cell.
   X: float32 or uint16 (x_is_u16), row-major n_rows x n_genes; batch_id may be NULL when n_batches == 1.
   Outputs are device doubles, overwritten; a data-parallel caller all-reduces them over the ranks. */
int scdef_count_stats(const void* X, int32_t x_is_u16, const int32_t* batch_id, int64_t n_rows, int32_t n_genes,
                      int32_t n_batches, double* batch_stats, double* gene_sums, float* lib, void* cuda_stream);

/* Counts kept as CSR on the device (float32 data, int32 column indices, int64 row pointers; the reference
   densifies `adata.X`, _scdef.py:491-496 — 40 GB of host memory at 1M x 5k — this keeps ~10 % of that in HBM).
   scdef_csr_row_stats: the per-cell constants of scdef_row_stats from the CSR rows.
   scdef_csr_gather_rows: out[j, :] = dense row `rows[j]` (out: n_rows x n_genes floats, overwritten); the
   result is passed to scdef_elbo_grad as `X_batch`, so every kernel downstream is the dense path. */
int scdef_csr_row_stats(const int64_t* indptr, const float* data, int64_t n_rows, float* lib, float* lgam,
                        void* cuda_stream);
int scdef_csr_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                          int32_t n_rows, int32_t n_genes, float* out, void* cuda_stream);

/* Post-fit Monte-Carlo assignment statistics of ONE layer (`scdef.tl.assign_confident`, tools/factor.py:1264-1379;
   the reference does this with two jax.lax.scan passes over n_samples keys).  For each of the n_rows cells:
   n_samples draws z = exp(mu + exp(rho) eps) of the n_cols kept factors (element (n, k) of the z block at
   mu[n*ld + cols[k]], rho likewise; cols: device int32[n_cols]), normalised to zhat; winner = argmax_k mean_s zhat;
   gap_s = zhat_s[winner] - max_{k != winner} zhat_s[k].
   stats (n_rows x 6, overwritten): clip(mean gap, -1, 1), clip(SD of the gap, 0, 1), clip(quantile_level-quantile of
   the gap with linear interpolation, 0, 1), clip(mean zhat[winner], 0, 1), max_k P(argmax = k), clip(1 - H(P)/log n_cols, 0, 1);
   argmax (n_rows, overwritten): winner slot in [0, n_cols).
   Noise: eps != NULL -> explicit draws (n_samples, n_rows, n_cols) (parity tests); otherwise Philox4x32-10 keyed by
   (seed, stream_id, row0 + n, k, sample), so a cell's draws do not depend on how the cells are sharded.
   n_cols in [2, 1024] (single-factor layers are constants, factor.py:1270-1281); 32*n_samples bytes of shared memory. */
int scdef_assign_confident(const float* mu, const float* rho, int64_t n_rows, int32_t ld, const int32_t* cols,
                           int32_t n_cols, int32_t n_samples, float quantile_level, uint64_t seed, int32_t stream_id,
                           int64_t row0, const float* eps, float* stats, int32_t* argmax, void* cuda_stream);

/* Monte-Carlo gene scores behind `tl.set_confident_signatures` (src/scdef/tools/factor.py:220-363; layer 0:
   models/_scdef.py:3975-3989), reduced on the fly instead of materialising the (draws, factors, genes) tensor of
   `_collect_hierarchy_mc_scores` (factor.py:271-291):
     score^s[f][g] = sum_r path[s][f][r] * clip(exp(w_mu[rows[r]][g] + exp(w_rho[rows[r]][g]) * eps[s][r][g]), 1e-15, 1e15)
   (path == NULL: identity, n_factors == n_rows — every factor's own draw).  w_mu / w_rho: the two halves of the W_0 block,
   (K0, n_genes) row-major; eps: (n_draws, n_rows, n_genes) N(0,1) draws supplied by the caller; path: (n_draws, n_factors, n_rows).
     mode 0: counts[f][g] += [score > tau[f]]                               (int32 (n_factors, n_genes), zeroed by the caller)
     mode 1: ref_score[s][f][i] = score at gene ref_idx[f][i]                (ref_idx (n_factors, kmax) int32, -1 = unused)
     mode 2: rank_count[s][f][i] += #{g : score^s[f][g] > ref_score[s][f][i]} (int32, zeroed by the caller; run mode 1 first)
   n_rows <= 128, n_factors <= 256. */
int scdef_signature_scores(const float* w_mu, const float* w_rho, int32_t n_genes, const int32_t* rows, int32_t n_rows,
                           const float* eps, const float* path, int32_t n_draws, int32_t n_factors, int32_t mode, const float* tau,
                           int32_t* counts, const int32_t* ref_idx, int32_t kmax, float* ref_score, int32_t* rank_count,
                           void* cuda_stream);

/* Cold start of a LogNormal block from Gamma draws, on the device (`init_var_params`, _scdef.py:1639-1648 for the z layers):
   m = clip(Gamma(shape, 1) / shape, clip_lo, clip_hi), v = m / var_div, (mu, rho = log sigma) = the LogNormal with mean m and
   variance v.  Element (n, k) is written at mu[n*ld + k] / rho[n*ld + k].  The reference draws with JAX keys, so parity is
   distributional; the draws here are Philox4x32-10 keyed by (seed, stream_id, row0 + n, k). */
int scdef_init_gamma_block(float* mu, float* rho, int64_t n_rows, int32_t n_cols, int32_t ld, float shape, float clip_lo,
                           float clip_hi, float var_div, uint64_t seed, int32_t stream_id, int64_t row0, void* cuda_stream);

/* Writes the N(0,1) draws the kernels would use for one block in PHILOX mode (debug / parity):
   block_kind 0 cell_scale, 1 z layer `layer`, 2 gene_scale, 3 brd, 4 W layer `layer`, 5 ard,
   6 alpha; out is (S, rows, cols). */
int scdef_noise_fill(scdef_ctx* ctx, const scdef_noise* noise, int block_kind, int layer, int S,
                     int B, float* out, void* cuda_stream);

/* Measurement (bench.py): CUDA-event pairs recorded on the launch stream around every launch
   of each kernel family, and a count of all kernels launched through this ABI. */
#define SCDEF_PROF_LIK 0          /* likelihood kernel (K2-K4) */
#define SCDEF_PROF_ADAM_LOCAL 1   /* dense Adam over the z block */
#define SCDEF_PROF_HIER 2
#define SCDEF_PROF_SAMPLE 3
#define SCDEF_PROF_WPRIOR 4
#define SCDEF_PROF_FINISH 5
#define SCDEF_PROF_ADAM_GLOBAL 6
#define SCDEF_PROF_PRIORS 7
#define SCDEF_PROF_ADAM_LOCAL_C 8 /* dense Adam over the cell-scale block */
#define SCDEF_PROF_LIK_MAIN 9     /* the fused likelihood kernel alone (without its pre-/post-kernels), LOCAL pass */
#define SCDEF_PROF_LIK_MAIN_GLOBAL 10 /* the same, GLOBAL pass */
#define SCDEF_PROF_KINDS 11
/* on: 0 = off, 1 = every kind, otherwise a mask (bit k+1 = kind k) so that a timed region can carry the
   event pairs of its dominant kernel only (two records per launch are not free on a ~30-launch step). */
int scdef_profile_enable(scdef_ctx* ctx, int on);
/* Sums the elapsed ms of the (up to 256 most recent) timed launches of each kind; syncs on them. */
int scdef_profile_read(scdef_ctx* ctx, double* ms_sum, int64_t* n_timed, int64_t* n_launched,
                       int64_t* total_launches, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SCDEF_B200_H */
