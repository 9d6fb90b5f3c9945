// C ABI of scdef_b200 (include/scdef_b200.h): host-side orchestration of the kernels.
// Every entry point only enqueues work on the caller's stream.
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>

#include "common.cuh"
#include "kernels_small.cuh"
#include "lik_simt.cuh"
#include "adam.cuh"
#include "lik_tc.cuh"
#include "tools.cuh"

using namespace scdef;

constexpr int PROF_RING = 256;

struct scdef_ctx {
  cudaEvent_t w0_event = nullptr;   // scdef_set_w0_grad_event: recorded in the GLOBAL pass once the W_0 gradient is final
  scdef_model_desc d;
  int L, Kt, K0;
  int K[SCDEF_MAX_LAYERS], off[SCDEF_MAX_LAYERS + 1], out_dim[SCDEF_MAX_LAYERS];
  int sm_count;
  char err[512];
  // profiling: CUDA-event pairs around every launch of each kernel family, on the launch stream
  bool prof_on, prof_ready;
  unsigned prof_mask;  // bit k: kind k is timed
  cudaEvent_t prof_ev[SCDEF_PROF_KINDS][PROF_RING][2];
  long long prof_count[SCDEF_PROF_KINDS];
  long long launches;
};

namespace {

enum { BLK_C = 0, BLK_S = 1, BLK_TAU = 2, BLK_OM = 3, BLK_A = 4, BLK_W = 5, BLK_Z = 5 + SCDEF_MAX_LAYERS };

// k_sample implementation: 2 = k_sample_v2 (default since round 2: 0.119 -> 0.057 ms per step at C5, bit-identical
// samples, profiles/r02_first_call_summary.txt), 1 = the first kernel (SCDEF_SAMPLE_V2=0, or the debug setter; kept as
// the cross-check of tests/test_variants_gpu.py)
int g_sample_impl = 0;
int sample_impl() {
  if (g_sample_impl == 0) {
    const char* e = getenv("SCDEF_SAMPLE_V2");
    g_sample_impl = (e && atoi(e) == 0) ? 1 : 2;
  }
  return g_sample_impl;
}

int fail(scdef_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

int check_launch(scdef_ctx* ctx, const char* what) {
  if (ctx) ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, SCDEF_ECUDA, "%s: %s", what, cudaGetErrorString(e));
  return SCDEF_OK;
}

struct ProfScope {
  scdef_ctx* c; int kind; cudaStream_t st; int slot;
  ProfScope(scdef_ctx* c_, int kind_, cudaStream_t st_) : c(c_), kind(kind_), st(st_), slot(-1) {
    if (c->prof_on && c->prof_ready && ((c->prof_mask >> kind) & 1u)) {
      slot = (int)(c->prof_count[kind] % PROF_RING);
      cudaEventRecord(c->prof_ev[kind][slot][0], st);
    }
  }
  ~ProfScope() {
    if (slot >= 0) {
      cudaEventRecord(c->prof_ev[kind][slot][1], st);
      c->prof_count[kind]++;
    }
  }
};

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Workspace layout: per block a sample buffer and a dELBO/dv buffer, (S, rows, cols) floats.
struct Plan {
  size_t smp[MAX_BLOCKS], G[MAX_BLOCKS];
  int rows[MAX_BLOCKS], cols[MAX_BLOCKS];
  size_t tc_off, tc_bytes;  // scratch of the tcgen05 likelihood kernel
  size_t total;
};

Plan make_plan(const scdef_ctx* c, int B, int S, int lik_impl) {
  Plan p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < MAX_BLOCKS; ++i) p.rows[i] = p.cols[i] = 0;
  p.rows[BLK_C] = B; p.cols[BLK_C] = 1;
  p.rows[BLK_S] = c->d.n_batches; p.cols[BLK_S] = c->d.n_genes;
  p.rows[BLK_TAU] = c->K0; p.cols[BLK_TAU] = 1;
  p.rows[BLK_OM] = c->K0; p.cols[BLK_OM] = 1;
  p.rows[BLK_A] = 1; p.cols[BLK_A] = 1;
  for (int l = 0; l < c->L; ++l) {
    p.rows[BLK_W + l] = c->K[l]; p.cols[BLK_W + l] = c->out_dim[l];
    p.rows[BLK_Z + l] = B; p.cols[BLK_Z + l] = c->K[l];
  }
  size_t o = 0;
  for (int i = 0; i < MAX_BLOCKS; ++i) {
    size_t n = (size_t)S * p.rows[i] * p.cols[i] * sizeof(float);
    if (i == BLK_W && lik_impl == SCDEF_LIK_TC) n = 0;  // W_0 samples never hit HBM
    p.smp[i] = o; o += align_up(n);
    p.G[i] = o; o += align_up(n);
  }
  p.tc_off = o;
  lik_tc_set_sm_count(c->sm_count);
  p.tc_bytes = (lik_impl == SCDEF_LIK_TC) ? lik_tc_workspace_bytes(B, S, c->K0, c->d.n_genes, c->d.n_batches) : 0;
  o += align_up(p.tc_bytes);
  p.total = o;
  return p;
}

NoiseTag make_tag(const scdef_noise* nz, int blk, int rows, int cols) {
  NoiseTag t;
  if (nz->coupling) { t.a = (uint32_t)rows; t.b = (uint32_t)cols; }
  else { t.a = 0x80000000u | (uint32_t)blk; t.b = 0u; }
  return t;
}

PhiloxKey make_key(const scdef_noise* nz) {
  PhiloxKey k;
  k.k0 = (uint32_t)(nz->seed ^ (nz->seed >> 32)) ^ (uint32_t)(nz->step >> 32) * 0x85EBCA6Bu;
  k.k1 = (uint32_t)nz->step;
  return k;
}

bool tc_ok(const scdef_ctx* c) {
  return lik_tc_supported(c->K0, c->d.n_genes, c->d.n_batches);
}

int resolve_lik_impl(const scdef_ctx* c, int requested) {
  if (requested == SCDEF_LIK_AUTO) return tc_ok(c) ? SCDEF_LIK_TC : SCDEF_LIK_SIMT;
  return requested;
}

}  // namespace

extern "C" {

int scdef_abi_version(void) { return SCDEF_ABI_VERSION; }

int scdef_ctx_create(const scdef_model_desc* desc, scdef_ctx** out) {
  if (!desc || !out) return SCDEF_EINVAL;
  if (desc->abi_version != SCDEF_ABI_VERSION) return SCDEF_EINVAL;
  if (desc->n_layers < 1 || desc->n_layers > SCDEF_MAX_LAYERS) return SCDEF_EINVAL;
  if (desc->n_genes < 1 || desc->n_batches < 1 || desc->n_cells < 0) return SCDEF_EINVAL;
  scdef_ctx* c = new (std::nothrow) scdef_ctx;
  if (!c) return SCDEF_EINVAL;
  memset(c, 0, sizeof(*c));
  c->d = *desc;
  c->L = desc->n_layers;
  c->off[0] = 0;
  for (int l = 0; l < c->L; ++l) {
    if (desc->layer_sizes[l] < 1) { delete c; return SCDEF_EINVAL; }
    c->K[l] = desc->layer_sizes[l];
    c->off[l + 1] = c->off[l] + c->K[l];
    c->out_dim[l] = l == 0 ? desc->n_genes : desc->layer_sizes[l - 1];
  }
  c->Kt = c->off[c->L];
  c->K0 = c->K[0];
  if (c->K0 > 256) { delete c; return SCDEF_EUNSUPPORTED; }
  int dev = 0;
  c->sm_count = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) c->sm_count = n;
  }
  cudaGetLastError();
  *out = c;
  return SCDEF_OK;
}

void scdef_ctx_destroy(scdef_ctx* ctx) {
  if (!ctx) return;
  if (ctx->prof_ready)
    for (int k = 0; k < SCDEF_PROF_KINDS; ++k)
      for (int i = 0; i < PROF_RING; ++i) { cudaEventDestroy(ctx->prof_ev[k][i][0]); cudaEventDestroy(ctx->prof_ev[k][i][1]); }
  delete ctx;
}

int scdef_profile_enable(scdef_ctx* ctx, int on) {
  if (!ctx) return SCDEF_EINVAL;
  if (on && !ctx->prof_ready) {
    for (int k = 0; k < SCDEF_PROF_KINDS; ++k)
      for (int i = 0; i < PROF_RING; ++i)
        if (cudaEventCreate(&ctx->prof_ev[k][i][0]) != cudaSuccess || cudaEventCreate(&ctx->prof_ev[k][i][1]) != cudaSuccess)
          return fail(ctx, SCDEF_ECUDA, "cudaEventCreate failed");
    ctx->prof_ready = true;
  }
  ctx->prof_on = on != 0;
  ctx->prof_mask = on == 1 ? 0xffffffffu : ((unsigned)on >> 1);
  return SCDEF_OK;
}

int scdef_profile_read(scdef_ctx* ctx, double* ms_sum, int64_t* n_timed, int64_t* n_launched, int64_t* total_launches, int reset) {
  if (!ctx) return SCDEF_EINVAL;
  for (int k = 0; k < SCDEF_PROF_KINDS; ++k) {
    double sum = 0.0;
    long long n = ctx->prof_count[k] < PROF_RING ? ctx->prof_count[k] : PROF_RING;
    if (ctx->prof_ready)
      for (long long i = 0; i < n; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(ctx->prof_ev[k][i][1]) != cudaSuccess ||
            cudaEventElapsedTime(&ms, ctx->prof_ev[k][i][0], ctx->prof_ev[k][i][1]) != cudaSuccess)
          return fail(ctx, SCDEF_ECUDA, "profile read failed for kind %d", k);
        sum += ms;
      }
    if (ms_sum) ms_sum[k] = sum;
    if (n_timed) n_timed[k] = n;
    if (n_launched) n_launched[k] = ctx->prof_count[k];
    if (reset) ctx->prof_count[k] = 0;
  }
  if (total_launches) *total_launches = ctx->launches;
  if (reset) ctx->launches = 0;
  return SCDEF_OK;
}

const char* scdef_last_error(scdef_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

size_t scdef_workspace_bytes(scdef_ctx* ctx, int B_max, int S, int lik_impl) {
  if (!ctx || B_max < 0 || S < 1) return 0;
  return make_plan(ctx, B_max < 1 ? 1 : B_max, S, resolve_lik_impl(ctx, lik_impl)).total + 256;
}

int scdef_elbo_grad(scdef_ctx* ctx, const scdef_call* call, const scdef_blocks* params,
                    const scdef_noise* noise, double* loss_out, const scdef_blocks* grads,
                    void* workspace, size_t workspace_bytes, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!call || !params || !noise || !loss_out || !grads || !workspace)
    return fail(ctx, SCDEF_EINVAL, "null argument");
  const int pass = call->pass;
  if (pass != SCDEF_PASS_LOCAL && pass != SCDEF_PASS_GLOBAL) return fail(ctx, SCDEF_EINVAL, "bad pass %d", pass);
  const int S = call->num_samples, B = call->batch;
  if (S < 1 || B < 0) return fail(ctx, SCDEF_EINVAL, "bad S=%d or B=%d", S, B);
  if (B > 0 && !call->indices) return fail(ctx, SCDEF_EINVAL, "indices is null");
  if (B > 0 && !call->X_batch && !ctx->d.X) return fail(ctx, SCDEF_EINVAL, "no counts: X_batch and desc.X are both null");
  if (((uintptr_t)workspace & 15) != 0) return fail(ctx, SCDEF_EALIGN, "workspace not 16-byte aligned");
  const int lik_impl = resolve_lik_impl(ctx, call->lik_impl);
  if (lik_impl == SCDEF_LIK_TC && !tc_ok(ctx))
    return fail(ctx, SCDEF_EUNSUPPORTED, "tcgen05 likelihood kernel does not take K0=%d", ctx->K0);
  const Plan plan = make_plan(ctx, B < 1 ? 1 : B, S, lik_impl);
  if (plan.total > workspace_bytes)
    return fail(ctx, SCDEF_EWORKSPACE, "workspace %zu < %zu bytes", workspace_bytes, plan.total);
  if (noise->mode == SCDEF_NOISE_EXPLICIT) {
    bool ok = noise->eps_cell_scale && noise->eps_z && noise->eps_gene_scale;
    if (ctx->d.use_brd) ok = ok && noise->eps_brd && noise->eps_ard;
    if (ctx->d.marginalize_alpha) ok = ok && noise->eps_alpha;
    for (int l = 0; l < ctx->L; ++l) ok = ok && noise->eps_W[l];
    if (!ok) return fail(ctx, SCDEF_EINVAL, "explicit noise: missing eps pointer");
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  char* ws = (char*)workspace;
  const scdef_model_desc& d = ctx->d;
  const int L = ctx->L, K0 = ctx->K0, G = d.n_genes, nb = d.n_batches;
  const bool global_pass = pass == SCDEF_PASS_GLOBAL;
  const bool explicit_noise = noise->mode == SCDEF_NOISE_EXPLICIT;
  const float gw = call->global_weight, scale = call->scale, anneal = call->annealing;
  const PhiloxKey key = make_key(noise);
  const float* sg = call->stop_gradients;
  const bool lik_on = !(sg[0] != 0.f);
  const bool tc = (lik_impl == SCDEF_LIK_TC);
  auto smp = [&](int b) { return (float*)(ws + plan.smp[b]); };
  auto Gb = [&](int b) { return (float*)(ws + plan.G[b]); };

  // ---- block table ----------------------------------------------------------------------
  // data-parallel row window of the per-cell noise (scdef_noise.local_row_offset / local_rows_total)
  const long long noise_rows_total = noise->local_rows_total > 0 ? noise->local_rows_total : B;
  const long long noise_row0 = noise->local_rows_total > 0 ? noise->local_row_offset : 0;
  if (noise_row0 < 0 || noise_row0 + B > noise_rows_total || noise_rows_total > 0x7fffffffLL)
    return fail(ctx, SCDEF_EINVAL, "noise row window out of range");
  BlockTable all;
  memset(&all, 0, sizeof(all));
  auto fill = [&](int blk, const float* base, int rows, int cols, int ld, int col0, bool local,
                  const float* eps, int eps_ld, int eps_col0, float mu_hi, float ent_w, bool stopped,
                  bool active, float* gbase, int g_rows, int g_ld, int g_col0) {
    BlockDesc& b = all.b[blk];
    const long long half = (long long)(local ? d.n_cells : rows) * ld;
    b.mu = base; b.rho = base ? base + half : nullptr;
    b.gather = local ? call->indices : nullptr;
    b.eps = explicit_noise ? eps : nullptr;
    b.fixed = nullptr;
    b.smp = smp(blk); b.G = Gb(blk);
    b.gmu = gbase; b.grho = gbase ? gbase + (long long)g_rows * g_ld : nullptr;
    b.rows = rows; b.cols = cols; b.ld = ld; b.col0 = col0;
    b.eps_ld = eps_ld; b.eps_col0 = eps_col0; b.g_ld = g_ld; b.g_col0 = g_col0;
    b.tag = make_tag(noise, blk, local ? (int)noise_rows_total : rows, cols);
    b.noise_e0 = local ? noise_row0 * cols : 0;
    b.mu_hi = mu_hi; b.ent_weight = ent_w; b.g_weight = 1.f;
    b.stopped = stopped; b.active = active;
  };
  const bool st0 = sg[0] != 0.f;
  fill(BLK_C, params->cell_scale, B, 1, 1, 0, true, noise->eps_cell_scale, 1, 0, LOG_1E6, scale,
       st0 || call->stop_cell_budgets != 0.f, true, grads->cell_scale, B, 1, 0);
  fill(BLK_S, params->gene_scale, nb, G, G, 0, false, noise->eps_gene_scale, G, 0, INFINITY, gw,
       st0 || call->stop_gene_budgets != 0.f, true, grads->gene_scale, nb, G, 0);
  fill(BLK_TAU, params->brd, K0, 1, 1, 0, false, noise->eps_brd, 1, 0, LOG_1E6, gw, st0, d.use_brd != 0,
       grads->brd, K0, 1, 0);
  fill(BLK_OM, params->ard, K0, 1, 1, 0, false, noise->eps_ard, 1, 0, LOG_1E6, gw, false, d.use_brd != 0,
       grads->ard, K0, 1, 0);
  fill(BLK_A, params->alpha, 1, 1, 1, 0, false, noise->eps_alpha, 1, 0, LOG_1E6, gw, false,
       d.marginalize_alpha != 0, grads->alpha, 1, 1, 0);
  for (int l = 0; l < L; ++l) {
    const bool stl = sg[l] != 0.f;
    fill(BLK_W + l, params->W[l], ctx->K[l], ctx->out_dim[l], ctx->out_dim[l], 0, false, noise->eps_W[l],
         ctx->out_dim[l], 0, LOG_1E6, anneal * gw, stl, true, grads->W[l], ctx->K[l], ctx->out_dim[l], 0);
    const bool pinned = (l == d.supervised_layer) && d.fixed_top_z;
    fill(BLK_Z + l, params->z, B, ctx->K[l], ctx->Kt, ctx->off[l], true, noise->eps_z, ctx->Kt, ctx->off[l],
         LOG_1E6, pinned ? 0.f : anneal * scale, stl || pinned, true, grads->z, B, ctx->Kt, ctx->off[l]);
    if (pinned) all.b[BLK_Z + l].fixed = d.fixed_top_z;
  }

  int rc;
  // ---- K1 sampling (+ entropy values) ----------------------------------------------------
  if (B > 0 || true) {
    BlockTable t;
    memset(&t, 0, sizeof(t));
    size_t max_el = 1;
    auto add = [&](int blk) {
      const BlockDesc& b = all.b[blk];
      if (!b.active || b.rows * b.cols == 0) return;
      if (tc && blk == BLK_W) return;  // W_0 is sampled inside the fused kernel
      t.b[t.n++] = b;
      size_t n = (size_t)S * b.rows * b.cols;
      if (n > max_el) max_el = n;
    };
    if (B > 0) add(BLK_C);
    add(BLK_S); add(BLK_TAU); add(BLK_OM); add(BLK_A);
    for (int l = 0; l < L; ++l) { add(BLK_W + l); if (B > 0) add(BLK_Z + l); }
    size_t gx = (max_el / 4 + 255) / 256;
    if (gx > (size_t)ctx->sm_count * 16) gx = (size_t)ctx->sm_count * 16;
    if (gx < 1) gx = 1;
    ProfScope ps(ctx, SCDEF_PROF_SAMPLE, st);
    if (sample_impl() == 2) {
      // one-dimensional grid sized per block; a thread owns a quad for a chunk of the samples (kernels_small.cuh)
      SampleGrid sgrid;
      long long quads_total = 0;
      int nq[MAX_BLOCKS];
      for (int i = 0; i < t.n; ++i) {
        const long long e0 = t.b[i].noise_e0, per = (long long)t.b[i].rows * t.b[i].cols;
        nq[i] = (int)(((e0 + per + 3) >> 2) - (e0 >> 2));
        quads_total += nq[i];
      }
      const long long target = (long long)ctx->sm_count * 2048;   // resident threads of the whole GPU
      int sch = (int)((target + quads_total - 1) / (quads_total > 0 ? quads_total : 1));
      if (sch < 1) sch = 1;
      if (sch > S) sch = S;
      int cta = 0;
      for (int i = 0; i < t.n; ++i) {
        sgrid.cta0[i] = cta;
        sgrid.sch[i] = sch;
        cta += (int)(((long long)nq[i] * sch + 255) / 256);
      }
      sgrid.cta0[t.n] = cta;
      if (cta > 0) SCDEF_LAUNCH(k_sample_v2, (unsigned)cta, 256, 0, st, t, sgrid, S, key, loss_out, gw);
    } else {
      SCDEF_LAUNCH(k_sample, dim3((unsigned)gx, t.n), 256, 0, st, t, S, key, loss_out, gw);
    }
    if ((rc = check_launch(ctx, "k_sample"))) return rc;
  }

  // In the LOCAL pass the priors of the global blocks only contribute to the loss value: nothing to do when the caller
  // discards it (SCDEF_FLAG_SKIP_LOSS).
  const bool value_only_skipped = !global_pass && (call->flags & SCDEF_FLAG_SKIP_LOSS);

  // ---- scale priors ------------------------------------------------------------------------
  {
    SimplePriorTable t;
    memset(&t, 0, sizeof(t));
    auto set = [&](int i, int blk, float a, const float* rate, const int64_t* gather, float rate_scale,
                   float weight, int is_cell, bool active, bool want_G) {
      SimplePrior& p = t.p[i];
      p.smp = smp(blk); p.G = want_G ? Gb(blk) : nullptr;
      p.rate = rate; p.gather = gather;
      p.rows = all.b[blk].rows; p.cols = all.b[blk].cols;
      p.a = a; p.lgam_a = lgammaf(a); p.rate_scale = rate_scale; p.weight = weight;
      p.is_cell = is_cell; p.active = active && p.rows * p.cols > 0;
    };
    set(0, BLK_C, d.cell_scale_shape, d.batch_lib_ratio, call->indices, d.cell_scale_shape, scale, 1, B > 0, !global_pass);
    const bool glob_on = !value_only_skipped;
    set(1, BLK_S, d.gene_scale_shape, d.gene_ratio, nullptr, d.gene_scale_shape, gw, 0, glob_on, global_pass);
    set(2, BLK_TAU, d.brd, nullptr, nullptr, d.brd / d.brd_mean, gw, 0, glob_on && d.use_brd != 0, global_pass);
    set(3, BLK_OM, d.shrinkage_shape, nullptr, nullptr, d.shrinkage_rate, gw, 0, glob_on && d.use_brd != 0, global_pass);
    set(4, BLK_A, 2.f, nullptr, nullptr, 2.f / fmaxf(call->alpha, 1e-8f), gw, 0, glob_on && d.marginalize_alpha != 0, global_pass);
    size_t max_el = (size_t)S * (size_t)((glob_on && (size_t)nb * G > (size_t)B) ? (size_t)nb * G : (size_t)B);
    size_t gx = (max_el + 255) / 256;
    if (gx > (size_t)ctx->sm_count * 8) gx = (size_t)ctx->sm_count * 8;
    if (gx < 1) gx = 1;
    ProfScope ps(ctx, SCDEF_PROF_PRIORS, st);
    SCDEF_LAUNCH(k_scale_priors, dim3((unsigned)gx, 5), 256, 0, st, t, S, loss_out);
    if ((rc = check_launch(ctx, "k_scale_priors"))) return rc;
  }

  // ---- W priors (value only in the LOCAL pass) ----------------------------------------------------
  for (int l = 0; l < L; ++l) {
    if (tc && l == 0) continue;  // fused into the tcgen05 kernel
    if (value_only_skipped) continue;
    WPriorArgs a;
    memset(&a, 0, sizeof(a));
    a.smp_W = smp(BLK_W + l);
    a.G_W = global_pass ? Gb(BLK_W + l) : nullptr;
    a.P = d.w_prior_shape[l]; a.R = d.w_prior_rate[l];
    a.P_scalar = d.w_prior_shape_scalar[l]; a.R_scalar = d.w_prior_rate_scalar[l];
    a.smp_tau = smp(BLK_TAU); a.smp_om = smp(BLK_OM);
    a.G_tau = Gb(BLK_TAU); a.G_om = Gb(BLK_OM);
    a.rows = ctx->K[l]; a.cols = ctx->out_dim[l];
    a.mode = (d.use_brd && l == 0) ? 1 : ((d.use_brd && l == 1) ? 2 : 0);
    a.K_l = (float)ctx->K[l];
    a.weight = gw;
    int warps = S * a.rows;
    ProfScope ps(ctx, SCDEF_PROF_WPRIOR, st);
    SCDEF_LAUNCH(k_wprior, (warps + 7) / 8, 256, 0, st, a, S, loss_out);
    if ((rc = check_launch(ctx, "k_wprior"))) return rc;
  }

  // ---- hierarchy ---------------------------------------------------------------------------
  // (its outputs — dELBO/dv of z, of W_l for l >= 1 and of alpha — are disjoint from the likelihood group's in the GLOBAL
  // pass, so SCDEF_FLAG_LIK_FIRST may run it after that group: the W_0 gradient, 97 % of the data-parallel all-reduce, is
  // then final a hierarchy kernel earlier)
  auto run_hier = [&]() -> int {
    if (B <= 0) return SCDEF_OK;
    HierArgs h;
    memset(&h, 0, sizeof(h));
    h.L = L;
    int wtot = 0;
    for (int l = 0; l < L; ++l) {
      h.K[l] = ctx->K[l]; h.off[l] = ctx->off[l];
      h.smp_z[l] = smp(BLK_Z + l); h.G_z[l] = Gb(BLK_Z + l);
      h.smp_W[l] = smp(BLK_W + l); h.G_W[l] = Gb(BLK_W + l);
      if (l >= 1) {
        h.wld[l] = ctx->K[l - 1] | 1;
        h.woff[l] = wtot;
        wtot += ctx->K[l] * h.wld[l];
      }
    }
    h.off[L] = ctx->Kt;
    h.wtot = wtot;
    h.smp_a = d.marginalize_alpha ? smp(BLK_A) : nullptr;
    h.G_a = (d.marginalize_alpha && global_pass) ? Gb(BLK_A) : nullptr;
    h.caf = d.cell_alpha_factor; h.idx = call->indices;
    h.B = B; h.pass = pass;
    const bool root_frozen = sg[L - 1] > 0.5f;
    h.top_layer = root_frozen ? L - 2 : L - 1;
    h.alpha = call->alpha; h.alpha_floor = d.alpha_floor; h.top_alpha = d.top_alpha;
    h.lgam_top = lgammaf(d.top_alpha); h.scale = scale;
    // register-tiled kernel when its shared-memory plan fits two CTAs per SM, else the warp-per-cell kernel
    HierTArgs ht;
    memset(&ht, 0, sizeof(ht));
    int wtot4 = 0, kmax = 1;
    for (int l = 0; l < L; ++l) {
      if (ctx->K[l] > kmax) kmax = ctx->K[l];
      if (l >= 1) {
        ht.wld[l] = (ctx->K[l - 1] + 3) / 4 * 4;
        ht.woff[l] = wtot4;
        wtot4 += ctx->K[l] * ht.wld[l];
      }
    }
    ht.wtot = wtot4; ht.kmax = kmax;
    const size_t smem_t = ((size_t)wtot4 * (global_pass ? 2 : 1) + (size_t)ctx->Kt * HLD + (size_t)kmax * HLD * (global_pass ? 1 : 3)) * sizeof(float);
    static const bool hier_old = getenv("SCDEF_HIER_OLD") != nullptr;
    if (smem_t <= 110 * 1024 && kmax <= 128 && !hier_old) {
      const int n_t32 = (B + HT - 1) / HT;
      int tpc = (int)((long long)n_t32 * S / (2 * ctx->sm_count * 2));  // aim for >= ~2 waves of 2 CTAs/SM
      if (tpc < 1) tpc = 1;
      if (tpc > n_t32) tpc = n_t32;
      ht.h = h;
      ht.tiles_per_cta = tpc;
      static const bool hier_mma = getenv("SCDEF_HIER_MMA") != nullptr;   // tensor-core form of the three small products (opt-in)
      ProfScope ps(ctx, SCDEF_PROF_HIER, st);
      if (hier_mma) {
        if (smem_t > 48 * 1024) cudaFuncSetAttribute(k_hier_tiled<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
        SCDEF_LAUNCH((k_hier_tiled<true>), dim3((n_t32 + tpc - 1) / tpc, S), HNT, smem_t, st, ht, S, loss_out);
      } else {
        if (smem_t > 48 * 1024) cudaFuncSetAttribute(k_hier_tiled<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t);
        SCDEF_LAUNCH((k_hier_tiled<false>), dim3((n_t32 + tpc - 1) / tpc, S), HNT, smem_t, st, ht, S, loss_out);
      }
    } else {
    // enough CTAs to fill the GPU, at least one round of 8 cells each
    int chunks = (ctx->sm_count * 2 + S - 1) / S;
    int max_chunks = (B + HIER_WARPS - 1) / HIER_WARPS;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    h.cells_per_block = ((B + chunks - 1) / chunks + HIER_WARPS - 1) / HIER_WARPS * HIER_WARPS;
    chunks = (B + h.cells_per_block - 1) / h.cells_per_block;
    size_t smem = ((size_t)2 * wtot + (size_t)3 * HIER_WARPS * ctx->Kt) * sizeof(float);
    if (smem > 200 * 1024) return fail(ctx, SCDEF_EUNSUPPORTED, "hierarchy needs %zu B of shared memory", smem);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_hier, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    {
      ProfScope ps(ctx, SCDEF_PROF_HIER, st);
      SCDEF_LAUNCH(k_hier, dim3(chunks, S), HIER_WARPS * 32, smem, st, h, S, loss_out);
    }
    }
    return check_launch(ctx, "k_hier");
  };
  const bool lik_first = pass == SCDEF_PASS_GLOBAL && (call->flags & SCDEF_FLAG_LIK_FIRST);
  if (!lik_first && (rc = run_hier())) return rc;

  // ---- likelihood (+ the whole W_0 path in tcgen05 mode) --------------------------------------------
  {
    const float* Xp = call->X_batch ? call->X_batch : d.X;
    const int64_t* idx_x = call->X_batch ? nullptr : call->indices;
    const bool lik_run = lik_on && B > 0;
    if (lik_run && !(call->flags & SCDEF_FLAG_SKIP_LOSS)) {   // a pure loss term: nothing to do when the loss is discarded
      const long long xlg_warps = d.x_lgamma_rowsum ? (long long)B : (long long)B * ((G + XLG_CHUNK - 1) / XLG_CHUNK);
      SCDEF_LAUNCH(k_x_lgamma, (unsigned)((xlg_warps + 7) / 8), 256, 0, st, Xp, idx_x, d.x_lgamma_rowsum, call->indices, B, G, scale, loss_out);
      if ((rc = check_launch(ctx, "k_x_lgamma"))) return rc;
    }
    if (tc) {
      LikTcArgs a;
      memset(&a, 0, sizeof(a));
      a.mu_W0 = params->W[0]; a.rho_W0 = params->W[0] + (long long)K0 * G;
      a.eps_W0 = explicit_noise ? noise->eps_W[0] : nullptr;
      a.key = key; a.tag = all.b[BLK_W].tag;
      a.smp_z0 = smp(BLK_Z); a.smp_c = smp(BLK_C); a.smp_s = smp(BLK_S);
      a.smp_tau = smp(BLK_TAU); a.smp_om = smp(BLK_OM);
      a.X = Xp; a.idx_x = idx_x; a.idx = call->indices; a.batch_id = d.batch_id;
      a.G_z0 = Gb(BLK_Z); a.G_c = Gb(BLK_C); a.G_s = Gb(BLK_S);
      a.G_tau = Gb(BLK_TAU); a.G_om = Gb(BLK_OM);
      a.gmu_W0 = grads->W[0]; a.grho_W0 = grads->W[0] ? grads->W[0] + (long long)K0 * G : nullptr;
      a.P = d.w_prior_shape[0]; a.R = d.w_prior_rate[0];
      a.P_scalar = d.w_prior_shape_scalar[0]; a.R_scalar = d.w_prior_rate_scalar[0];
      a.use_brd = d.use_brd; a.B = B; a.G = G; a.K0 = K0; a.nb = nb; a.pass = pass; a.S = S;
      a.scale = scale; a.global_weight = gw; a.anneal = anneal; a.w0_stopped = st0; a.lik_on = lik_run;
      a.reuse_w0 = (call->flags & SCDEF_FLAG_REUSE_W0_SAMPLES) ? 1 : 0;
      a.want_loss = (call->flags & SCDEF_FLAG_SKIP_LOSS) ? 0 : 1;
      a.ws = ws + plan.tc_off; a.ws_bytes = plan.tc_bytes;
      {
        ProfScope ps(ctx, SCDEF_PROF_LIK, st);
        // the fused kernel alone, LOCAL and GLOBAL launches timed separately
        const int mk = pass == SCDEF_PASS_GLOBAL ? SCDEF_PROF_LIK_MAIN_GLOBAL : SCDEF_PROF_LIK_MAIN;
        const bool pm = ctx->prof_on && ctx->prof_ready && lik_run && ((ctx->prof_mask >> mk) & 1u);
        const int slot = (int)(ctx->prof_count[mk] % PROF_RING);
        if (pm) { a.ev0 = ctx->prof_ev[mk][slot][0]; a.ev1 = ctx->prof_ev[mk][slot][1]; }
        rc = lik_tc_launch(a, loss_out, ctx->sm_count, st);
        if (pm && rc == 0) ctx->prof_count[mk]++;
      }
      ctx->launches += 4;
      if (rc) return fail(ctx, rc, "lik_tc_launch failed (%d)", rc);
      if ((rc = check_launch(ctx, "k_lik_tc"))) return rc;
    } else if (lik_run) {
      LikArgs a;
      memset(&a, 0, sizeof(a));
      a.smp_z0 = smp(BLK_Z); a.smp_W0 = smp(BLK_W); a.smp_c = smp(BLK_C); a.smp_s = smp(BLK_S);
      a.X = Xp; a.idx_x = idx_x; a.idx = call->indices; a.batch_id = d.batch_id;
      a.G_z0 = Gb(BLK_Z); a.G_c = Gb(BLK_C); a.G_W0 = Gb(BLK_W); a.G_s = Gb(BLK_S);
      a.B = B; a.G = G; a.K0 = K0; a.nb = nb; a.pass = pass; a.scale = scale;
      size_t smem = ((size_t)LT * (K0 + 1) + (size_t)K0 * LW_LD + (size_t)LT * LQ_LD) * sizeof(float);
      if (smem > 48 * 1024) cudaFuncSetAttribute(k_lik_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((G + LT - 1) / LT, (B + LT - 1) / LT, S);
      {
        ProfScope ps(ctx, SCDEF_PROF_LIK, st);
        ProfScope pm(ctx, pass == SCDEF_PASS_GLOBAL ? SCDEF_PROF_LIK_MAIN_GLOBAL : SCDEF_PROF_LIK_MAIN, st);
        SCDEF_LAUNCH(k_lik_simt, grid, 256, smem, st, a, S, loss_out);
      }
      if ((rc = check_launch(ctx, "k_lik_simt"))) return rc;
    }
  }

  // the W_0 gradient of the GLOBAL pass is final here on the tcgen05 path (k_tc_w0_final_flat wrote it)
  const bool w0_final_early = pass == SCDEF_PASS_GLOBAL && tc;
  if (w0_final_early && ctx->w0_event) cudaEventRecord(ctx->w0_event, st);
  if (lik_first && (rc = run_hier())) return rc;

  // ---- (mu, rho) gradients of the blocks of this pass ------------------------------------------
  {
    BlockTable t;
    memset(&t, 0, sizeof(t));
    size_t max_el = 1;
    auto add = [&](int blk) {
      const BlockDesc& b = all.b[blk];
      if (!b.gmu || b.rows * b.cols == 0) return;
      if (tc && blk == BLK_W) return;  // written by the fused tcgen05 path
      t.b[t.n++] = b;
      size_t n = (size_t)b.rows * b.cols;
      if (n > max_el) max_el = n;
    };
    if (global_pass) {
      add(BLK_S); add(BLK_TAU); add(BLK_OM); add(BLK_A);
      for (int l = 0; l < L; ++l) add(BLK_W + l);
    } else if (B > 0) {
      add(BLK_C);
      for (int l = 0; l < L; ++l) add(BLK_Z + l);
    }
    if (t.n > 0) {
      size_t gx = (max_el + 31) / 32;   // a CTA per 32 elements of the largest block
      if (gx > (size_t)ctx->sm_count * 16) gx = (size_t)ctx->sm_count * 16;
      ProfScope ps(ctx, SCDEF_PROF_FINISH, st);
      SCDEF_LAUNCH(k_finish, dim3((unsigned)gx, t.n), 256, 0, st, t, S, key);
      if ((rc = check_launch(ctx, "k_finish"))) return rc;
    }
  }
  if (pass == SCDEF_PASS_GLOBAL && ctx->w0_event && !w0_final_early) cudaEventRecord(ctx->w0_event, st);   // CUDA-core path: k_finish wrote it
  return SCDEF_OK;
}

int scdef_set_w0_grad_event(scdef_ctx* ctx, void* cuda_event) {
  if (!ctx) return SCDEF_EINVAL;
  ctx->w0_event = (cudaEvent_t)cuda_event;
  return SCDEF_OK;
}

int scdef_adam_local(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* grads,
                     const scdef_blocks* m, const scdef_blocks* v, const int64_t* indices, int B,
                     int64_t step, float lr, float b1, float b2, float eps,
                     const int32_t* freeze_z_layers, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!params || !grads || !m || !v || step < 1) return fail(ctx, SCDEF_EINVAL, "bad argument");
  if (B > 0 && !indices) return fail(ctx, SCDEF_EINVAL, "indices is null");
  if (!ctx->d.row_slot) return fail(ctx, SCDEF_EINVAL, "desc.row_slot is null");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  AdamHyper h;
  h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps;
  h.bc1 = (float)(1.0 - pow((double)b1, (double)step));
  h.bc2 = (float)(1.0 - pow((double)b2, (double)step));
  h.inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, (double)step)));
  h.inv_bc2 = (float)(1.0 / (1.0 - pow((double)b2, (double)step)));
  const long long N = ctx->d.n_cells;
  if (N == 0) return SCDEF_OK;
  int rc;
  if (B > 0) {
    SCDEF_LAUNCH(k_set_slots, (B + 255) / 256, 256, 0, st, ctx->d.row_slot, indices, B, 0);
    if ((rc = check_launch(ctx, "k_set_slots"))) return rc;
  }
  for (int which = 0; which < 2; ++which) {
    AdamLocalArgs a;
    memset(&a, 0, sizeof(a));
    a.p = which ? params->z : params->cell_scale;
    a.m = which ? m->z : m->cell_scale;
    a.v = which ? v->z : v->cell_scale;
    a.g = which ? grads->z : grads->cell_scale;
    a.row_slot = ctx->d.row_slot;
    a.N = N; a.C = which ? ctx->Kt : 1; a.B = B;
    if (which && freeze_z_layers)
      for (int l = 0; l < ctx->L; ++l)
        if (freeze_z_layers[l]) { a.frozen_lo[a.n_frozen] = ctx->off[l]; a.frozen_hi[a.n_frozen] = ctx->off[l + 1]; a.n_frozen++; }
    const long long half = N * a.C;
    const bool vec4 = (a.C % 4 == 0) && ((half * 4) % 16 == 0) && ((long long)B * a.C * 4 % 16 == 0) &&
                      (((uintptr_t)a.p | (uintptr_t)a.m | (uintptr_t)a.v | (uintptr_t)a.g) % 16 == 0);
    const long long nvec = vec4 ? half / 4 : half;
    long long blocks = (nvec + 511) / 512;  // two vectors per thread
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const bool small = nvec < (1LL << 30);
    {
      ProfScope ps(ctx, which ? SCDEF_PROF_ADAM_LOCAL : SCDEF_PROF_ADAM_LOCAL_C, st);
      dim3 grid((unsigned)blocks, 2);
      if (vec4 && small) SCDEF_LAUNCH((k_adam_local<4, uint32_t>), grid, 256, 0, st, a, h);
      else if (vec4) SCDEF_LAUNCH((k_adam_local<4, unsigned long long>), grid, 256, 0, st, a, h);
      else if (small) SCDEF_LAUNCH((k_adam_local<1, uint32_t>), grid, 256, 0, st, a, h);
      else SCDEF_LAUNCH((k_adam_local<1, unsigned long long>), grid, 256, 0, st, a, h);
    }
    if ((rc = check_launch(ctx, "k_adam_local"))) return rc;
  }
  if (B > 0) {
    SCDEF_LAUNCH(k_set_slots, (B + 255) / 256, 256, 0, st, ctx->d.row_slot, indices, B, 1);
    if ((rc = check_launch(ctx, "k_set_slots(clear)"))) return rc;
  }
  return SCDEF_OK;
}

int scdef_adam_bias_table(float b1, float b2, float* host_out, int max_len) {
  // (1/(1-b1^k), 1/(1-b2^k)) for k = 1.., with exactly the arithmetic scdef_adam_local uses for its per-step factors
  if (!host_out || max_len < 1) return SCDEF_EINVAL;
  int k = 1;
  for (; k <= max_len; ++k) {
    const float i1 = (float)(1.0 / (1.0 - pow((double)b1, (double)k)));
    const float i2 = (float)(1.0 / (1.0 - pow((double)b2, (double)k)));
    host_out[2 * (k - 1)] = i1;
    host_out[2 * (k - 1) + 1] = i2;
    if (i1 == 1.f && i2 == 1.f) return k;
  }
  return max_len;
}

int scdef_adam_local_lazy(scdef_ctx* ctx, int phase, const scdef_blocks* params, const scdef_blocks* grads,
                          const scdef_blocks* m, const scdef_blocks* v, const int64_t* indices, int B,
                          int64_t step, float lr, float b1, float b2, float eps,
                          const int32_t* freeze_z_layers, int32_t* row_step, const float* bc_table,
                          int bc_len, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!params || !m || !v || !row_step || !bc_table || bc_len < 1 || step < 0) return fail(ctx, SCDEF_EINVAL, "bad argument");
  if (phase != SCDEF_LAZY_CATCHUP && phase != SCDEF_LAZY_UPDATE && phase != SCDEF_LAZY_CATCHUP_ALL)
    return fail(ctx, SCDEF_EINVAL, "bad phase %d", phase);
  if (phase != SCDEF_LAZY_CATCHUP_ALL && B > 0 && !indices) return fail(ctx, SCDEF_EINVAL, "indices is null");
  if (phase == SCDEF_LAZY_UPDATE && (!grads || step < 1)) return fail(ctx, SCDEF_EINVAL, "UPDATE needs grads and step >= 1");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const long long N = ctx->d.n_cells;
  if (N == 0) return SCDEF_OK;
  if (phase != SCDEF_LAZY_CATCHUP_ALL && B == 0) return SCDEF_OK;
  AdamHyper h;
  h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps;
  const double s1 = (double)(step < 1 ? 1 : step);
  h.bc1 = (float)(1.0 - pow((double)b1, s1));
  h.bc2 = (float)(1.0 - pow((double)b2, s1));
  h.inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, s1)));
  h.inv_bc2 = (float)(1.0 / (1.0 - pow((double)b2, s1)));
  int rc;
  AdamLazyPair pr;
  memset(&pr, 0, sizeof(pr));
  long long max_elems = 1;
  for (int which = 0; which < 2; ++which) {
    AdamLazyArgs& a = pr.a[which];
    a.p = which ? params->z : params->cell_scale;
    a.m = which ? m->z : m->cell_scale;
    a.v = which ? v->z : v->cell_scale;
    a.g = grads ? (which ? grads->z : grads->cell_scale) : nullptr;
    a.idx = phase == SCDEF_LAZY_CATCHUP_ALL ? nullptr : indices;
    a.row_step = row_step;
    a.bc_tab = reinterpret_cast<const float2*>(bc_table);
    a.N = N; a.C = which ? ctx->Kt : 1; a.B = B; a.tab_len = bc_len;
    a.monotone = (b1 <= 0.95f && b2 >= 0.99f && b2 < 1.f && b1 > 0.f) ? 1 : 0;  // b1 (1-b1^k)/(1-b1^(k+1)) sqrt((1-b2^(k+1))/(b2 (1-b2^k))) < 1 for all k
    a.target = phase == SCDEF_LAZY_CATCHUP ? (int)step - 1 : (int)step;
    if (which && freeze_z_layers)
      for (int l = 0; l < ctx->L; ++l)
        if (freeze_z_layers[l]) { a.frozen_lo[a.n_frozen] = ctx->off[l]; a.frozen_hi[a.n_frozen] = ctx->off[l + 1]; a.n_frozen++; }
    const long long rows = a.idx ? B : N;
    if (rows * a.C > max_elems) max_elems = rows * a.C;
  }
  {
    // one launch for both local blocks: the short serial replay of the (C = 1) cell-scale block hides under z's
    long long blocks = (max_elems + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    ProfScope ps(ctx, SCDEF_PROF_ADAM_LOCAL, st);
    if (phase == SCDEF_LAZY_UPDATE) SCDEF_LAUNCH(k_adam_lazy_update, dim3((unsigned)blocks, 2, 2), 256, 0, st, pr, h);
    else SCDEF_LAUNCH(k_adam_lazy_catchup, dim3((unsigned)blocks, 2, 2), 256, 0, st, pr, h);
  }
  if ((rc = check_launch(ctx, "k_adam_lazy"))) return rc;
  // the rows are now current up to `target`
  {
    const long long rows = phase == SCDEF_LAZY_CATCHUP_ALL ? N : B;
    const int target = phase == SCDEF_LAZY_CATCHUP ? (int)step - 1 : (int)step;
    SCDEF_LAUNCH(k_set_row_step, (unsigned)((rows + 255) / 256), 256, 0, st, row_step, phase == SCDEF_LAZY_CATCHUP_ALL ? nullptr : indices, rows, target);
    if ((rc = check_launch(ctx, "k_set_row_step"))) return rc;
  }
  return SCDEF_OK;
}

int scdef_adam_global(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* grads,
                      const scdef_blocks* m, const scdef_blocks* v, int64_t step, float lr, float b1,
                      float b2, float eps, int freeze_w, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!params || !grads || !m || !v || step < 1) return fail(ctx, SCDEF_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  AdamHyper h;
  h.lr = lr; h.b1 = b1; h.b2 = b2; h.eps = eps;
  h.bc1 = (float)(1.0 - pow((double)b1, (double)step));
  h.bc2 = (float)(1.0 - pow((double)b2, (double)step));
  h.inv_bc1 = (float)(1.0 / (1.0 - pow((double)b1, (double)step)));
  h.inv_bc2 = (float)(1.0 / (1.0 - pow((double)b2, (double)step)));
  AdamSegTable t;
  memset(&t, 0, sizeof(t));
  int max_n = 1;
  auto add = [&](float* p, const float* g, float* mm, float* vv, int n_half, int clip, int frozen) {
    AdamSeg& s = t.s[t.n++];
    s.p = p; s.g = g; s.m = mm; s.v = vv; s.n_half = n_half; s.clip = clip; s.frozen = frozen;
    if (2 * n_half > max_n) max_n = 2 * n_half;
  };
  const int K0 = ctx->K0, G = ctx->d.n_genes, nb = ctx->d.n_batches;
  add(params->gene_scale, grads->gene_scale, m->gene_scale, v->gene_scale, nb * G, 1, 0);
  add(params->brd, grads->brd, m->brd, v->brd, K0, 1, 0);
  for (int l = 0; l < ctx->L; ++l)
    add(params->W[l], grads->W[l], m->W[l], v->W[l], ctx->K[l] * ctx->out_dim[l], 1, freeze_w ? 1 : 0);
  add(params->ard, grads->ard, m->ard, v->ard, K0, 0, 0);
  add(params->alpha, grads->alpha, m->alpha, v->alpha, 1, 0, 0);
  for (int i = 0; i < t.n; ++i)
    if (!t.s[i].p || !t.s[i].g || !t.s[i].m || !t.s[i].v) return fail(ctx, SCDEF_EINVAL, "null block pointer");
  int gx = (max_n + 255) / 256;
  if (gx > ctx->sm_count * 8) gx = ctx->sm_count * 8;
  {
    ProfScope ps(ctx, SCDEF_PROF_ADAM_GLOBAL, st);
    SCDEF_LAUNCH(k_adam_global, dim3(gx, t.n), 256, 0, st, t, h);
  }
  return check_launch(ctx, "k_adam_global");
}

int scdef_posterior(scdef_ctx* ctx, const scdef_blocks* params, const scdef_blocks* means,
                    const scdef_blocks* vars, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!params) return fail(ctx, SCDEF_EINVAL, "null params");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  PostSegTable t;
  memset(&t, 0, sizeof(t));
  long long max_n = 1;
  auto add = [&](const float* p, float* mean, float* var, long long n) {
    if (!p || (!mean && !var) || n == 0) return;
    PostSeg& s = t.s[t.n++];
    s.p = p; s.mean = mean; s.var = var; s.n = n;
    if (n > max_n) max_n = n;
  };
  const long long N = ctx->d.n_cells;
  const int K0 = ctx->K0, G = ctx->d.n_genes, nb = ctx->d.n_batches;
#define MV(f) (means ? means->f : nullptr), (vars ? vars->f : nullptr)
  add(params->cell_scale, MV(cell_scale), N);
  add(params->z, MV(z), N * ctx->Kt);
  add(params->gene_scale, MV(gene_scale), (long long)nb * G);
  add(params->brd, MV(brd), K0);
  for (int l = 0; l < ctx->L; ++l) add(params->W[l], MV(W[l]), (long long)ctx->K[l] * ctx->out_dim[l]);
  add(params->ard, MV(ard), K0);
  add(params->alpha, MV(alpha), 1);
#undef MV
  if (t.n == 0) return SCDEF_OK;
  long long gx = (max_n + 255) / 256;
  if (gx > (long long)ctx->sm_count * 16) gx = (long long)ctx->sm_count * 16;
  SCDEF_LAUNCH(k_posterior, dim3((unsigned)gx, t.n), 256, 0, st, t);
  return check_launch(ctx, "k_posterior");
}

int scdef_csr_row_stats(const int64_t* indptr, const float* data, int64_t n_rows, float* lib, float* lgam, void* cuda_stream) {
  if (!indptr || !data || n_rows < 0) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  SCDEF_LAUNCH(k_csr_row_stats, (unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)cuda_stream, reinterpret_cast<const long long*>(indptr), data, n_rows, lib, lgam);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_csr_gather_rows(const int64_t* indptr, const int32_t* indices, const float* data, const int64_t* rows,
                          int32_t n_rows, int32_t n_genes, float* out, void* cuda_stream) {
  if (!indptr || !indices || !data || !out || n_rows < 0 || n_genes < 1) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  if (!rows) return SCDEF_EINVAL;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_rows * n_genes, st);
  SCDEF_LAUNCH(k_csr_gather_rows, (n_rows + 7) / 8, 256, 0, st, reinterpret_cast<const long long*>(indptr), indices, data,
                                                      reinterpret_cast<const long long*>(rows), n_rows, n_genes, out);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_u16_row_stats(const uint16_t* X, int64_t n_rows, int32_t n_genes, float* lib, float* lgam, void* cuda_stream) {
  if (!X || n_rows < 0 || n_genes < 1) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  SCDEF_LAUNCH(k_u16_row_stats, (unsigned)((n_rows + 7) / 8), 256, 0, (cudaStream_t)cuda_stream, X, n_rows, n_genes, lib, lgam);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_u16_gather_rows(const uint16_t* X, const int64_t* rows, int32_t n_rows, int32_t n_genes, float* out, void* cuda_stream) {
  if (!X || !out || n_rows < 0 || n_genes < 1) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  if (!rows) return SCDEF_EINVAL;
  const long long total = (long long)n_rows * n_genes;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = (int)((total + 255) / 256 < (long long)sms * 8 ? (total + 255) / 256 : (long long)sms * 8);
  SCDEF_LAUNCH(k_u16_gather_rows, blocks, 256, 0, (cudaStream_t)cuda_stream, X, reinterpret_cast<const long long*>(rows), n_rows, n_genes, out);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_count_stats(const void* X, int32_t x_is_u16, const int32_t* batch_id, int64_t n_rows, int32_t n_genes,
                      int32_t n_batches, double* batch_stats, double* gene_sums, float* lib, void* cuda_stream) {
  if (!X || !batch_stats || !gene_sums || n_rows < 0 || n_genes < 1 || n_batches < 1) return SCDEF_EINVAL;
  if (n_batches > 1 && !batch_id) return SCDEF_EINVAL;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaMemsetAsync(batch_stats, 0, sizeof(double) * 3 * (size_t)(n_batches + 1), st);
  cudaMemsetAsync(gene_sums, 0, sizeof(double) * (size_t)(n_batches + 1) * n_genes, st);
  if (n_rows == 0) return SCDEF_OK;
  const unsigned blocks = (unsigned)((n_rows + 31) / 32);
  if (x_is_u16) SCDEF_LAUNCH((k_count_stats<uint16_t>), blocks, 256, 0, st, (const uint16_t*)X, batch_id, n_rows, n_genes, n_batches, batch_stats, gene_sums, lib);
  else SCDEF_LAUNCH((k_count_stats<float>), blocks, 256, 0, st, (const float*)X, batch_id, n_rows, n_genes, n_batches, batch_stats, gene_sums, lib);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_assignment_counts(scdef_ctx* ctx, const scdef_blocks* params, int64_t* counts, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!params || !params->z || !counts) return fail(ctx, SCDEF_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  cudaMemsetAsync(counts, 0, sizeof(int64_t) * ctx->Kt, st);
  const long long N = ctx->d.n_cells;
  if (N == 0) return SCDEF_OK;
  AssignArgs a;
  memset(&a, 0, sizeof(a));
  a.mu = params->z; a.rho = params->z + N * ctx->Kt;
  a.N = N; a.L = ctx->L; a.Kt = ctx->Kt;
  for (int l = 0; l < ctx->L; ++l) { a.K[l] = ctx->K[l]; a.off[l] = ctx->off[l]; }
  long long blocks = (N + 7) / 8;
  if (blocks > (long long)ctx->sm_count * 8) blocks = (long long)ctx->sm_count * 8;
  SCDEF_LAUNCH(k_assignment_counts, (unsigned)blocks, 256, sizeof(unsigned int) * ctx->Kt, st, a, reinterpret_cast<unsigned long long*>(counts));
  return check_launch(ctx, "k_assignment_counts");
}

/* bring-up hook (not part of the public header): dump the tcgen05 rate tile R as (S,B,G) */
void scdef_debug_tc_dump(float* p) { lik_tc_set_debug(p); }
void scdef_debug_pair_stuck(int* out) { lik_tc_pair_stuck(out); }
void scdef_debug_set_sample_impl(int v) { g_sample_impl = (v == 2) ? 2 : 1; }
/* host logic under test: n draws of the device initialiser's Gamma(shape, 1) sampler, computed on the HOST from the same code */
void scdef_debug_gamma_host(float shape, unsigned long long seed, int stream_id, long long n, float* out_host) {
  PhiloxKey key; key.k0 = (uint32_t)seed; key.k1 = (uint32_t)(seed >> 32);
  NoiseTag tag; tag.a = 0x696e6974u; tag.b = (uint32_t)stream_id << 8;
  for (long long i = 0; i < n; ++i) out_host[i] = philox_gamma(shape, key, tag, (unsigned long long)i);
}
void scdef_debug_tc_geom(int B, int S, int K0, int G, int* out) { lik_tc_debug_geom(B, S, K0, G, out); }
int scdef_debug_flat_chunk_of(long long f, long long T, int P) { return lik_tc_debug_flat_chunk_of(f, T, P); }
void scdef_debug_pair_timing(unsigned long long* out) { cudaDeviceSynchronize(); lik_tc_read_pair_timing(out); }

int scdef_row_stats(const float* X, int64_t n_rows, int32_t n_genes, float* lib, float* lgam, void* cuda_stream) {
  if (!X || n_rows < 0 || n_genes < 1) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  long long blocks = (n_rows + 7) / 8;
  SCDEF_LAUNCH(k_row_stats, (unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream, X, n_rows, n_genes, lib, lgam);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_assign_confident(const float* mu, const float* rho, int64_t n_rows, int32_t ld, const int32_t* cols,
                           int32_t n_cols, int32_t n_samples, float quantile_level, uint64_t seed, int32_t stream_id,
                           int64_t row0, const float* eps, float* stats, int32_t* argmax, void* cuda_stream) {
  if (!mu || !rho || !cols || !stats || !argmax) return SCDEF_EINVAL;
  if (n_rows < 0 || ld < 1 || n_cols < 2 || n_cols > ld || n_samples < 1 || row0 < 0) return SCDEF_EINVAL;
  if (!(quantile_level >= 0.f && quantile_level <= 1.f)) return SCDEF_EINVAL;
  if (n_cols > 1024) return SCDEF_EUNSUPPORTED;                 // 8 quads of 4 factors per lane
  const size_t smem = (size_t)8 * (size_t)n_samples * sizeof(float);   // the S gaps of each of the CTA's 8 cells
  if (smem > 200 * 1024) return SCDEF_EUNSUPPORTED;
  if (n_rows == 0) return SCDEF_OK;
  ConfidentArgs a;
  a.mu = mu; a.rho = rho; a.cols = cols; a.eps = eps; a.stats = stats; a.argmax = argmax;
  a.N = n_rows; a.row0 = row0; a.ld = ld; a.K = n_cols; a.S = n_samples; a.q_level = quantile_level;
  a.key.k0 = (uint32_t)seed; a.key.k1 = (uint32_t)(seed >> 32);
  a.tag.a = 0x61736366u;  /* 'ascf' */
  a.tag.b = (uint32_t)stream_id << 8;   // the high half of the 64-bit quad counter is added to the low byte range
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (n_rows + 7) / 8;
  if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int quads = (n_cols + 3) / 4;
  if (quads <= 32) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_assign_confident<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SCDEF_LAUNCH((k_assign_confident<1>), (unsigned)blocks, 256, smem, st, a);
  } else if (quads <= 64) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_assign_confident<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SCDEF_LAUNCH((k_assign_confident<2>), (unsigned)blocks, 256, smem, st, a);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_assign_confident<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SCDEF_LAUNCH((k_assign_confident<8>), (unsigned)blocks, 256, smem, st, a);
  }
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_signature_scores(const float* w_mu, const float* w_rho, int32_t n_genes, const int32_t* rows, int32_t n_rows,
                           const float* eps, const float* path, int32_t n_draws, int32_t n_factors, int32_t mode, const float* tau,
                           int32_t* counts, const int32_t* ref_idx, int32_t kmax, float* ref_score, int32_t* rank_count,
                           void* cuda_stream) {
  if (!w_mu || !w_rho || !rows || !eps || n_genes < 1 || n_rows < 1 || n_draws < 1 || n_factors < 1) return SCDEF_EINVAL;
  if (n_rows > SIG_MAX_R || n_factors > SIG_MAX_F) return SCDEF_EUNSUPPORTED;
  if (!path && n_factors != n_rows) return SCDEF_EINVAL;
  if (mode == 0 ? (!tau || !counts) : (mode == 1 || mode == 2) ? (!ref_idx || !ref_score || kmax < 1 || (mode == 2 && !rank_count)) : true)
    return SCDEF_EINVAL;
  SigArgs a;
  memset(&a, 0, sizeof(a));
  a.w_mu = w_mu; a.w_rho = w_rho; a.rows = rows; a.eps = eps; a.path = path; a.tau = tau; a.ref_idx = ref_idx;
  a.ref_score = ref_score; a.counts = counts; a.rank_count = rank_count;
  a.R = n_rows; a.G = n_genes; a.S = n_draws; a.F = n_factors; a.kmax = kmax; a.mode = mode;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (n_genes + SIG_TG - 1) / SIG_TG;
  int chunks = (2 * sms + tiles - 1) / tiles;          // about two CTAs per SM
  if (chunks > n_draws) chunks = n_draws;
  a.draws_per_cta = (n_draws + chunks - 1) / chunks;
  chunks = (n_draws + a.draws_per_cta - 1) / a.draws_per_cta;
  const size_t smem = (size_t)(SIG_MAX_R * SIG_TG + SIG_FC * (SIG_MAX_R + 1) + SIG_FC * (SIG_TG + 1)) * sizeof(float);
  cudaFuncSetAttribute(k_signature_scores, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  SCDEF_LAUNCH(k_signature_scores, dim3(tiles, chunks), 256, smem, (cudaStream_t)cuda_stream, a);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_init_gamma_block(float* mu, float* rho, int64_t n_rows, int32_t n_cols, int32_t ld, float shape, float clip_lo,
                           float clip_hi, float var_div, uint64_t seed, int32_t stream_id, int64_t row0, void* cuda_stream) {
  if (!mu || !rho || n_rows < 0 || n_cols < 1 || ld < n_cols || row0 < 0) return SCDEF_EINVAL;
  if (!(shape > 0.f) || !(clip_lo > 0.f) || !(clip_hi >= clip_lo) || !(var_div > 0.f)) return SCDEF_EINVAL;
  if (n_rows == 0) return SCDEF_OK;
  InitGammaArgs a;
  a.mu = mu; a.rho = rho; a.N = n_rows; a.row0 = row0; a.K = n_cols; a.ld = ld;
  a.shape = shape; a.clip_lo = clip_lo; a.clip_hi = clip_hi; a.var_div = var_div;
  a.key.k0 = (uint32_t)seed; a.key.k1 = (uint32_t)(seed >> 32);
  a.tag.a = 0x696e6974u;  /* 'init' */
  a.tag.b = (uint32_t)stream_id << 8;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long blocks = (n_rows * n_cols + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  SCDEF_LAUNCH(k_init_gamma_block, (unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream, a);
  return cudaGetLastError() == cudaSuccess ? SCDEF_OK : SCDEF_ECUDA;
}

int scdef_noise_fill(scdef_ctx* ctx, const scdef_noise* noise, int block_kind, int layer, int S, int B,
                     float* out, void* cuda_stream) {
  if (!ctx) return SCDEF_EINVAL;
  if (!noise || !out || S < 1) return fail(ctx, SCDEF_EINVAL, "bad argument");
  int blk, rows, cols;
  const int K0 = ctx->K0;
  switch (block_kind) {
    case 0: blk = BLK_C; rows = B; cols = 1; break;
    case 1: if (layer < 0 || layer >= ctx->L) return fail(ctx, SCDEF_EINVAL, "bad layer");
            blk = BLK_Z + layer; rows = B; cols = ctx->K[layer]; break;
    case 2: blk = BLK_S; rows = ctx->d.n_batches; cols = ctx->d.n_genes; break;
    case 3: blk = BLK_TAU; rows = K0; cols = 1; break;
    case 4: if (layer < 0 || layer >= ctx->L) return fail(ctx, SCDEF_EINVAL, "bad layer");
            blk = BLK_W + layer; rows = ctx->K[layer]; cols = ctx->out_dim[layer]; break;
    case 5: blk = BLK_OM; rows = K0; cols = 1; break;
    case 6: blk = BLK_A; rows = 1; cols = 1; break;
    default: return fail(ctx, SCDEF_EINVAL, "bad block kind");
  }
  if (rows * cols == 0) return SCDEF_OK;
  long long e0 = 0;
  int tag_rows = rows;
  if (block_kind <= 1 && noise->local_rows_total > 0) {
    if (noise->local_row_offset < 0 || noise->local_row_offset + B > noise->local_rows_total)
      return fail(ctx, SCDEF_EINVAL, "noise row window out of range");
    e0 = noise->local_row_offset * cols;
    tag_rows = (int)noise->local_rows_total;
  }
  long long total = (long long)S * rows * cols;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)ctx->sm_count * 16) blocks = (long long)ctx->sm_count * 16;
  SCDEF_LAUNCH(k_noise_fill, (unsigned)blocks, 256, 0, (cudaStream_t)cuda_stream, out, S, rows * cols, make_key(noise),
                                                                          make_tag(noise, blk, tag_rows, cols), e0);
  return check_launch(ctx, "k_noise_fill");
}

}  // extern "C"
