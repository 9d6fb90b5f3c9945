// Generic per-block kernels: reparameterised sampling, scale priors, W priors, the
// hierarchical z priors, and the final (mu, rho) gradient assembly.
// Math: SURVEY.md Appendix A.1-A.5; reference `_scdef.py:1850-2094`, `jax_utils.py:6-46`.
#pragma once
#include "common.cuh"

namespace scdef {

// -----------------------------------------------------------------------------------------
// K1: v = clip(exp(clip(mu) + exp(clip(rho)) * eps), 1e-15, 1e15) for every block and MC
// sample; the s == 0 slice also accumulates the LogNormal entropies (jax_utils.py:6-18).
// grid = (ceil(max_elems/256), n_blocks)
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sample(const BlockTable tab, int S, PhiloxKey key,
                                                double* loss_out, float global_weight) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const BlockDesc& b = tab.b[blockIdx.y];
  const int per = b.rows * b.cols;
  // one Philox call yields the normals of 4 flat elements; the block's elements are [e0, e0+per) of the stream
  const long long e0 = b.noise_e0;
  const long long q0 = e0 >> 2;
  const int nquad = (int)(((e0 + per + 3) >> 2) - q0);
  const long long total = (long long)S * nquad;
  double ent = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(e / nquad);
    const int qd = (int)(e - (long long)s * nquad);
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!b.eps && !b.fixed) n4 = philox_normal4(key, b.tag, (uint32_t)s, (uint32_t)(q0 + qd));
    const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long gi = (q0 + qd) * 4 + q - e0;
      if (gi < 0 || gi >= per) continue;
      const int i = (int)gi;
      const int r = i / b.cols, c = i - r * b.cols;
      const long long row = b.gather ? b.gather[r] : r;
      float v;
      if (b.fixed) {
        v = b.fixed[row * b.cols + c];
      } else {
        const float mu = b.mu[row * b.ld + b.col0 + c];
        const float rho = b.rho[row * b.ld + b.col0 + c];
        const float mut = fminf(fmaxf(mu, LOG_1E_10), b.mu_hi);
        const float rhot = fminf(fmaxf(rho, LOG_1E_10), LOG_1E10);
        const float eps = b.eps ? b.eps[((long long)s * b.rows + r) * b.eps_ld + b.eps_col0 + c] : nn[q];
        v = fminf(fmaxf(expf(mut + expf(rhot) * eps), V_LO), V_HI);
        if (s == 0) ent += (double)(mut + rhot + HALF_LOG_2PI_PLUS_HALF);
      }
      b.smp[(long long)s * per + i] = v;
    }
  }
  // entropy value: ent_weight already holds w_H * scale_H (* global_weight for global blocks)
  double tot = block_sum_d(ent, red);
  if (threadIdx.x == 0 && tot != 0.0 && b.active) {
    bool is_local = b.gather != nullptr;
    atomicAdd(&loss_out[is_local ? 0 : 1], -tot * (double)b.ent_weight);
  }
}

// -----------------------------------------------------------------------------------------
// k_sample, second version (the default; SCDEF_SAMPLE_V2=0 selects the first; same draws, same arithmetic, bit-identical
// samples: `tests/test_variants_gpu.py`).  k_sample launches ceil(max_elems/4/256) CTAs for EVERY block of the table, so
// at C5 eleven of the thirteen block rows are thousands of CTAs without work, and every (sample, quad) thread redoes the
// gather, two integer divisions and an expf(rho) per element.  Here the grid is one dimension sized per block
// (`SampleGrid.cta0`), a thread owns one quad (4 consecutive elements) for a CHUNK of the S samples and keeps
// (clip(mu), sigma) of its elements in registers.
// -----------------------------------------------------------------------------------------
struct SampleGrid {
  int cta0[MAX_BLOCKS + 1];   // first CTA of each table block; cta0[n] = total
  int sch[MAX_BLOCKS];        // sample chunks per quad of the block
};

__global__ void __launch_bounds__(256) k_sample_v2(const BlockTable tab, const SampleGrid sg, int S, PhiloxKey key,
                                                   double* loss_out, float global_weight) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  (void)global_weight;
  __shared__ double red[32];
  int bi = 0;
  while (bi + 1 < tab.n && (int)blockIdx.x >= sg.cta0[bi + 1]) ++bi;
  const BlockDesc& b = tab.b[bi];
  const int per = b.rows * b.cols;
  const long long e0 = b.noise_e0;
  const long long q0 = e0 >> 2;
  const int nquad = (int)(((e0 + per + 3) >> 2) - q0);
  const int sch = sg.sch[bi];
  const long long t = (long long)((int)blockIdx.x - sg.cta0[bi]) * 256 + threadIdx.x;
  const int chunk = (int)(t / nquad), qd = (int)(t - (long long)chunk * nquad);
  double ent = 0.0;
  if (chunk < sch) {
    const int s_begin = (int)((long long)chunk * S / sch), s_end = (int)((long long)(chunk + 1) * S / sch);
    // the thread's four elements: flat index, clipped mean, sigma (or the pinned value)
    int idx[4];
    float mut[4], sig[4], fixedv[4];
    long long eps_off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long gi = (q0 + qd) * 4 + q - e0;
      idx[q] = -1; mut[q] = 0.f; sig[q] = 0.f; fixedv[q] = 0.f; eps_off[q] = 0;
      if (gi < 0 || gi >= per) continue;
      const int i = (int)gi;
      const int r = i / b.cols, c = i - r * b.cols;
      const long long row = b.gather ? b.gather[r] : r;
      idx[q] = i;
      if (b.fixed) {
        fixedv[q] = b.fixed[row * b.cols + c];
      } else {
        const float mu = b.mu[row * b.ld + b.col0 + c];
        const float rho = b.rho[row * b.ld + b.col0 + c];
        mut[q] = fminf(fmaxf(mu, LOG_1E_10), b.mu_hi);
        const float rhot = fminf(fmaxf(rho, LOG_1E_10), LOG_1E10);
        sig[q] = expf(rhot);
        eps_off[q] = (long long)r * b.eps_ld + b.eps_col0 + c;
        if (chunk == 0) ent += (double)(mut[q] + rhot + HALF_LOG_2PI_PLUS_HALF);   // the s == 0 slice carries the entropy
      }
    }
    for (int s = s_begin; s < s_end; ++s) {
      float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!b.eps && !b.fixed) n4 = philox_normal4(key, b.tag, (uint32_t)s, (uint32_t)(q0 + qd));
      const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (idx[q] < 0) continue;
        float v;
        if (b.fixed) {
          v = fixedv[q];
        } else {
          const float eps = b.eps ? b.eps[(long long)s * b.rows * b.eps_ld + eps_off[q]] : nn[q];
          v = fminf(fmaxf(expf(mut[q] + sig[q] * eps), V_LO), V_HI);
        }
        b.smp[(long long)s * per + idx[q]] = v;
      }
    }
  }
  double tot = block_sum_d(ent, red);
  if (threadIdx.x == 0 && tot != 0.0 && b.active) {
    bool is_local = b.gather != nullptr;
    atomicAdd(&loss_out[is_local ? 0 : 1], -tot * (double)b.ent_weight);
  }
}

// -----------------------------------------------------------------------------------------
// Final assembly (A.1): d loss/d mu = -1[mu in clip] ( (1/S) sum_s 1[v in clip] G v + w_H ),
//                       d loss/d rho = -1[rho in clip]( (1/S) sum_s 1[..] G v eps sigma + w_H ).
// grid = (ceil(max_elems/256), n_blocks); only blocks of the requested pass are in `tab`.
// -----------------------------------------------------------------------------------------
constexpr int FIN_SPLIT = 8;  // warps of a CTA that share 32 elements and split their loop over the MC samples

// CTA = 32 consecutive elements (lanes: every load is one coalesced 128 B line of an MC-sample plane) x 8 warps that
// take the samples s = warp, warp + 8, ...; the eight partial sums meet in shared memory.  (The first version gave an
// element to 8 adjacent lanes: 16 B per sample plane and instruction, half of every sector unused — 30 us per launch
// at C5 for 41 MB.)
__global__ void __launch_bounds__(256) k_finish(const BlockTable tab, int S, PhiloxKey key) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  (void)key;
  __shared__ float part[2][FIN_SPLIT][32];
  const BlockDesc& b = tab.b[blockIdx.y];
  const int per = b.rows * b.cols;
  const float invS = 1.f / (float)S;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_groups = (per + 31) / 32;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int i = grp * 32 + lane;
    const bool live = i < per;
    const int r = live ? i / b.cols : 0, c = live ? i - r * b.cols : 0;
    const bool on = live && b.active && !b.stopped;
    float mu = 0.f, rho = 0.f, gv = 0.f, gve = 0.f;
    if (on) {
      const long long row = b.gather ? b.gather[r] : r;
      mu = b.mu[row * b.ld + b.col0 + c];
      rho = b.rho[row * b.ld + b.col0 + c];
      const float mut = fminf(fmaxf(mu, LOG_1E_10), b.mu_hi);
      for (int s = w; s < S; s += FIN_SPLIT) {
        const long long e = (long long)s * per + i;
        const float v = b.smp[e];
        if (v > V_LO && v < V_HI) {
          // inside the sample clip v = exp(mu~ + sigma eps)  =>  sigma eps = log v - mu~ : no noise needed here
          const float t = b.G[e] * v;
          gv += t;
          gve = fmaf(t, logf(v) - mut, gve);
        }
      }
    }
    part[0][w][lane] = gv;
    part[1][w][lane] = gve;
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int k = 1; k < FIN_SPLIT; ++k) { gv += part[0][k][lane]; gve += part[1][k][lane]; }
      float dmu = 0.f, drho = 0.f;
      if (on) {
        const bool in_mu = (mu >= LOG_1E_10) && (mu <= b.mu_hi);
        const bool in_rho = (rho >= LOG_1E_10) && (rho <= LOG_1E10);
        dmu = in_mu ? -(gv * invS * b.g_weight + b.ent_weight) : 0.f;
        drho = in_rho ? -(gve * invS * b.g_weight + b.ent_weight) : 0.f;
      }
      if (live) {
        b.gmu[(long long)r * b.g_ld + b.g_col0 + c] = dmu;
        b.grho[(long long)r * b.g_ld + b.g_col0 + c] = drho;
      }
    }
    __syncthreads();
  }
}

// -----------------------------------------------------------------------------------------
// Scale priors (A.4 last lines; `_scdef.py:1946-1990`): Gamma(shape a, rate b_e) on the
// samples of cell_scale, gene_scale, brd, ard, alpha.  STORES G (first writer of these buffers).
// -----------------------------------------------------------------------------------------
struct SimplePrior {
  const float* smp;
  float* G;          // nullptr -> value only
  const float* rate; // per-element rate factor (gathered through `gather` by row) or nullptr
  const int64_t* gather;
  int rows, cols;
  float a, lgam_a, rate_scale; // rate = rate_scale * (rate ? rate[...] : 1)
  float weight;      // multiplies value and G: scale (cell part) or global_weight
  int is_cell;       // accumulate into loss_out[0] (cell part) or [1]
  int active;
};
struct SimplePriorTable {
  SimplePrior p[5];
};

__global__ void __launch_bounds__(256) k_scale_priors(const SimplePriorTable tab, int S, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const SimplePrior& p = tab.p[blockIdx.y];
  if (!p.active) return;
  const int per = p.rows * p.cols;
  const long long total = (long long)S * per;
  double val = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int i = (int)(e % per);
    float b = p.rate_scale;
    if (p.rate) {
      int r = i / p.cols, c = i - r * p.cols;
      long long row = p.gather ? p.gather[r] : r;
      b *= p.rate[row * p.cols + c];
    }
    float v = p.smp[e];
    val += (double)gamma_lp(v, logf(v), p.a, b, p.lgam_a, logf(b));
    if (p.G) p.G[e] = p.weight * ((p.a - 1.f) / v - b);
  }
  double tot = block_sum_d(val, red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[p.is_cell ? 0 : 1], -tot * (double)p.weight / (double)S);
}

// -----------------------------------------------------------------------------------------
// W priors (A.4; `_scdef.py:2014-2044`).  One warp per (s, row k).  STORES G_W; atomically
// adds the brd/ard terms.  mode 0: A=P, B=R*K_l.  mode 1 (l=0, BRD): A=P/tau_k,
// B=R*K0/(tau_k*omega_k).  mode 2 (l=1, BRD): A=P, B=R*K1/tau_col.
// -----------------------------------------------------------------------------------------
struct WPriorArgs {
  const float* smp_W;  // (S, rows, cols)
  float* G_W;          // (S, rows, cols) or nullptr (value only)
  const float* P;      // (rows, cols) or nullptr
  const float* R;
  float P_scalar, R_scalar;
  const float* smp_tau; // (S, K0)
  const float* smp_om;  // (S, K0)
  float* G_tau;         // (S, K0) or nullptr
  float* G_om;
  int rows, cols, mode;
  float K_l;
  float weight;         // global_weight
};

__global__ void __launch_bounds__(256) k_wprior(const WPriorArgs a, int S, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nrows = S * a.rows;
  double val = 0.0;
  if (warp < nrows) {
    const int s = warp / a.rows, k = warp - s * a.rows;
    const float* w_row = a.smp_W + ((long long)s * a.rows + k) * a.cols;
    float* g_row = a.G_W ? a.G_W + ((long long)s * a.rows + k) * a.cols : nullptr;
    float tau = 1.f, om = 1.f;
    if (a.mode == 1) {
      tau = a.smp_tau[s * a.rows + k];
      om = a.smp_om[s * a.rows + k];
    }
    const bool scalar_P = (a.P == nullptr);
    float A_row = 0.f, lgam_row = 0.f, psi_row = 0.f;
    if (scalar_P) {
      A_row = (a.mode == 1) ? a.P_scalar / tau : a.P_scalar;
      lgam_row = lgammaf(A_row);
      if (a.mode == 1) psi_row = digammaf(A_row);
    }
    float acc_tau = 0.f, acc_om = 0.f;
    for (int c = lane; c < a.cols; c += 32) {
      float w = w_row[c];
      float lw = logf(w);
      float P = scalar_P ? a.P_scalar : a.P[k * a.cols + c];
      float R = a.R ? a.R[k * a.cols + c] : a.R_scalar;
      float A, Bm, lg;
      if (a.mode == 1) {
        A = scalar_P ? A_row : P / tau;
        Bm = R * a.K_l / (tau * om);
        lg = scalar_P ? lgam_row : lgammaf(A);
        float lB = logf(Bm);
        if (g_row) {
          float psi = scalar_P ? psi_row : digammaf(A);
          float dA = lw - psi + lB;
          float dB = -w + A / Bm;
          acc_tau += dA * (-A / tau) + dB * (-Bm / tau);
          acc_om += dB * (-Bm / om);
        }
        val += (double)gamma_lp(w, lw, A, Bm, lg, lB);
      } else {
        A = P;
        lg = scalar_P ? lgam_row : lgammaf(A);
        if (a.mode == 2) {
          float tc = a.smp_tau[s * a.cols + c];
          Bm = R * a.K_l / tc;
          if (g_row) atomicAdd(&a.G_tau[s * a.cols + c], a.weight * (-w + A / Bm) * (-Bm / tc));
        } else {
          Bm = R * a.K_l;
        }
        val += (double)gamma_lp(w, lw, A, Bm, lg, logf(Bm));
      }
      if (g_row) g_row[c] = a.weight * ((A - 1.f) / w - Bm);
    }
    if (a.mode == 1 && g_row) {
      acc_tau = warp_sum(acc_tau);
      acc_om = warp_sum(acc_om);
      if (lane == 0) {
        atomicAdd(&a.G_tau[s * a.rows + k], a.weight * acc_tau);
        atomicAdd(&a.G_om[s * a.rows + k], a.weight * acc_om);
      }
    }
  }
  double tot = block_sum_d(val, red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[1], -tot * (double)a.weight / (double)S);
}

// -----------------------------------------------------------------------------------------
// K5: hierarchical z priors (A.4; `_scdef.py:2047-2094`).  grid = (cell chunks, S), 8 warps,
// one warp per cell.  W_l (l>=1) samples of this MC sample live in shared memory (row stride
// padded to an odd number of words so that both access directions are conflict-free).
// LOCAL pass: stores G_z (all layers).  GLOBAL pass: accumulates dW_l = z_l^T P_{l-1} in shared
// memory and adds it to G_W; adds the alpha gradient when alpha is marginalised.
// -----------------------------------------------------------------------------------------
struct HierArgs {
  int L;
  int K[SCDEF_MAX_LAYERS];
  int off[SCDEF_MAX_LAYERS + 1];
  int wld[SCDEF_MAX_LAYERS];   // padded row stride of W_l in smem (l >= 1)
  int woff[SCDEF_MAX_LAYERS];  // offset of W_l in the smem W area (l >= 1)
  int wtot;                    // floats in the smem W area
  const float* smp_z[SCDEF_MAX_LAYERS];  // (S,B,K_l)
  const float* smp_W[SCDEF_MAX_LAYERS];  // (S,K_l,K_{l-1}), l>=1
  float* G_z[SCDEF_MAX_LAYERS];          // LOCAL
  float* G_W[SCDEF_MAX_LAYERS];          // GLOBAL (l>=1)
  const float* smp_a;  // (S) or nullptr
  float* G_a;          // (S) GLOBAL & marginalize
  const float* caf;    // (N) cell_alpha_factor
  const int64_t* idx;  // (B)
  int B, pass, cells_per_block, top_layer; // top_layer = index of the "active top" layer
  float alpha, alpha_floor, top_alpha, lgam_top, scale;
};

constexpr int HIER_WARPS = 8;

__global__ void __launch_bounds__(HIER_WARPS * 32) k_hier(const HierArgs a, int S, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ float sm[];
  __shared__ double red[32];
  const int Kt = a.off[a.L];
  float* Wsm = sm;                       // [wtot]
  float* GWsm = Wsm + a.wtot;            // [wtot] (GLOBAL only, else unused but allocated)
  float* zs = GWsm + a.wtot;             // [HIER_WARPS][Kt]
  float* Ps = zs + HIER_WARPS * Kt;      // [HIER_WARPS][Kt]  P_l at off[l]
  float* gzs = Ps + HIER_WARPS * Kt;     // [HIER_WARPS][Kt]
  const int s = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool global_pass = (a.pass == SCDEF_PASS_GLOBAL);

  for (int l = 1; l < a.L; ++l) {
    const int rows = a.K[l], cols = a.K[l - 1];
    const float* src = a.smp_W[l] + (long long)s * rows * cols;
    for (int e = tid; e < rows * cols; e += blockDim.x) {
      int j = e / cols, k = e - j * cols;
      Wsm[a.woff[l] + j * a.wld[l] + k] = src[e];
    }
  }
  if (global_pass)
    for (int e = tid; e < a.wtot; e += blockDim.x) GWsm[e] = 0.f;
  __syncthreads();

  const float alpha_value = a.smp_a ? a.smp_a[s] : a.alpha;
  const int j_begin = blockIdx.x * a.cells_per_block;
  const int j_end = min(a.B, j_begin + a.cells_per_block);
  double val = 0.0;
  float dal_acc = 0.f;
  float* z = zs + warp * Kt;
  float* P = Ps + warp * Kt;
  float* gz = gzs + warp * Kt;

  for (int j0 = j_begin; j0 < j_end; j0 += HIER_WARPS) {
    const int j = j0 + warp;
    const bool valid = j < j_end;
    if (valid) {
      for (int l = 0; l < a.L; ++l)
        for (int k = lane; k < a.K[l]; k += 32)
          z[a.off[l] + k] = a.smp_z[l][((long long)s * a.B + j) * a.K[l] + k];
      const float caf = a.caf[a.idx[j]];
      const float araw = alpha_value * caf;
      const float al = fmaxf(araw, a.alpha_floor);
      const float lgam_al = lgammaf(al);
      const float log_al = logf(al);
      const float psi_al = a.G_a ? digammaf(al) : 0.f;
      __syncwarp();
      float dal = 0.f;
      for (int l = a.L - 1; l >= 0; --l) {
        const int Kl = a.K[l];
        const bool top = (l == a.top_layer);
        for (int k = lane; k < Kl; k += 32) {
          float zz = z[a.off[l] + k];
          float lz = logf(zz);
          float g;
          if (top) {
            val += (double)gamma_lp(zz, lz, a.top_alpha, a.top_alpha, a.lgam_top, logf(a.top_alpha));
            g = (a.top_alpha - 1.f) / zz - a.top_alpha;
          } else {
            float m = 1.f;
            if (l < a.L - 1) {
              const int Ku = a.K[l + 1];
              const float* Wl = Wsm + a.woff[l + 1];
              const int ld = a.wld[l + 1];
              const float* zu = z + a.off[l + 1];
              m = 0.f;
              for (int u = 0; u < Ku; ++u) m = fmaf(zu[u], Wl[u * ld + k], m);
            }
            float rate = al / m;
            float lrate = log_al - logf(m);
            val += (double)gamma_lp(zz, lz, al, rate, lgam_al, lrate);
            g = (al - 1.f) / zz - rate;
            if (l < a.L - 1) P[a.off[l] + k] = rate * (zz / m - 1.f);
            dal += lz - psi_al + lrate + 1.f - zz / m;
          }
          gz[a.off[l] + k] = g;
        }
      }
      if (a.G_a && araw > a.alpha_floor) dal_acc += dal * caf;
      __syncwarp();
      if (!global_pass) {
        // children -> parents: G_{z_l}[u] += sum_k P_{l-1}[k] W_l[u][k]  (layer l-1 ordinary)
        for (int l = 1; l < a.L; ++l) {
          if (l - 1 == a.top_layer) continue;
          const int Kl = a.K[l], Kc = a.K[l - 1];
          const float* Wl = Wsm + a.woff[l];
          const int ld = a.wld[l];
          const float* Pc = P + a.off[l - 1];
          for (int u = lane; u < Kl; u += 32) {
            float acc = 0.f;
            for (int k = 0; k < Kc; ++k) acc = fmaf(Pc[k], Wl[u * ld + k], acc);
            gz[a.off[l] + u] += acc;
          }
        }
        __syncwarp();
        for (int l = 0; l < a.L; ++l)
          for (int k = lane; k < a.K[l]; k += 32)
            a.G_z[l][((long long)s * a.B + j) * a.K[l] + k] = a.scale * gz[a.off[l] + k];
      }
    }
    if (global_pass) {
      __syncthreads();
      const int nvalid = min(HIER_WARPS, j_end - j0);
      for (int l = 1; l < a.L; ++l) {
        if (l - 1 == a.top_layer) continue;
        const int rows = a.K[l], cols = a.K[l - 1], ld = a.wld[l];
        for (int e = tid; e < rows * cols; e += blockDim.x) {
          int u = e / cols, k = e - u * cols;
          float acc = 0.f;
          for (int w = 0; w < nvalid; ++w) acc = fmaf(zs[w * Kt + a.off[l] + u], Ps[w * Kt + a.off[l - 1] + k], acc);
          GWsm[a.woff[l] + u * ld + k] += acc;
        }
      }
      __syncthreads();
    }
  }

  if (global_pass) {
    for (int l = 1; l < a.L; ++l) {
      if (l - 1 == a.top_layer) continue;
      const int rows = a.K[l], cols = a.K[l - 1], ld = a.wld[l];
      float* dst = a.G_W[l] + (long long)s * rows * cols;
      for (int e = tid; e < rows * cols; e += blockDim.x) {
        int u = e / cols, k = e - u * cols;
        atomicAdd(&dst[e], a.scale * GWsm[a.woff[l] + u * ld + k]);
      }
    }
    if (a.G_a) {
      dal_acc = warp_sum(dal_acc);
      if (lane == 0 && dal_acc != 0.f) atomicAdd(&a.G_a[s], a.scale * dal_acc);
    }
  }
  double tot = block_sum_d(val, red);
  if (tid == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)S);
}

// -----------------------------------------------------------------------------------------
// K5, register-tiled version (used whenever its shared-memory plan fits).  Same math as k_hier.
// CTA = (MC sample s, a run of 32-cell tiles); lane = cell, so every per-layer matrix product
//   m_l = z_{l+1} W_{l+1},   d z_{l+1} += P_l W_{l+1}^T,   d W_{l+1} += z_{l+1}^T P_l
// is a small GEMM on smem tiles stored [k][cell] (row stride 33: conflict-free with lanes on the
// cells AND with lanes on k), with 4x register blocking and LDS.128 on the W rows.
// Layers are processed top-down; only the gradient buffers of two adjacent layers are live.
// -----------------------------------------------------------------------------------------
struct HierTArgs {
  HierArgs h;
  int tiles_per_cta;
  int wld[SCDEF_MAX_LAYERS];   // row stride of W_l in smem (multiple of 4)
  int woff[SCDEF_MAX_LAYERS];
  int wtot, kmax;
};
constexpr int HT = 32, HLD = 33;
constexpr int HNW = 16, HNT = HNW * 32;   // warps / threads per CTA of k_hier_tiled (2 CTAs per SM: 32 warps)

// ---- warp-level tensor-core form of the three small products (template parameter MMA of k_hier_tiled) ----------------
// mma.sync.m16n8k8 tf32 with both operands split on the fly into tf32 hi + tf32 lo (three products lo*hi + hi*lo + hi*hi:
// ~2^-21 relative, fp32 accumulate), on the SAME shared-memory tiles as the CUDA-core loops.  g = lane >> 2, t = lane & 3:
//   A (16 x 8, row): a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]
//   B ( 8 x 8, col): b0 = B[t][g], b1 = B[t+4][g]
//   C (16 x 8)     : c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
// (`tests/test_host.py::test_hier_mma_fragment_maps` replays these index maps in NumPy against a plain product.)
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// C[16 x 8] += A B for one k-step; af / bf return element (row, col) of A / (k, n) of B (0 outside the matrices)
template <typename FA, typename FB>
__device__ __forceinline__ void mma_step_3xtf32(float (&c)[4], int lane, FA af, FB bf) {
  const int g = lane >> 2, t = lane & 3;
  uint32_t ah[4], al[4], bh[2], bl[2];
  tf32_split(af(g, t), ah[0], al[0]);
  tf32_split(af(g + 8, t), ah[1], al[1]);
  tf32_split(af(g, t + 4), ah[2], al[2]);
  tf32_split(af(g + 8, t + 4), ah[3], al[3]);
  tf32_split(bf(t, g), bh[0], bl[0]);
  tf32_split(bf(t + 4, g), bh[1], bl[1]);
  mma_tf32_16x8x8(c, al, bh);
  mma_tf32_16x8x8(c, ah, bl);
  mma_tf32_16x8x8(c, ah, bh);
}

template <bool MMA>
__global__ void __launch_bounds__(HNT, 2) k_hier_tiled(const HierTArgs t, int S, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ __align__(16) float sm[];
  __shared__ double red[32];
  const HierArgs& a = t.h;
  const int L = a.L, Kt = a.off[L];
  const bool global_pass = (a.pass == SCDEF_PASS_GLOBAL);
  float* Wsm = sm;                                   // [wtot]
  float* GWsm = Wsm + t.wtot;                        // [wtot]           GLOBAL only
  float* zt = GWsm + (global_pass ? t.wtot : 0);     // [Kt][HLD]
  float* Pm = zt + Kt * HLD;                         // [kmax][HLD]      m_l, then P_l in place
  float* gzA = Pm + t.kmax * HLD;                    // [kmax][HLD]      LOCAL only: layer l+1
  float* gzB = gzA + t.kmax * HLD;                   // [kmax][HLD]      LOCAL only: layer l
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int l = 1; l < L; ++l) {
    const int rows = a.K[l], cols = a.K[l - 1], ld = t.wld[l];
    const float* src = a.smp_W[l] + (long long)s * rows * cols;
    // one warp per row, lanes along the row: coalesced, no index divisions, several loads in flight
#pragma unroll 4
    for (int u = warp; u < rows; u += HNW) {
      for (int k = lane; k < ld; k += 32) {
        Wsm[t.woff[l] + u * ld + k] = k < cols ? __ldg(src + u * cols + k) : 0.f;
        if (global_pass) GWsm[t.woff[l] + u * ld + k] = 0.f;
      }
    }
  }
  const float alpha_value = a.smp_a ? a.smp_a[s] : a.alpha;
  double val = 0.0;
  float dal_acc = 0.f;

  for (int tile = blockIdx.x * t.tiles_per_cta; tile < (blockIdx.x + 1) * t.tiles_per_cta; ++tile) {
    const int j0 = tile * HT;
    if (j0 >= a.B) break;
    __syncthreads();
    // z samples of the tile, all layers: coalesced over k, stored [k][cell]
    for (int l = 0; l < L; ++l) {
      const int Kl = a.K[l];
      const float* zsrc = a.smp_z[l] + ((long long)s * a.B + j0) * Kl;
      float* zdst = zt + a.off[l] * HLD;
      // one warp per cell, lanes along k: coalesced reads; smem bank = (k + c) % 32, conflict-free
#pragma unroll
      for (int ci = 0; ci < HT / HNW; ++ci) {
        const int c = warp + HNW * ci;
        const bool ok = j0 + c < a.B;
        for (int k = lane; k < Kl; k += 32) zdst[k * HLD + c] = ok ? __ldg(zsrc + c * Kl + k) : 1.f;
      }
    }
    const int jl = j0 + lane;
    const bool cvalid = jl < a.B;
    const float caf = cvalid ? a.caf[a.idx[jl]] : 1.f;
    const float araw = alpha_value * caf;
    const float al = fmaxf(araw, a.alpha_floor);
    const float lgam_al = lgammaf(al), log_al = logf(al);
    const float psi_al = a.G_a ? digammaf(al) : 0.f;
    float dal = 0.f;
    __syncthreads();

    for (int l = L - 1; l >= 0; --l) {
      const int Kl = a.K[l];
      const bool top = (l == a.top_layer);
      const bool has_parent = (l < L - 1) && !top;
      // (1) m_l = z_{l+1} W_{l+1}  -> Pm[k][cell]
      if (has_parent) {
        const int Ku = a.K[l + 1], ld = t.wld[l + 1];
        const float* Wl = Wsm + t.woff[l + 1];
        if (MMA) {
          // m[cell][k] = sum_u z[u][cell] W[u][k]:  M = the 32 cells (2 m-tiles), N = Kl, K = Ku; a warp per (m-tile, n-tile)
          const float* zb = zt + a.off[l + 1] * HLD;
          const int n_nt = (Kl + 7) / 8;
          for (int tile = warp; tile < 2 * n_nt; tile += HNW) {
            const int m0 = (tile & 1) * 16, n0 = (tile >> 1) * 8;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k0 = 0; k0 < Ku; k0 += 8)
              mma_step_3xtf32(c, lane,
                              [&](int r, int kk) { return k0 + kk < Ku ? zb[(k0 + kk) * HLD + m0 + r] : 0.f; },
                              [&](int kk, int n) { return (k0 + kk < Ku && n0 + n < Kl) ? Wl[(k0 + kk) * ld + n0 + n] : 0.f; });
            const int g = lane >> 2, tq = lane & 3;
            if (n0 + 2 * tq < Kl) { Pm[(n0 + 2 * tq) * HLD + m0 + g] = c[0]; Pm[(n0 + 2 * tq) * HLD + m0 + g + 8] = c[2]; }
            if (n0 + 2 * tq + 1 < Kl) { Pm[(n0 + 2 * tq + 1) * HLD + m0 + g] = c[1]; Pm[(n0 + 2 * tq + 1) * HLD + m0 + g + 8] = c[3]; }
          }
        }
        const float* zu = zt + a.off[l + 1] * HLD + lane;
        for (int k0 = warp * 4; !MMA && k0 < Kl; k0 += 4 * HNW) {
          float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll 4
          for (int u = 0; u < Ku; ++u) {
            const float z = zu[u * HLD];
            const float4 w = *reinterpret_cast<const float4*>(Wl + u * ld + k0);
            acc0 = fmaf(z, w.x, acc0); acc1 = fmaf(z, w.y, acc1); acc2 = fmaf(z, w.z, acc2); acc3 = fmaf(z, w.w, acc3);
          }
          Pm[(k0 + 0) * HLD + lane] = acc0;
          if (k0 + 1 < Kl) Pm[(k0 + 1) * HLD + lane] = acc1;
          if (k0 + 2 < Kl) Pm[(k0 + 2) * HLD + lane] = acc2;
          if (k0 + 3 < Kl) Pm[(k0 + 3) * HLD + lane] = acc3;
        }
        __syncthreads();
      }
      // (2) prior value, own gradient, P_l (in place over m_l)
      float* gz_cur = (l & 1) ? gzA : gzB;   // LOCAL: layers alternate between the two buffers
      for (int k = warp; k < Kl; k += HNW) {
        const float zz = zt[(a.off[l] + k) * HLD + lane];
        const float lz = __logf(zz);   // MUFU path (as in the likelihood epilogue): <= 2 ulp, far inside the 1e-4 bar
        float g;
        if (top) {
          if (cvalid) val += (double)gamma_lp(zz, lz, a.top_alpha, a.top_alpha, a.lgam_top, logf(a.top_alpha));
          g = (a.top_alpha - 1.f) / zz - a.top_alpha;
        } else {
          const float m = has_parent ? Pm[k * HLD + lane] : 1.f;
          const float inv_m = __frcp_rn(m);                // one reciprocal serves rate and z/m
          const float rate = al * inv_m, lrate = log_al - __logf(m);
          const float zm = zz * inv_m;
          if (cvalid) val += (double)gamma_lp(zz, lz, al, rate, lgam_al, lrate);
          g = __fdividef(al - 1.f, zz) - rate;
          if (has_parent) Pm[k * HLD + lane] = cvalid ? rate * (zm - 1.f) : 0.f;
          if (cvalid) dal += lz - psi_al + lrate + 1.f - zm;
        }
        if (!global_pass) gz_cur[k * HLD + lane] = g;
      }
      __syncthreads();
      // (3) children -> parents
      if (has_parent) {
        const int Ku = a.K[l + 1], ld = t.wld[l + 1];
        if (!global_pass) {
          // gz_{l+1}[u][cell] += sum_k P_l[k][cell] W_{l+1}[u][k]
          float* gz_par = ((l + 1) & 1) ? gzA : gzB;
          const float* Wl = Wsm + t.woff[l + 1];
          if (MMA) {
            // gz[u][cell] += sum_k P[k][cell] W[u][k]:  M = cells, N = Ku, K = Kl
            const int n_nt = (Ku + 7) / 8;
            for (int tile = warp; tile < 2 * n_nt; tile += HNW) {
              const int m0 = (tile & 1) * 16, n0 = (tile >> 1) * 8;
              float c[4] = {0.f, 0.f, 0.f, 0.f};
              for (int k0 = 0; k0 < Kl; k0 += 8)
                mma_step_3xtf32(c, lane,
                                [&](int r, int kk) { return k0 + kk < Kl ? Pm[(k0 + kk) * HLD + m0 + r] : 0.f; },
                                [&](int kk, int n) { return (k0 + kk < Kl && n0 + n < Ku) ? Wl[(n0 + n) * ld + k0 + kk] : 0.f; });
              const int g = lane >> 2, tq = lane & 3;
              if (n0 + 2 * tq < Ku) { gz_par[(n0 + 2 * tq) * HLD + m0 + g] += c[0]; gz_par[(n0 + 2 * tq) * HLD + m0 + g + 8] += c[2]; }
              if (n0 + 2 * tq + 1 < Ku) { gz_par[(n0 + 2 * tq + 1) * HLD + m0 + g] += c[1]; gz_par[(n0 + 2 * tq + 1) * HLD + m0 + g + 8] += c[3]; }
            }
          }
          for (int u0 = warp * 4; !MMA && u0 < Ku; u0 += 4 * HNW) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k0 = 0; k0 < Kl; k0 += 4) {
              float p[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) p[i] = (k0 + i < Kl) ? Pm[(k0 + i) * HLD + lane] : 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (u0 + q < Ku) {
                  const float4 w = *reinterpret_cast<const float4*>(Wl + (u0 + q) * ld + k0);
                  acc[q] = fmaf(p[0], w.x, fmaf(p[1], w.y, fmaf(p[2], w.z, fmaf(p[3], w.w, acc[q]))));
                }
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (u0 + q < Ku) gz_par[(u0 + q) * HLD + lane] += acc[q];
          }
        } else {
          // GW_{l+1}[u][k] += sum_cell z_{l+1}[u][cell] P_l[k][cell]   (lanes on k)
          float* GWl = GWsm + t.woff[l + 1];
          const float* zu = zt + a.off[l + 1] * HLD;
          if (MMA) {
            // GW[u][k] += sum_cell z[u][cell] P[k][cell]:  M = Ku (m-tiles of 16), N = Kl, K = the 32 cells
            const int n_mt = (Ku + 15) / 16, n_nt = (Kl + 7) / 8;
            for (int tile = warp; tile < n_mt * n_nt; tile += HNW) {
              const int m0 = (tile % n_mt) * 16, n0 = (tile / n_mt) * 8;
              float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int k0 = 0; k0 < HT; k0 += 8)
                mma_step_3xtf32(c, lane,
                                [&](int r, int kk) { return m0 + r < Ku ? zu[(m0 + r) * HLD + k0 + kk] : 0.f; },
                                [&](int kk, int n) { return n0 + n < Kl ? Pm[(n0 + n) * HLD + k0 + kk] : 0.f; });
              const int g = lane >> 2, tq = lane & 3;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int u = m0 + g + (i >> 1) * 8, k = n0 + 2 * tq + (i & 1);
                if (u < Ku && k < Kl) GWl[u * ld + k] += c[i];
              }
            }
          }
          for (int u0 = warp * 4; !MMA && u0 < Ku; u0 += 4 * HNW) {
            float acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
            const int nkb = (Kl + 31) / 32;
            for (int c = 0; c < HT; ++c) {
              float z4[4], p4[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) z4[q] = (u0 + q < Ku) ? zu[(u0 + q) * HLD + c] : 0.f;
#pragma unroll
              for (int i = 0; i < 4; ++i) p4[i] = (i < nkb && lane + 32 * i < Kl) ? Pm[(lane + 32 * i) * HLD + c] : 0.f;
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[q][i] = fmaf(z4[q], p4[i], acc[q][i]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (u0 + q < Ku && lane + 32 * i < Kl) GWl[(u0 + q) * ld + lane + 32 * i] += acc[q][i];
          }
        }
        __syncthreads();
      }
      // (4) LOCAL: the gradient of layer l+1 is complete -> global (coalesced over k); layer 0 / childless layers at once
      if (!global_pass) {
        for (int which = 0; which < 2; ++which) {
          int lw;
          if (which == 0) { if (l == L - 1) continue; lw = l + 1; }     // parent layer: complete after this iteration
          else { if (l != 0) continue; lw = 0; }                          // bottom layer: no children
          const float* gsrc = (lw & 1) ? gzA : gzB;
          const int Kw = a.K[lw];
          float* gdst = a.G_z[lw] + ((long long)s * a.B + j0) * Kw;
#pragma unroll
          for (int ci = 0; ci < HT / HNW; ++ci) {
            const int c = warp + HNW * ci;
            if (j0 + c < a.B)
              for (int k = lane; k < Kw; k += 32) gdst[c * Kw + k] = a.scale * gsrc[k * HLD + c];
          }
        }
        __syncthreads();
      }
    }
    if (a.G_a && cvalid && araw > a.alpha_floor) dal_acc += dal * caf;
  }

  if (global_pass) {
    __syncthreads();
    for (int l = 1; l < L; ++l) {
      if (l - 1 == a.top_layer) continue;
      const int rows = a.K[l], cols = a.K[l - 1], ld = t.wld[l];
      float* dst = a.G_W[l] + (long long)s * rows * cols;
      for (int u = warp; u < rows; u += HNW)
        for (int k = lane; k < cols; k += 32) atomicAdd(&dst[u * cols + k], a.scale * GWsm[t.woff[l] + u * ld + k]);
    }
    if (a.G_a) {
      dal_acc = warp_sum(dal_acc);
      if (lane == 0 && dal_acc != 0.f) atomicAdd(&a.G_a[s], a.scale * dal_acc);
    }
  }
  double tot = block_sum_d(val, red);
  if (tid == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)S);
}

}  // namespace scdef
