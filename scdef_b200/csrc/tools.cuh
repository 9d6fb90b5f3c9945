// Post-fit Monte-Carlo assignment of cells to layers: the per-layer statistics of `scdef.tl.assign_confident`
// (reference `tools/factor.py:1264-1379`), one warp per cell.
//
// The reference draws S reparameterised samples z^(s) = exp(mu + sigma eps) of the kept factors of a layer for every
// cell, normalises them to zhat, and reduces over the S samples twice (two lax.scan passes with the same keys):
//   pass 1: argmax counts and the mean of zhat               -> winner f* = argmax_k mean zhat, p = counts / S
//   pass 2: gap^(s) = zhat[f*] - max_{g != f*} zhat[g]        -> lower (1 - credible_level) quantile, mean, SD of the gap
// Here both passes run in one kernel: the lane that owns factor k keeps its count and zhat sum in registers, the
// draws of pass 2 are regenerated from the same Philox counters (or re-read from the injected eps), the S gaps of the
// cell live in the warp's slice of shared memory, and the two order statistics of the quantile are found by rank
// counting.  HBM traffic is the 2 K parameters of the cell in and 7 numbers out; the kernel is bound by the
// FP32 / MUFU work of 2 S K samples per cell.
#pragma once
#include "common.cuh"

namespace scdef {

struct ConfidentArgs {
  const float* mu;      // element (n, c) at mu[n * ld + c]
  const float* rho;
  const int32_t* cols;  // (K) column of kept factor k
  const float* eps;     // explicit draws (S, N, K) or nullptr -> Philox
  float* stats;         // (N, 6): effect_size, posterior_sd, confidence, winner_mass, winner_probability, entropy_confidence
  int32_t* argmax;      // (N) slot of the winner among the kept factors
  long long N, row0;    // row0: global index of row 0 (data parallelism: the draws of a cell do not depend on the sharding)
  int ld, K, S;
  float q_level;        // 1 - credible_level
  PhiloxKey key;
  NoiseTag tag;
};

// (value, index) arg-max over the warp, first maximum on ties (np / jnp.argmax)
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
  }
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// MAXQ: quads (4 consecutive kept factors) per lane; lane owns quads lane, lane + 32, ...
template <int MAXQ>
__global__ void __launch_bounds__(256) k_assign_confident(const ConfidentArgs a) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ float gaps_all[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* gaps = gaps_all + (size_t)wib * a.S;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int K = a.K, S = a.S, Q = (K + 3) >> 2;
  const float invS = 1.f / (float)S;

  for (long long n = warp0; n < a.N; n += nwarps) {
    float mu[MAXQ][4], sg[MAXQ][4], sz[MAXQ][4];
    int cnt[MAXQ][4];
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = 4 * (lane + 32 * i) + c;
        mu[i][c] = 0.f; sg[i][c] = 0.f; sz[i][c] = 0.f; cnt[i][c] = 0;
        if (k < K) {
          const int col = a.cols[k];
          mu[i][c] = a.mu[n * a.ld + col];
          sg[i][c] = expf(a.rho[n * a.ld + col]);                 // factor.py:1265
        }
      }
    // zhat of sample s for the factors this lane owns (0 for k >= K); identical in both passes
    auto sample = [&](int s, float (&zh)[MAXQ][4]) {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i) {
        const int q = lane + 32 * i;
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (q < Q) {
          if (a.eps) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (4 * q + c < K) e[c] = a.eps[((long long)s * a.N + n) * K + 4 * q + c];
          } else {
            const unsigned long long qi = (unsigned long long)(a.row0 + n) * (unsigned long long)Q + (unsigned long long)q;
            NoiseTag t = a.tag;
            t.b += (uint32_t)(qi >> 32);
            const float4 n4 = philox_normal4(a.key, t, (uint32_t)s, (uint32_t)qi);
            e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool on = 4 * q + c < K;
          zh[i][c] = on ? expf(fmaf(sg[i][c], e[c], mu[i][c])) : 0.f;   // 1289
          part += zh[i][c];
        }
      }
      const float denom = fmaxf(warp_sum(part), 1e-30f);                // 1290
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) zh[i][c] = zh[i][c] / denom;
    };

    // ---- pass 1: argmax counts and the sum of zhat (factor.py:1286-1302) ----
    for (int s = 0; s < S; ++s) {
      float zh[MAXQ][4];
      sample(s, zh);
      float best = -INFINITY;
      int bi = 0x7fffffff;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = 4 * (lane + 32 * i) + c;
          if (k < K) {
            sz[i][c] += zh[i][c];
            if (zh[i][c] > best) { best = zh[i][c]; bi = k; }   // ascending k within the lane: first maximum kept
          }
        }
      warp_argmax(best, bi);
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) cnt[i][c] += (4 * (lane + 32 * i) + c == bi) ? 1 : 0;
    }
    // winner = argmax of the mean zhat (1304-1305); p = counts / S (1303): its maximum and entropy (1361-1367)
    float wbest = -INFINITY, pmax = 0.f, H = 0.f;
    int winner = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < MAXQ; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = 4 * (lane + 32 * i) + c;
        if (k < K) {
          const float m = sz[i][c] * invS;
          if (m > wbest) { wbest = m; winner = k; }
          const float p = (float)cnt[i][c] * invS;
          pmax = fmaxf(pmax, p);
          H -= p * logf(fminf(fmaxf(p, 1e-30f), 1.f));
        }
      }
    warp_argmax(wbest, winner);
    pmax = warp_max(pmax);
    H = warp_sum(H);

    // ---- pass 2: the gap between the winner and its runner-up, per sample (1317-1329) ----
    float sum_gap = 0.f, sum_wm = 0.f;
    for (int s = 0; s < S; ++s) {
      float zh[MAXQ][4];
      sample(s, zh);
      float wm = -INFINITY, ru = -INFINITY;
#pragma unroll
      for (int i = 0; i < MAXQ; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = 4 * (lane + 32 * i) + c;
          if (k < K) {
            if (k == winner) wm = zh[i][c];
            else ru = fmaxf(ru, zh[i][c]);
          }
        }
      wm = warp_max(wm);
      ru = warp_max(ru);
      const float gap = wm - ru;
      if (lane == 0) gaps[s] = gap;
      sum_gap += gap;
      sum_wm += wm;
    }
    __syncwarp();
    const float mean_gap = sum_gap * invS;
    float var = 0.f;
    for (int i = lane; i < S; i += 32) {
      const float d = gaps[i] - mean_gap;
      var = fmaf(d, d, var);
    }
    var = warp_sum(var) * invS;                                        // population SD (jnp.std), 1336
    // empirical quantile with linear interpolation (jnp.quantile default), 1334: order statistics by rank counting
    const float pos = a.q_level * (float)(S - 1);
    const int lo = (int)floorf(pos), hi = min((int)ceilf(pos), S - 1);
    const float wh = pos - (float)lo;
    float vlo = -INFINITY, vhi = -INFINITY;
    for (int i = lane; i < S; i += 32) {
      const float gi = gaps[i];
      int rank = 0;
      for (int j = 0; j < S; ++j) {
        const float gj = gaps[j];
        rank += (gj < gi || (gj == gi && j < i)) ? 1 : 0;
      }
      if (rank == lo) vlo = gi;
      if (rank == hi) vhi = gi;
    }
    vlo = warp_max(vlo);
    vhi = warp_max(vhi);
    const float conf = vlo * (1.f - wh) + vhi * wh;
    if (lane == 0) {
      float* o = a.stats + n * 6;
      o[0] = fminf(fmaxf(mean_gap, -1.f), 1.f);                        // 1343
      o[1] = fminf(fmaxf(sqrtf(var), 0.f), 1.f);                       // 1351
      o[2] = fminf(fmaxf(conf, 0.f), 1.f);                             // 1354
      o[3] = fminf(fmaxf(sum_wm * invS, 0.f), 1.f);                    // 1358
      o[4] = fminf(fmaxf(pmax, 0.f), 1.f);                             // 1361
      o[5] = fminf(fmaxf(1.f - H / logf((float)K), 0.f), 1.f);         // 1367
      a.argmax[n] = winner;
    }
    __syncwarp();   // the warp's gap slice is reused by its next cell
  }
}

// =====================================================================================================================
// Cold start of a LogNormal block from Gamma draws on the device (`init_var_params`, reference
// `models/_scdef.py:1639-1648` for z: mean m = clip(Gamma(a, 1) / a), variance m / var_div, moment-matched to (mu, log sigma);
// SURVEY.md §8 f3).  The reference draws with JAX keys, the host path of this package with NumPy's Generator: parity is at
// the level of the distribution only, so the draws here come from the package's Philox stream (Marsaglia-Tsang
// squeeze, boosted by U^(1/a) for a < 1).  One thread per element; the draw of element (row0 + n, k) does not depend
// on how the cells are sharded.
// =====================================================================================================================
struct InitGammaArgs {
  float* mu;     // element (n, k) at mu[n * ld + k]
  float* rho;
  long long N, row0;
  int K, ld;
  float shape, clip_lo, clip_hi, var_div;
  PhiloxKey key;
  NoiseTag tag;
};

// host + device: the CPU test draws from the same code (`scdef_debug_gamma_host`)
__host__ __device__ inline float philox_gamma(float a, PhiloxKey key, NoiseTag tag, unsigned long long elem) {
  const float k32 = 2.3283064365386963e-10f;   // 2^-32
  const bool boost = a < 1.f;
  const float aa = boost ? a + 1.f : a;
  const float d = aa - 0.33333333f, c = 1.f / sqrtf(9.f * d);
  NoiseTag t = tag;
  t.b += (uint32_t)(elem >> 32);
  float g = d;   // the mode, should all attempts be rejected (probability < 1e-20)
  float ub = 0.5f;
  for (uint32_t attempt = 0; attempt < 16u; ++attempt) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)elem, t.a, t.b, attempt), key);
    const float u1 = fminf(((float)r.x + 0.5f) * k32, 0.99999994f), u2 = ((float)r.y + 0.5f) * k32;
    const float x = sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);   // N(0,1)
    const float u = fminf(((float)r.z + 0.5f) * k32, 0.99999994f);
    if (attempt == 0u) ub = fminf(((float)r.w + 0.5f) * k32, 0.99999994f);
    float v = 1.f + c * x;
    if (v <= 0.f) continue;
    v = v * v * v;
    if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) { g = d * v; break; }
  }
  if (boost) g *= expf(logf(ub) / a);                                           // Gamma(a) = Gamma(a + 1) U^(1/a)
  return g;
}

__global__ void __launch_bounds__(256) k_init_gamma_block(const InitGammaArgs a) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long total = a.N * a.K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long n = e / a.K;
    const int k = (int)(e - n * a.K);
    const float g = philox_gamma(a.shape, a.key, a.tag, (unsigned long long)(a.row0 + n) * (unsigned long long)a.K + (unsigned long long)k);
    const double m = (double)fminf(fmaxf(g / a.shape, a.clip_lo), a.clip_hi);
    const double v = m / (double)a.var_div;
    // LogNormal with mean m and variance v (`_scdef.py:1496-1505`): mu = log(m^2 / sqrt(m^2 + v)), sigma^2 = log(1 + v / m^2)
    a.mu[n * a.ld + k] = (float)log(m * m / sqrt(m * m + v));
    a.rho[n * a.ld + k] = (float)log(sqrt(log1p(v / (m * m))));
  }
}


// -----------------------------------------------------------------------------------------
// tl.set_confident_signatures on the device (`/root/reference/src/scdef/tools/factor.py:220-363`): Monte-Carlo gene
// scores of the factors of the upper layers,  score^s[f][g] = sum_r path^s[f][r] W_0^s[r][g],  W_0^s = clip(exp(mu +
// sigma eps^s)) — or, with path == nullptr, the row's own draw (layer 0, `_scdef.py:3975-3989`).  The (S, F, G) score
// tensor the reference materialises (`_collect_hierarchy_mc_scores`, 271-291) is never stored: a CTA samples one
// 64-gene tile of W_0^s into shared memory, forms the scores of 32 factors at a time with a register-tiled FP32
// product, and reduces them on the spot to what the statistics need:
//   mode 0 COUNT     counts[f][g]        += [score > tau_f]                      (exceedance confidences, 305-308)
//   mode 1 REFSCORE  ref_score[s][f][i]   = score at gene ref_idx[f][i]          (the signature genes' own scores)
//   mode 2 RANK      rank_count[s][f][i] += #{g : score[g] > ref_score[s][f][i]}  (gene i is in the draw's top-k list
//                                                                                 iff its count is < k; 354-357)
// grid = (gene tiles, draw chunks), 256 threads: thread (fy, gx) owns factors 2 fy, 2 fy + 1 of the chunk and genes
// 4 gx .. 4 gx + 3 of the tile.
// -----------------------------------------------------------------------------------------
struct SigArgs {
  const float* w_mu; const float* w_rho;   // W_0 halves, (K0_all, G) row-major
  const int* rows;                         // [R] rows of W_0 behind the path's columns
  const float* eps;                        // [S][R][G] N(0,1)
  const float* path;                       // [S][F][R] or nullptr (identity, F == R)
  const float* tau;                        // [F]            (mode 0)
  const int* ref_idx;                      // [F][kmax], -1 padded (modes 1, 2)
  float* ref_score;                        // [S][F][kmax]
  int* counts;                             // [F][G]         (mode 0)
  int* rank_count;                         // [S][F][kmax]   (mode 2)
  int R, G, S, F, kmax, mode, draws_per_cta;
};
constexpr int SIG_TG = 64, SIG_FC = 32, SIG_MAX_R = 128, SIG_MAX_F = 256;

__global__ void __launch_bounds__(256) k_signature_scores(const SigArgs a) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ __align__(16) float sig_sm[];
  float* Wt = sig_sm;                                // [R][SIG_TG]
  float* Pt = Wt + SIG_MAX_R * SIG_TG;               // [SIG_FC][R + 1]
  float* Sc = Pt + SIG_FC * (SIG_MAX_R + 1);         // [SIG_FC][SIG_TG + 1]
  const int tid = threadIdx.x, fy = tid >> 4, gx = tid & 15;
  const int g0 = blockIdx.x * SIG_TG;
  const int s_begin = blockIdx.y * a.draws_per_cta, s_end = min(a.S, s_begin + a.draws_per_cta);
  const int n_fc = (a.F + SIG_FC - 1) / SIG_FC;
  int cnt[SIG_MAX_F / SIG_FC][2][4];
#pragma unroll
  for (int c = 0; c < SIG_MAX_F / SIG_FC; ++c)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) cnt[c][i][j] = 0;

  for (int s = s_begin; s < s_end; ++s) {
    __syncthreads();
    // the draw's tile of W_0 (lanes along the genes: coalesced)
    for (int e = tid; e < a.R * SIG_TG; e += 256) {
      const int r = e / SIG_TG, gl = e - r * SIG_TG, g = g0 + gl;
      float w = 0.f;
      if (g < a.G) {
        const long long o = (long long)a.rows[r] * a.G + g;
        const float u = a.w_mu[o] + expf(a.w_rho[o]) * a.eps[((long long)s * a.R + r) * a.G + g];
        w = fminf(fmaxf(expf(u), 1e-15f), 1e15f);
      }
      Wt[r * SIG_TG + gl] = w;
    }
#pragma unroll 1
    for (int c = 0; c < n_fc; ++c) {
      const int f0 = c * SIG_FC;
      __syncthreads();
      if (a.path) {
        for (int e = tid; e < SIG_FC * a.R; e += 256) {
          const int fl = e / a.R, r = e - fl * a.R;
          Pt[fl * (SIG_MAX_R + 1) + r] = (f0 + fl < a.F) ? a.path[((long long)s * a.F + f0 + fl) * a.R + r] : 0.f;
        }
      }
      __syncthreads();
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      if (a.path) {
        const float* p0 = Pt + (2 * fy) * (SIG_MAX_R + 1), * p1 = p0 + (SIG_MAX_R + 1);
        for (int r = 0; r < a.R; ++r) {
          const float4 w = *reinterpret_cast<const float4*>(Wt + r * SIG_TG + 4 * gx);
          const float x0 = p0[r], x1 = p1[r];
          acc[0][0] = fmaf(x0, w.x, acc[0][0]); acc[0][1] = fmaf(x0, w.y, acc[0][1]);
          acc[0][2] = fmaf(x0, w.z, acc[0][2]); acc[0][3] = fmaf(x0, w.w, acc[0][3]);
          acc[1][0] = fmaf(x1, w.x, acc[1][0]); acc[1][1] = fmaf(x1, w.y, acc[1][1]);
          acc[1][2] = fmaf(x1, w.z, acc[1][2]); acc[1][3] = fmaf(x1, w.w, acc[1][3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = f0 + 2 * fy + i;
          if (f < a.F) {
            const float4 w = *reinterpret_cast<const float4*>(Wt + f * SIG_TG + 4 * gx);
            acc[i][0] = w.x; acc[i][1] = w.y; acc[i][2] = w.z; acc[i][3] = w.w;
          }
        }
      }
      if (a.mode == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = f0 + 2 * fy + i;
          const float t = f < a.F ? a.tau[f] : 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const bool hit = f < a.F && g0 + 4 * gx + j < a.G && acc[i][j] > t;
#pragma unroll
            for (int cc = 0; cc < SIG_MAX_F / SIG_FC; ++cc) cnt[cc][i][j] += (cc == c && hit) ? 1 : 0;
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) Sc[(2 * fy + i) * (SIG_TG + 1) + 4 * gx + j] = acc[i][j];
        __syncthreads();
        for (int e = tid; e < SIG_FC * a.kmax; e += 256) {
          const int fl = e / a.kmax, i = e - fl * a.kmax, f = f0 + fl;
          if (f >= a.F) continue;
          const int gi = a.ref_idx[(long long)f * a.kmax + i];
          if (gi < 0) continue;
          const long long o = ((long long)s * a.F + f) * a.kmax + i;
          if (a.mode == 1) {
            if (gi >= g0 && gi < g0 + SIG_TG) a.ref_score[o] = Sc[fl * (SIG_TG + 1) + gi - g0];
          } else {
            const float rs = a.ref_score[o];
            int n = 0;
            const int gmax = min(SIG_TG, a.G - g0);
            for (int g = 0; g < gmax; ++g) n += Sc[fl * (SIG_TG + 1) + g] > rs ? 1 : 0;
            if (n) atomicAdd(&a.rank_count[o], n);
          }
        }
      }
    }
  }
  if (a.mode == 0) {
#pragma unroll
    for (int c = 0; c < SIG_MAX_F / SIG_FC; ++c)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int f = c * SIG_FC + 2 * fy + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = g0 + 4 * gx + j;
          if (f < a.F && g < a.G && cnt[c][i][j]) atomicAdd(&a.counts[(long long)f * a.G + g], cnt[c][i][j]);
        }
      }
  }
}

}  // namespace scdef
