// K8: fused Adam (optax.adam semantics, `_scdef.py:3079-3082, 3112-3116, 3149-3162`) and the
// posterior summaries (`_scdef.py:3394-3476`).  HBM-bound: 24 B/param (read p,m,v; write p,m,v)
// plus the gradient (dense for globals, minibatch-sparse for locals).
#pragma once
#include "common.cuh"

namespace scdef {

struct AdamHyper {
  float lr, b1, b2, eps, bc1, bc2;  // bc = 1 - b^t
  float inv_bc1, inv_bc2;           // reciprocals (the kernels multiply instead of dividing)
};

// One optax-Adam element update.  The operation sequence is pinned with explicit intrinsics (no compiler
// FMA contraction) so that the dense kernel and the lazy replay are bit-identical.
__device__ __forceinline__ void adam_elem(float& p, float& m, float& v, float g, const AdamHyper& h,
                                          bool write_p, float clip_hi) {
  m = __fmaf_rn(__fsub_rn(1.f, h.b1), g, __fmul_rn(h.b1, m));
  v = __fmaf_rn(__fmul_rn(__fsub_rn(1.f, h.b2), g), g, __fmul_rn(h.b2, v));
  const float mh = __fmul_rn(m, h.inv_bc1), vh = __fmul_rn(v, h.inv_bc2);
  const float upd = __fdividef(mh, __fadd_rn(__fsqrt_rn(vh), h.eps));
  const float pn = __fmaf_rn(-h.lr, upd, p);
  if (write_p) p = fminf(pn, clip_hi);  // clip_params lower bounds (-1e10) can never bind after fminf of finite
}

// ---- globals: a table of segments, one per block --------------------------------------------
struct AdamSeg {
  float *p, *m, *v;
  const float* g;
  int n_half;      // elements of one half (mu or rho)
  int clip;        // clip_params applies (mu<=1e4, rho<=10)
  int frozen;      // freeze_w: parameters keep their value, moments still update
};
struct AdamSegTable {
  AdamSeg s[SCDEF_MAX_LAYERS + 4];
  int n;
};

// grid = (blocks, n_segments)
__global__ void __launch_bounds__(256) k_adam_global(const AdamSegTable tab, const AdamHyper h) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const AdamSeg& sg = tab.s[blockIdx.y];
  const int total = 2 * sg.n_half;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    float p = sg.p[i], m = sg.m[i], v = sg.v[i];
    float hi = INFINITY;
    if (sg.clip) hi = (i < sg.n_half) ? 1e4f : 10.f;
    float lo = sg.clip ? -1e10f : -INFINITY;
    adam_elem(p, m, v, sg.g[i], h, !sg.frozen, hi);
    if (!sg.frozen) p = fmaxf(p, lo);
    else if (sg.clip) p = fmaxf(fminf(p, hi), lo);  // clip_params still runs on the restored W
    sg.p[i] = p;
    sg.m[i] = m;
    sg.v[i] = v;
  }
}

// ---- locals: dense over all N rows; gradient rows come from the minibatch via row_slot -------
__global__ void k_set_slots(int32_t* row_slot, const int64_t* idx, int B, int clear) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < B) row_slot[idx[j]] = clear ? -1 : j;
}

struct AdamLocalArgs {
  float *p, *m, *v;      // (2, N, C)
  const float* g;        // (2, B, C) compact
  const int32_t* row_slot;
  long long N;
  int C, B;
  int n_frozen;          // frozen column ranges [lo, hi)
  int frozen_lo[SCDEF_MAX_LAYERS], frozen_hi[SCDEF_MAX_LAYERS];
};

// grid = (blocks, 2 halves); flat vector index over one half (N*C/VEC vectors), two vectors per
// thread per iteration so that six 16-byte loads are in flight before the first dependent use.
template <int VEC, typename IdxT>
__global__ void __launch_bounds__(256) k_adam_local(const AdamLocalArgs a, const AdamHyper h) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const IdxT C = (IdxT)a.C, CV = (IdxT)(a.C / VEC);
  const IdxT nvec = (IdxT)a.N * CV;
  const int hsel = blockIdx.y;
  const long long hoff = (long long)hsel * a.N * a.C;
  float* __restrict__ P = a.p + hoff;
  float* __restrict__ M = a.m + hoff;
  float* __restrict__ V = a.v + hoff;
  const float* __restrict__ G = a.g + (long long)hsel * a.B * a.C;
  const IdxT stride = (IdxT)gridDim.x * blockDim.x;
  for (IdxT t0 = (IdxT)blockIdx.x * blockDim.x + threadIdx.x; t0 < nvec; t0 += 2 * stride) {
    float p[2][VEC], m[2][VEC], v[2][VEC], g[2][VEC];
    IdxT tt[2] = {t0, t0 + stride};
    int c0[2], slot[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      live[u] = tt[u] < nvec;
      if (!live[u]) continue;
      const IdxT i = tt[u] * VEC;
      if (VEC == 4) {
        float4 p4 = *reinterpret_cast<const float4*>(P + i), m4 = *reinterpret_cast<const float4*>(M + i),
               v4 = *reinterpret_cast<const float4*>(V + i);
        p[u][0] = p4.x; p[u][1] = p4.y; p[u][2] = p4.z; p[u][VEC - 1] = p4.w;
        m[u][0] = m4.x; m[u][1] = m4.y; m[u][2] = m4.z; m[u][VEC - 1] = m4.w;
        v[u][0] = v4.x; v[u][1] = v4.y; v[u][2] = v4.z; v[u][VEC - 1] = v4.w;
      } else {
        p[u][0] = P[i]; m[u][0] = M[i]; v[u][0] = V[i];
      }
      const IdxT n = tt[u] / CV;
      c0[u] = (int)((tt[u] - n * CV) * VEC);
      slot[u] = a.row_slot[n];
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!live[u]) continue;
#pragma unroll
      for (int q = 0; q < VEC; ++q) g[u][q] = 0.f;
      if (slot[u] >= 0) {
        const float* gp = G + (long long)slot[u] * C + c0[u];
        if (VEC == 4) {
          float4 g4 = *reinterpret_cast<const float4*>(gp);
          g[u][0] = g4.x; g[u][1] = g4.y; g[u][2] = g4.z; g[u][VEC - 1] = g4.w;
        } else {
          g[u][0] = gp[0];
        }
      }
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        bool frozen = false;
        for (int f = 0; f < a.n_frozen; ++f) frozen |= (c0[u] + q >= a.frozen_lo[f]) && (c0[u] + q < a.frozen_hi[f]);
        adam_elem(p[u][q], m[u][q], v[u][q], g[u][q], h, !frozen, INFINITY);
      }
      const IdxT i = tt[u] * VEC;
      if (VEC == 4) {
        *reinterpret_cast<float4*>(P + i) = make_float4(p[u][0], p[u][1], p[u][2], p[u][VEC - 1]);
        *reinterpret_cast<float4*>(M + i) = make_float4(m[u][0], m[u][1], m[u][2], m[u][VEC - 1]);
        *reinterpret_cast<float4*>(V + i) = make_float4(v[u][0], v[u][1], v[u][2], v[u][VEC - 1]);
      } else {
        P[i] = p[u][0]; M[i] = m[u][0]; V[i] = v[u][0];
      }
    }
  }
}

// ---- lazy (exact) variant of the dense local Adam ------------------------------------------------------
// A row outside the minibatch sees g = 0: m <- b1 m, v <- b2 v, p <- p - lr * m^ / (sqrt(v^) + eps).  Those steps
// depend only on the row's own state and the step number, so they can be replayed in registers when the row is
// next needed instead of streaming all N rows through HBM every step.  The replay executes the SAME fp32
// operations as k_adam_local (same adam_elem, same per-step bias-correction factors from `bc_tab`), so the
// result is bit-identical to the dense kernel.
struct AdamLazyArgs {
  float *p, *m, *v;        // (2, N, C)
  const float* g;          // (2, B, C) compact (UPDATE phase)
  const int64_t* idx;      // (B) rows, or nullptr = all N rows (full catch-up)
  int32_t* row_step;       // (N) last Adam step applied to the row
  const float2* bc_tab;    // [k-1] = (1/(1-b1^k), 1/(1-b2^k)); entries beyond the table are the last one (= 1, 1)
  long long N;
  int C, B, tab_len;
  int monotone;            // 1: |update| is decreasing over zero-gradient steps (enables the constant-p shortcut)
  int target;              // CATCHUP: replay steps row_step+1 .. target;  UPDATE: the step being applied
  int n_frozen;
  int frozen_lo[SCDEF_MAX_LAYERS], frozen_hi[SCDEF_MAX_LAYERS];
};

// one thread per element of the selected rows; grid = (blocks, 2 halves, 2 blocks {cell_scale, z})
struct AdamLazyPair {
  AdamLazyArgs a[2];
};

__global__ void __launch_bounds__(256) k_adam_lazy_catchup(const AdamLazyPair pr, AdamHyper h) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const AdamLazyArgs& a = pr.a[blockIdx.z];
  const long long rows = a.idx ? a.B : a.N;
  const long long total = rows * a.C;
  const long long hoff = (long long)blockIdx.y * a.N * a.C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long j = t / a.C;
    const int c = (int)(t - j * a.C);
    const long long n = a.idx ? a.idx[j] : j;
    const int t0 = a.row_step[n];
    if (t0 >= a.target) continue;
    const long long i = hoff + n * a.C + c;
    float p = a.p[i], m = a.m[i], v = a.v[i];
    if (m == 0.f && v == 0.f) continue;  // never touched since the optimiser was reset: every replayed step is a no-op
    bool frozen = false;
    for (int f = 0; f < a.n_frozen; ++f) frozen |= (c >= a.frozen_lo[f]) && (c < a.frozen_hi[f]);
    int k = t0 + 1;
    // phase 1: full steps, until the update no longer changes p in fp32.  |update| shrinks by >= ~10 % per
    // zero-gradient step (a.monotone: verified on the host from the bias-correction table), so from then on p is
    // constant and only the two moment decays have to be replayed (phase 2); once m is exactly 0 only v (phase 3).
    for (; k <= a.target; ++k) {
      if (m == 0.f) break;
      const float2 bc = a.bc_tab[min(k, a.tab_len) - 1];
      h.inv_bc1 = bc.x;
      h.inv_bc2 = bc.y;
      const float p_before = p;
      adam_elem(p, m, v, 0.f, h, !frozen, INFINITY);
      if (a.monotone && p == p_before) { ++k; break; }
    }
    // phases 2 and 3: with g = 0 the dense kernel's  m = fma(1-b1, 0, b1*m)  and  v = fma((1-b2)*0, 0, b2*v)  are exactly
    // b1*m and b2*v (adding +0 changes nothing but the sign of a zero, fixed below), so a replayed step is two, then one,
    // multiplications.  m stops changing when it is 0 or a denormal that b1*m rounds back to itself: checked every 8 steps.
    int remaining = a.target - k + 1;
    while (remaining > 0) {
      const float m_prev = m;
      const int n = remaining < 8 ? remaining : 8;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (u < n) { m = __fmul_rn(h.b1, m); v = __fmul_rn(h.b2, v); }
      remaining -= n;
      if (m == m_prev) break;
    }
    for (; remaining >= 8; remaining -= 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) v = __fmul_rn(h.b2, v);
    }
    for (; remaining > 0; --remaining) v = __fmul_rn(h.b2, v);
    if (m == 0.f) m = 0.f;   // -0 -> +0, as the dense kernel's fma leaves it
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
  }
}

// applies step `target` with the minibatch gradient to the minibatch rows (already caught up to target-1)
__global__ void __launch_bounds__(256) k_adam_lazy_update(const AdamLazyPair pr, const AdamHyper h) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const AdamLazyArgs& a = pr.a[blockIdx.z];
  const long long total = (long long)a.B * a.C;
  const long long hoff = (long long)blockIdx.y * a.N * a.C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long j = t / a.C;
    const int c = (int)(t - j * a.C);
    const long long n = a.idx[j];
    const long long i = hoff + n * a.C + c;
    float p = a.p[i], m = a.m[i], v = a.v[i];
    bool frozen = false;
    for (int f = 0; f < a.n_frozen; ++f) frozen |= (c >= a.frozen_lo[f]) && (c < a.frozen_hi[f]);
    adam_elem(p, m, v, a.g[((long long)blockIdx.y * a.B + j) * a.C + c], h, !frozen, INFINITY);
    a.p[i] = p; a.m[i] = m; a.v[i] = v;
  }
}

__global__ void k_set_row_step(int32_t* row_step, const int64_t* idx, long long n, int value) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) row_step[idx ? idx[j] : j] = value;
}

// ---- posterior summaries ----------------------------------------------------------------------
struct PostSeg {
  const float* p;   // (2, n)
  float* mean;      // (n) or nullptr
  float* var;       // (n) or nullptr
  long long n;
};
struct PostSegTable {
  PostSeg s[SCDEF_MAX_LAYERS + 6];
  int n;
};

__global__ void __launch_bounds__(256) k_posterior(const PostSegTable tab) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const PostSeg& sg = tab.s[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n;
       i += (long long)gridDim.x * blockDim.x) {
    float mu = sg.p[i], ls = sg.p[sg.n + i];
    float s2 = expf(ls);
    s2 *= s2;
    if (sg.mean) sg.mean[i] = expf(mu + 0.5f * s2);
    if (sg.var) sg.var[i] = (expf(s2) - 1.f) * expf(2.f * mu + s2);
  }
}

// ---- CSR-resident counts: per-cell constants and minibatch densification (one warp per row) ------
__global__ void __launch_bounds__(256) k_csr_row_stats(const long long* indptr, const float* data, long long n_rows,
                                                       float* lib, float* lgam) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  float sl = 0.f, sg = 0.f;
  for (long long p = indptr[row] + lane; p < indptr[row + 1]; p += 32) {
    const float x = data[p];
    sl += x;
    if (x > 1.5f) sg += lgammaf(x + 1.f);
  }
  sl = warp_sum(sl);
  sg = warp_sum(sg);
  if (lane == 0) {
    if (lib) lib[row] = sl;
    if (lgam) lgam[row] = sg;
  }
}

__global__ void __launch_bounds__(256) k_csr_gather_rows(const long long* indptr, const int* indices, const float* data,
                                                         const long long* rows, int n_rows, int G, float* out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= n_rows) return;
  const long long row = rows[j];
  float* dst = out + (long long)j * G;
  for (long long p = indptr[row] + lane; p < indptr[row + 1]; p += 32) dst[indices[p]] = data[p];
}

// ---- hard assignment counts per factor (one warp per cell, shared-memory histogram per CTA) -------
struct AssignArgs {
  const float* mu;    // (N, Kt)
  const float* rho;   // (N, Kt)
  long long N;
  int L, Kt;
  int K[SCDEF_MAX_LAYERS], off[SCDEF_MAX_LAYERS];
};

__global__ void __launch_bounds__(256) k_assignment_counts(const AssignArgs a, unsigned long long* counts) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ unsigned int hist[];
  for (int i = threadIdx.x; i < a.Kt; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp0; row < a.N; row += nwarps) {
    const float* mu = a.mu + row * a.Kt;
    const float* rho = a.rho + row * a.Kt;
    for (int l = 0; l < a.L; ++l) {
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int k = lane; k < a.K[l]; k += 32) {
        float s2 = expf(rho[a.off[l] + k]);
        s2 *= s2;
        const float v = expf(mu[a.off[l] + k] + 0.5f * s2);   // same expression as k_posterior
        if (v > best || bi == 0x7fffffff) { best = v; bi = k; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
      }
      if (lane == 0 && bi != 0x7fffffff) atomicAdd(&hist[a.off[l] + bi], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.Kt; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

// ---- per-cell constants of a resident count matrix ---------------------------------------------
__global__ void __launch_bounds__(256) k_row_stats(const float* X, long long n_rows, int G, float* lib,
                                                   float* lgam) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  float sl = 0.f, sg = 0.f;
  for (int g = lane; g < G; g += 32) {
    float x = X[row * G + g];
    sl += x;
    if (x > 1.5f) sg += lgammaf(x + 1.f);
  }
  sl = warp_sum(sl);
  sg = warp_sum(sg);
  if (lane == 0) {
    if (lib) lib[row] = sl;
    if (lgam) lgam[row] = sg;
  }
}

// ---- counts stored as uint16 (half the HBM of float32: 10 GB instead of 20 GB at 1M x 5000) -------------------------
// per-cell constants, as k_row_stats
__global__ void __launch_bounds__(256) k_u16_row_stats(const uint16_t* X, long long n_rows, int G, float* lib, float* lgam) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  float sl = 0.f, sg = 0.f;
  for (int g = lane; g < G; g += 32) {
    const float x = (float)X[row * G + g];
    sl += x;
    if (x > 1.5f) sg += lgammaf(x + 1.f);
  }
  sl = warp_sum(sl);
  sg = warp_sum(sg);
  if (lane == 0) {
    if (lib) lib[row] = sl;
    if (lgam) lgam[row] = sg;
  }
}
// out[j, :] = float(X[rows[j], :]) : the minibatch as the dense float32 X_batch every kernel downstream reads
__global__ void __launch_bounds__(256) k_u16_gather_rows(const uint16_t* X, const long long* rows, int n_rows, int G, float* out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long total = (long long)n_rows * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / G), g = (int)(i - (long long)j * G);
    out[i] = (float)X[rows[j] * G + g];
  }
}

// ---- the count statistics of load_adata (`_scdef.py:498-573`) on the device ------------------------------------------
// Per batch b (and over all cells, slot nb): [number of cells, sum of library sizes, sum of squared library sizes] and the
// per-gene totals.  T = float or uint16_t.  One block = 32 rows x all genes: the gene sums of the block's rows are
// accumulated per batch in registers over the rows (rows of one block mostly share few batches) and added with double atomics.
template <typename T>
__global__ void __launch_bounds__(256) k_count_stats(const T* X, const int* batch_id, long long n_rows, int G, int nb,
                                                     double* st, double* gs, float* lib_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ float lib_sm[32];
  __shared__ int b_sm[32];
  const long long r0 = (long long)blockIdx.x * 32;
  const int rows = (int)min((long long)32, n_rows - r0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // library sizes: 8 warps x 4 rows
  for (int rr = warp; rr < rows; rr += 8) {
    const T* row = X + (r0 + rr) * G;
    float acc = 0.f;
    for (int g = lane; g < G; g += 32) acc += (float)row[g];
    acc = warp_sum(acc);
    if (lane == 0) {
      lib_sm[rr] = acc;
      b_sm[rr] = batch_id ? batch_id[r0 + rr] : 0;
      if (lib_out) lib_out[r0 + rr] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int rr = 0; rr < rows; ++rr) {
      const double l = (double)lib_sm[rr];
      const int b = b_sm[rr];
      atomicAdd(&st[nb * 3 + 0], 1.0); atomicAdd(&st[nb * 3 + 1], l); atomicAdd(&st[nb * 3 + 2], l * l);
      if (nb > 1) { atomicAdd(&st[b * 3 + 0], 1.0); atomicAdd(&st[b * 3 + 1], l); atomicAdd(&st[b * 3 + 2], l * l); }
    }
  }
  // gene sums: thread = gene (strided), rows in order; a run of rows of the same batch is summed before it is flushed
  for (int g = threadIdx.x; g < G; g += 256) {
    float all = 0.f, run = 0.f;
    int cur = b_sm[0];
    for (int rr = 0; rr < rows; ++rr) {
      const float x = (float)X[(r0 + rr) * G + g];
      all += x;
      if (nb > 1) {
        if (b_sm[rr] != cur) {
          if (run != 0.f) atomicAdd(&gs[(long long)cur * G + g], (double)run);
          run = 0.f;
          cur = b_sm[rr];
        }
        run += x;
      }
    }
    if (nb > 1 && run != 0.f) atomicAdd(&gs[(long long)cur * G + g], (double)run);
    if (all != 0.f) atomicAdd(&gs[(long long)nb * G + g], (double)all);
  }
}

__global__ void __launch_bounds__(256) k_noise_fill(float* out, int S, int per, PhiloxKey key, NoiseTag tag, long long e0) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long total = (long long)S * per;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    int s = (int)(e / per);
    int i = (int)(e - (long long)s * per);
    out[e] = philox_normal(key, tag, (uint32_t)s, (uint32_t)(e0 + i));
  }
}

}  // namespace scdef
