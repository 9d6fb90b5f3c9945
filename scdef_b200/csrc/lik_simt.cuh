// v1 likelihood kernel: fp32 CUDA-core tiles (any K0 up to 256).  The on-device reference
// for the tcgen05 kernel and the fallback for shapes that one does not take.
// Math: SURVEY.md A.3; reference `_scdef.py:2096-2136`.
//   R = z0 W0,  Lambda = c_n * s_{b(n),g} * R,  ll = sum x log Lambda - Lambda (- lgamma(x+1), added elsewhere)
//   E = x - Lambda,  Q = E / R
//   LOCAL : G_z0 += scale * Q W0^T,  G_c += scale * sum_g E / c
//   GLOBAL: G_W0 += scale * z0^T Q,  G_s[b,g] += scale * sum_{n in b} E / s
#pragma once
#include "common.cuh"

namespace scdef {

struct LikArgs {
  const float* smp_z0;  // (S,B,K0)
  const float* smp_W0;  // (S,K0,G)
  const float* smp_c;   // (S,B)
  const float* smp_s;   // (S,nb,G)
  const float* X;       // rows of length G: batch-major (B,G) if idx_x == nullptr else gathered
  const int64_t* idx_x; // row gather into X (resident matrix) or nullptr
  const int64_t* idx;   // (B) rows in the per-cell constants
  const int32_t* batch_id;  // (N)
  float* G_z0;  // (S,B,K0)  LOCAL
  float* G_c;   // (S,B)     LOCAL
  float* G_W0;  // (S,K0,G)  GLOBAL
  float* G_s;   // (S,nb,G)  GLOBAL
  int B, G, K0, nb, pass;
  float scale;
};

constexpr int LT = 64;        // tile edge (cells and genes)
constexpr int LW_LD = LT + 4; // W tile row stride (float4-aligned, conflict-free for 4 rows)
constexpr int LQ_LD = LT + 1;
constexpr int LIK_MAX_NB_SMEM = 8;

// grid = (gene tiles, cell tiles, S), block = 256
__global__ void __launch_bounds__(256) k_lik_simt(const LikArgs a, int S, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ float sm[];
  __shared__ double red[32];
  __shared__ float rowE[LT];
  __shared__ float colE[LIK_MAX_NB_SMEM][LT];
  __shared__ int tile_b[LT];
  const int K0 = a.K0;
  const int z_ld = K0 + 1;
  float* zt = sm;                   // [LT][z_ld]
  float* wt = zt + LT * z_ld;       // [K0][LW_LD]
  float* Qt = wt + K0 * LW_LD;      // [LT][LQ_LD]
  const int s = blockIdx.z;
  const int g0 = blockIdx.x * LT, j0 = blockIdx.y * LT;
  const int tid = threadIdx.x;
  const bool global_pass = (a.pass == SCDEF_PASS_GLOBAL);

  for (int e = tid; e < LT * K0; e += 256) {
    int r = e / K0, k = e - r * K0;
    int j = j0 + r;
    zt[r * z_ld + k] = j < a.B ? a.smp_z0[((long long)s * a.B + j) * K0 + k] : 0.f;
  }
  for (int e = tid; e < K0 * LT; e += 256) {
    int k = e / LT, g = e - k * LT;
    wt[k * LW_LD + g] = (g0 + g) < a.G ? a.smp_W0[((long long)s * K0 + k) * a.G + g0 + g] : 0.f;
  }
  if (tid < LT) {
    rowE[tid] = 0.f;
    int j = j0 + tid;
    tile_b[tid] = j < a.B ? a.batch_id[a.idx[j]] : 0;
  }
  for (int e = tid; e < LIK_MAX_NB_SMEM * LT; e += 256) (&colE[0][0])[e] = 0.f;
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.f;
  for (int k = 0; k < K0; ++k) {
    float4 w4 = *reinterpret_cast<const float4*>(&wt[k * LW_LD + tx * 4]);
    float zr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) zr[i] = zt[(ty * 4 + i) * z_ld + k];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = fmaf(zr[i], w4.x, acc[i][0]);
      acc[i][1] = fmaf(zr[i], w4.y, acc[i][1]);
      acc[i][2] = fmaf(zr[i], w4.z, acc[i][2]);
      acc[i][3] = fmaf(zr[i], w4.w, acc[i][3]);
    }
  }

  double val = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = ty * 4 + i, j = j0 + r;
    float rsum = 0.f;
    if (j < a.B) {
      const float c = a.smp_c[(long long)s * a.B + j];
      const int b = tile_b[r];
      const long long xrow = a.idx_x ? a.idx_x[j] : j;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int gl = tx * 4 + jj, g = g0 + gl;
        float q = 0.f;
        if (g < a.G) {
          const float sg = a.smp_s[((long long)s * a.nb + b) * a.G + g];
          const float x = a.X[xrow * a.G + g];
          const float R = acc[i][jj];
          const float lam = c * sg * R;
          const float E = x - lam;
          q = E / R;
          val += (double)((x > 0.f ? x * logf(lam) : 0.f) - lam);
          rsum += E;
          if (global_pass) {
            float ds = a.scale * E / sg;
            if (a.nb <= LIK_MAX_NB_SMEM) atomicAdd(&colE[b][gl], ds);
            else atomicAdd(&a.G_s[((long long)s * a.nb + b) * a.G + g], ds);
          }
        }
        Qt[r * LQ_LD + gl] = q;
      }
      if (!global_pass) atomicAdd(&rowE[r], rsum / c);
    } else {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Qt[r * LQ_LD + tx * 4 + jj] = 0.f;
    }
  }
  __syncthreads();

  if (!global_pass) {
    // G_z0[j][k] += scale * sum_g Q[r][g] W[k][g]
    const int r = tid >> 2, j = j0 + r;
    if (j < a.B) {
      for (int k = tid & 3; k < K0; k += 4) {
        float d = 0.f;
#pragma unroll 8
        for (int g = 0; g < LT; ++g) d = fmaf(Qt[r * LQ_LD + g], wt[k * LW_LD + g], d);
        atomicAdd(&a.G_z0[((long long)s * a.B + j) * K0 + k], a.scale * d);
      }
    }
    if (tid < LT && j0 + tid < a.B)
      atomicAdd(&a.G_c[(long long)s * a.B + j0 + tid], a.scale * rowE[tid]);
  } else {
    // G_W0[k][g] += scale * sum_r z[r][k] Q[r][g]
    const int gl = tid & 63, g = g0 + gl;
    if (g < a.G) {
      for (int k = tid >> 6; k < K0; k += 4) {
        float d = 0.f;
#pragma unroll 8
        for (int r = 0; r < LT; ++r) d = fmaf(zt[r * z_ld + k], Qt[r * LQ_LD + gl], d);
        atomicAdd(&a.G_W0[((long long)s * K0 + k) * a.G + g], a.scale * d);
      }
      if (a.nb <= LIK_MAX_NB_SMEM && tid < LT) {
        for (int b = 0; b < a.nb; ++b) {
          float v = colE[b][gl];
          if (v != 0.f) atomicAdd(&a.G_s[((long long)s * a.nb + b) * a.G + g], v);
        }
      }
    }
  }
  double tot = block_sum_d(val, red);
  if (tid == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)S);
}

// - scale * sum_j sum_g lgamma(x_jg + 1): parameter-free part of the Poisson log-pmf
// (`_scdef.py:2119`).  With the per-cell sums precomputed (resident counts): one warp per minibatch row, grid = ceil(B/8).
// Without them (host placement: the rows arrive per step): one warp per (row, 512-gene chunk), grid = ceil(B*chunks/8) —
// a warp per whole row left this kernel at 32 CTAs x 157 lgammaf per lane, ~0.09 ms of an otherwise idle GPU per step.
constexpr int XLG_CHUNK = 512;
__global__ void __launch_bounds__(256) k_x_lgamma(const float* X, const int64_t* idx_x,
                                                  const float* rowsum, const int64_t* idx_cells,
                                                  int B, int G, float scale, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double val = 0.0;
  if (rowsum) {
    if (w < B && lane == 0) val = (double)rowsum[idx_cells[w]];
  } else {
    const int chunks = (G + XLG_CHUNK - 1) / XLG_CHUNK;
    const long long j = w / chunks;
    const int c = (int)(w - j * chunks);
    if (j < B) {
      const long long row = idx_x ? idx_x[j] : j;
      const int g1 = min(G, (c + 1) * XLG_CHUNK);
      float acc = 0.f;
      for (int g = c * XLG_CHUNK + lane; g < g1; g += 32) {
        const float x = X[row * G + g];
        if (x > 1.5f) acc += lgammaf(x + 1.f);
      }
      val = (double)acc;
    }
  }
  double tot = block_sum_d(val, red);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&loss_out[0], tot * (double)scale);
}

}  // namespace scdef
