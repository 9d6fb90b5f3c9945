// The fused Poisson-likelihood kernel of BOTH passes: one persistent CTA-pair kernel, two orientations
// (included by lik_tc.cu inside its anonymous namespace, after lik_tc_pair.cuh whose PTX wrappers it uses).
//
// Algebra that takes the rank-1 terms out of the tile loop (reference `_scdef.py:2096-2136`):
//   Lambda = c s R,   Q = (x - Lambda) / R = x / R - c s^T            with  R = z0 W
//   dz0 = Q W^T   = (X/R) W^T - c (W s_b)^T            (X/R) is zero wherever x = 0, i.e. for ~90 % of the entries
//   dW  = z0^T Q  = z0^T (X/R) - sum_b ctilde_b s_b^T        ctilde_b = sum_{n in b} c_n z_n
//   dc_n = (rowsum_g x_ng - c_n z_n . (W s_b)) / c_n
//   ll  = sum_{x>0} x log R + sum_n rowsumX_n log c_n + sum_{b,g} colsumX_b[g] log s_bg - sum_n c_n z_n . (W s_b) - lgamma terms
// so the tile loop needs ONLY x and R:  q = x > 0 ? x / R : 0  (one MUFU.RCP, one multiply, one select per element) and,
// when the loss is wanted, x log R.  The cell scale, the gene-scale tile, Lambda and E are gone from the epilogue; the rank-1
// corrections are a K0-dot per (sample, cell) (`k_tc_rank1`) and nb FMAs per element of the W_0 fold.
//
// One kernel template, two orientations of the same tile pipeline ("rows" = the M dimension of the pair MMAs):
//   LOCAL  (GENES = false): rows = cells.  ROW unit A1 = z0^s of 128 cells (K-major), static per segment (rb = cell block, s);
//          COL stream = W_0^s tile images, pair-tile = 128 genes.  D2 = dz0[rows, k0] accumulated over the segment,
//          drained to the partial buffer (`k_tc_rank1` sums the slots and adds the rank-1 term).
//   GLOBAL (GENES = true):  rows = genes.  ROW unit A1 = W_0^s of 128 genes (MN-major, assembled from two 64-gene tile
//          images with 1 KB bulk copies), static per segment (rb = gene block, s); COL stream = z0^s, pair-tile = 128 cells
//          (image `zimg_g`: one 56 KB bulk copy per stage = both 64-cell halves, the CTA's own first).
//          The second GEMM runs PER CTA (cta_group::1): D2[k0 (M = 128), the CTA's 128 genes (N)] += z_h^T Q_h^T, i.e.
//          A = z of half h (MN-major, M = k0), B = the CTA's Q tile (K-major, N = genes), K = 64 cells — the contraction is
//          over the cells, which the rate GEMM has as columns, and the result comes out with k0 on the tensor-memory lanes.
//          So the FOLD  t = W (scale dW' - sum_b ctS_b[k] s_b[g] - gw B_k) + gw (A_k - 1),  d mu += t,  sum t log W += t log W
//          runs with thread = k0 row: A_k, B_k, ctS_b[k] are per-thread constants and W^s[k][32 genes] is four 16 B loads per
//          plane.  Both accumulators (128 lanes x 128 genes each) live in tensor memory next to R and D2 (4 x 128 columns),
//          and are flushed to the partial buffer when the gene block changes.
// Rate GEMM   D1[256 rows, 128 cols] = A1 B1      (cta_group::2, M = 256, N = 128 split over the pair, K = KP, 3 split-bf16 products)
// second GEMM LOCAL : D2[256 rows, KP] += Q_h B2_h   (cta_group::2, M = 256, N = KP split over the pair, K = 64 cols of half h)
//             GLOBAL: see above (cta_group::1, each CTA's own issuing thread; the peer's doubles as the TMA relay)
// Flattened work f = (rb * S + s) * n_pt + pt, cut into P contiguous chunks, one per CTA pair (148 SMs -> 74 pairs).
// Barrier protocol: that of k_lik_pair_flat (lik_tc_pair.cuh; modelled in tests/test_pair_protocol.py); in the GLOBAL
// orientation the fold reads W^s from the tile images in global memory, so the ROW unit is released by the tensor core as in LOCAL.
#pragma once

struct FlatGeom {
  int KP;
  int n_rblk;   // row blocks (256 rows each): LOCAL cell blocks, GLOBAL gene blocks
  int n_pt;     // column pair-tiles (128 columns) per segment: LOCAL gene pair-tiles, GLOBAL cell pair-tiles
  int P;        // CTA pairs
  int slots;    // partial-buffer slots: LOCAL per (sample), GLOBAL per gene block
  int n_units;  // 128-row units in the count image of this orientation (= 2 * n_rblk)
};

constexpr uint32_t FOLD_GRP = 144;                 // 8 k0 rows x 16 B of one 8-gene group, padded so that the 4 groups of a warp fall into different banks
constexpr uint32_t FOLD_SCRATCH = 2 * 4 * FOLD_GRP;   // hi and lo plane

struct FlatSmem {
  uint32_t a1, st[2], q, fold, misc, total;
  uint32_t a1_plane, b1_plane, b2_piece, b2_off, stage, q_plane;
  uint32_t z_rs, zb_rs, u_rs;
};
__host__ __device__ inline FlatSmem plan_flat_smem(int KP) {
  FlatSmem s;
  s.z_rs = (uint32_t)(KP / 8) * 128u;     // [row][k0] layouts: bytes per 8-row group
  s.zb_rs = (uint32_t)(KP / 16) * 128u;   // the same for a k0 half
  s.u_rs = 16u * 128u;                    // assembled W unit [k0][128 genes]: bytes per 8-k0 group
  s.a1_plane = (uint32_t)KP * 256u;       // 128 rows x KP x bf16
  s.b1_plane = (uint32_t)KP * 128u;       // 64 columns x KP
  s.b2_piece = (uint32_t)KP * 64u;        // 64 columns x KP/2
  s.b2_off = 2u * s.b1_plane;
  s.stage = s.b2_off + 4u * s.b2_piece;
  s.q_plane = 16u * Q_RS;
  uint32_t o = 0;
  s.a1 = o; o += 2u * s.a1_plane;
  s.st[0] = o; o += s.stage;
  s.st[1] = o; o += s.stage;
  s.q = o; o += 2u * s.q_plane;
  s.fold = o; o += (uint32_t)(MAX_NB * 128 * 4);   // GLOBAL, more than one batch: the unit's gene-scale samples of the segment
  s.misc = o;
  s.total = o + 1024u;
  return s;
}


__device__ __forceinline__ float rcp_ftz(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_ftz(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// Per-element arithmetic of the epilogue, per orientation (measured on a B200 at C5, ms per launch):
//   0 = clamp R at 1e-35 and multiply (no select), 1 = select on x > 0 with the raw rcp.approx.ftz, 2 = select with __fdividef.
//   LOCAL:  0 -> 0.282, 1 -> 0.237, 2 -> 0.243        GLOBAL (also x log2 R): 0 -> 0.375, 1 -> 0.376, 2 -> 0.386
#ifndef SCDEF_MATH_LOCAL
#define SCDEF_MATH_LOCAL 1
#endif
#ifndef SCDEF_MATH_GLOBAL
#define SCDEF_MATH_GLOBAL 0
#endif
#ifndef FOLD_UNROLL
#define FOLD_UNROLL 4   // measured: 0.375 -> 0.346 ms per GLOBAL launch at C5 against the rolled loop (no register moves for the prefetch rotation)
#endif
constexpr int FOLD_UNROLL_N = FOLD_UNROLL;
constexpr float LN2_F = 0.69314718055994530942f;
constexpr float LG2_V_HI_M = 49.828777f;   // log2(0.9999e15) = -log2(1.0001e-15) to 1e-8: the margins of V_LO_M / V_HI_M

// 32 lanes x 4 consecutive 32-bit columns (the fold works in groups of four elements to keep its live state small)
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8_wait3(uint32_t (&a)[8], uint32_t (&b)[8], uint32_t (&c)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                 "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                 "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
               :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
               "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
}

struct FlatPos {   // position in the flattened space
  int rb, s, pt, g;
  __device__ __forceinline__ void next(int n_pt, int S) {
    if (++pt == n_pt) { pt = 0; ++g; if (++s == S) { s = 0; ++rb; } }
  }
};

// tensor-memory columns.  LOCAL: R[2] at 0 / 128, dz0[2] at 256 / 384.  GLOBAL: ONE R buffer at 0 (the whole-pair-tile
// epilogue has both halves of R in registers ~1k cycles after the rate GEMM commits, and the next rate GEMM's result is
// only needed a pair-tile later), D2 = dW[k0, 128 genes] at 128, and the two fold accumulators at 256 / 384 (d mu,
// sum t log W; 128 genes each): the accumulators of a gene block never occupy registers.
constexpr uint32_t FLAT_TMEM_D2_LOCAL = 256, FLAT_TMEM_D2_GENES = 128, FLAT_TMEM_ACC_MU = 256, FLAT_TMEM_ACC_LW = 384;

// z0 samples -> GLOBAL-orientation column images: per (s, 128-cell pair-tile, rank) one contiguous stage
//   [Z_own hi | Z_own lo | Z_other hi | Z_other lo],  Z_x = z[the 64 cells of half x][all k0]  as [cell][k0], 8-row-group stride z_rs
// Z_own (half = rank) is the CTA's N-half of the rate GEMM's B (K-major); Z_0 / Z_1 are the A operands (MN-major, M = k0) of
// the CTA's second GEMMs of the two halves.
__global__ void __launch_bounds__(256) k_tc_zprep_g(const LikTcArgs a, uint8_t* zimg, int KP, int n_ptc) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const FlatSmem L = plan_flat_smem(KP);
  const int nchunk = KP / 8;
  const long long per_pt = 128ll * nchunk;   // (cell of the pair-tile, k0 chunk)
  const long long total = (long long)a.S * n_ptc * per_pt;
  const bool vec = (a.K0 & 3) == 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long sp = t / per_pt;
    const int i = (int)(t - sp * per_pt);
    const int s = (int)(sp / n_ptc), pt = (int)(sp - (long long)s * n_ptc);
    const int cl = i % 128, c = i / 128;          // cell of the pair-tile, k0 chunk
    const int j = pt * 128 + cl, k = c * 8;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (j < a.B) {
      const float* row = a.smp_z0 + ((long long)s * a.B + j) * a.K0;
      if (vec) {
        if (k + 4 <= a.K0) { float4 v = *reinterpret_cast<const float4*>(row + k); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
        if (k + 8 <= a.K0) { float4 v = *reinterpret_cast<const float4*>(row + k + 4); x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (k + e < a.K0) x[e] = row[k + e];
      }
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    uint8_t* base = zimg + sp * (2ll * L.stage);
    const int h = cl >> 6, r = cl & 63;           // 64-cell half, row inside it
    const uint32_t off = (uint32_t)(r >> 3) * L.z_rs + (uint32_t)c * 128u + (uint32_t)(r & 7) * 16u;
    for (int rk = 0; rk < 2; ++rk) {              // own half first in each rank's stage
      uint8_t* d = base + (long long)rk * L.stage + (rk == h ? 0u : 2u * L.b1_plane) + off;
      *reinterpret_cast<uint4*>(d) = hi;
      *reinterpret_cast<uint4*>(d + L.b1_plane) = lo;
    }
  }
}

// minibatch counts -> GLOBAL-orientation image (row = gene, column = cell), same addressing as k_tc_xprep:
// element (gene = unit*128 + q*32 + lane, cell = tile*64 + v*4 + i) at ((((tile*units + unit)*4 + q)*16 + v)*32 + lane)*4 + i
__global__ void __launch_bounds__(256) k_tc_xprep_t(const LikTcArgs a, float* ximg, int n_ctiles, int n_units) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long total = (long long)n_ctiles * n_units * 4 * 16 * 32;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(t % 32);               // consecutive threads = consecutive genes of one cell: coalesced reads
    long long rest = t / 32;
    const int q = (int)(rest % 4); rest /= 4;
    const int unit = (int)(rest % n_units); rest /= n_units;
    const int v = (int)(rest % 16);
    const int tile = (int)(rest / 16);
    const int g = unit * 128 + q * 32 + lane, j0 = tile * 64 + v * 4;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (g < a.G) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (j0 + i < a.B) x[i] = a.X[(a.idx_x ? a.idx_x[j0 + i] : (long long)(j0 + i)) * a.G + g];
    }
    *reinterpret_cast<float4*>(ximg + ((((((long long)tile * n_units + unit) * 4 + q) * 16 + v) * 32 + lane) * 4)) = make_float4(x[0], x[1], x[2], x[3]);
  }
}

// sum_{b,g} colsumX_b[g] log s^s_{b,g}  (the gene-scale part of sum x log Lambda), all samples
__global__ void __launch_bounds__(256) k_tc_loss_gs(const LikTcArgs a, const float* colsumX, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const long long n = (long long)a.S * a.nb * a.G, per = (long long)a.nb * a.G;
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float cx = colsumX[i % per];
    if (cx > 0.f) acc += (double)(cx * __logf(a.smp_s[i]));
  }
  const double tot = block_sum_d(acc, red);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
}

// The rank-1 parts, one warp per (sample, cell):  dot = z_n . (W s_b)
//   LOCAL : G_z0[s][j][k] += scale (sum_slots part - c Ws_b[k]),   G_c[s][j] += scale (rowsumX_j / c - dot)
//   loss  : - scale / S * (rowsumX_j log c - c dot)
__global__ void __launch_bounds__(256) k_tc_rank1(const LikTcArgs a, const TcScratch ws, const float* part, int slots_ld,
                                                  long long T, int P, int n_pt, int do_grad, int do_loss, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), total = (long long)a.S * a.B;
  double lacc = 0.0;
  if (w < total) {
    const int s = (int)(w / a.B), j = (int)(w - (long long)s * a.B);
    const float c = a.smp_c[w];
    const int b = ws.brow[j];
    const float* Wsb = ws.Ws + ((long long)s * a.nb + b) * a.K0;
    const float* z = a.smp_z0 + w * a.K0;
    int nsl = 0;
    long long pbase = 0;
    if (do_grad) {
      // the chunks that hold a piece of segment (rb, s): rb = j / 256
      const long long f0 = ((long long)(j / CB) * a.S + s) * n_pt;
      nsl = flat_chunk_of(f0 + n_pt - 1, T, P) - flat_chunk_of(f0, T, P) + 1;
      pbase = ((long long)s * slots_ld * a.B + j) * a.K0;
    }
    float dot = 0.f;
    if ((a.K0 & 3) == 0) {
      // one float4 per lane and array: a single round of loads for K0 <= 128
      for (int k = lane * 4; k < a.K0; k += 128) {
        const float4 z4 = *reinterpret_cast<const float4*>(z + k), w4 = *reinterpret_cast<const float4*>(Wsb + k);
        dot = fmaf(z4.x, w4.x, fmaf(z4.y, w4.y, fmaf(z4.z, w4.z, fmaf(z4.w, w4.w, dot))));
        if (do_grad) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int q = 0; q < nsl; ++q) {
            const float4 p4 = *reinterpret_cast<const float4*>(part + pbase + (long long)q * a.B * a.K0 + k);
            acc.x += p4.x; acc.y += p4.y; acc.z += p4.z; acc.w += p4.w;
          }
          float4 g4 = *reinterpret_cast<float4*>(a.G_z0 + w * a.K0 + k);
          g4.x += a.scale * (acc.x - c * w4.x); g4.y += a.scale * (acc.y - c * w4.y);
          g4.z += a.scale * (acc.z - c * w4.z); g4.w += a.scale * (acc.w - c * w4.w);
          *reinterpret_cast<float4*>(a.G_z0 + w * a.K0 + k) = g4;
        }
      }
    } else {
      for (int k = lane; k < a.K0; k += 32) {
        const float wsk = Wsb[k];
        dot = fmaf(z[k], wsk, dot);
        if (do_grad) {
          float acc = 0.f;
          for (int q = 0; q < nsl; ++q) acc += part[pbase + (long long)q * a.B * a.K0 + k];
          a.G_z0[w * a.K0 + k] += a.scale * (acc - c * wsk);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) {
      const float rs = ws.rowsumx[j];
      if (do_grad) a.G_c[w] += a.scale * (rs / c - dot);
      if (do_loss) lacc = (double)(rs * __logf(c) - c * dot);
    }
  }
  if (do_loss) {
    const double tot = block_sum_d(lacc, red);
    if (threadIdx.x == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
  }
}

// W_0: entropy value and the final (mu, rho) gradients from the gene-block slots of the GLOBAL orientation
__global__ void __launch_bounds__(256) k_tc_w0_final_flat(const LikTcArgs a, const float* part, long long T, int P, int per_rb,
                                                          double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const long long n = (long long)a.K0 * a.G;
  const float wH = a.anneal * a.global_weight, invS = 1.f / (float)a.S;
  double ent = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float mu = a.mu_W0[i], rho = a.rho_W0[i];
    ent += (double)(fminf(fmaxf(mu, LOG_1E_10), LOG_1E6) + fminf(fmaxf(rho, LOG_1E_10), LOG_1E10) + HALF_LOG_2PI_PLUS_HALF);
    if (a.gmu_W0) {
      float dmu = 0.f, drho = 0.f;
      if (!a.w0_stopped) {
        const int rb = (int)((i % a.G) / 256);
        const long long f0 = (long long)rb * per_rb;
        const int nsl = part ? flat_chunk_of(f0 + per_rb - 1, T, P) - flat_chunk_of(f0, T, P) + 1 : 0;
        float gv = 0.f, glw = 0.f;
        for (int c = 0; c < nsl; ++c) {
          gv += part[((long long)c * 2 + 0) * n + i];
          glw += part[((long long)c * 2 + 1) * n + i];
        }
        const float gve = glw - fminf(fmaxf(mu, LOG_1E_10), LOG_1E6) * gv;   // sum_s t_s (sigma eps)_s, sigma eps = log W - mu~
        const bool in_mu = mu >= LOG_1E_10 && mu <= LOG_1E6, in_rho = rho >= LOG_1E_10 && rho <= LOG_1E10;
        dmu = in_mu ? -(gv * invS + wH) : 0.f;
        drho = in_rho ? -(gve * invS + wH) : 0.f;
      }
      a.gmu_W0[i] = dmu;
      a.grho_W0[i] = drho;
    }
  }
  const double tot = block_sum_d(ent, red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[1], -tot * (double)wH);
}

// ---- the kernel ------------------------------------------------------------------------------------------------
template <bool GENES, bool WANT_LL, bool MAT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PIPE_NT, 1)   // 96 registers: 112 x 576 threads is refused at launch
    k_lik_flat(const LikTcArgs a, const FlatGeom gm, const TcScratch ws, double* loss_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const FlatSmem L = plan_flat_smem(KP);
  FlatMisc* mi = reinterpret_cast<FlatMisc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  PT_DECL;

  // ---- work of this pair ----
  const int n_pt = gm.n_pt, P = gm.P, S = a.S;
  const long long T = (long long)gm.n_rblk * S * n_pt;
  const int chunk = blockIdx.x >> 1;
  const long long f0 = (long long)chunk * T / P, f1 = (long long)(chunk + 1) * T / P;
  const int n_outer = (int)(f1 - f0);
  const int U = 2 * n_outer;
  FlatPos cur0;
  {
    const long long sg = f0 / n_pt;
    cur0.rb = (int)(sg / S); cur0.s = (int)(sg % S); cur0.pt = (int)(f0 % n_pt); cur0.g = 0;
  }
  const int n_seg = n_outer > 0 ? (int)((f1 - 1) / n_pt - f0 / n_pt) + 1 : 0;
  constexpr int ND2 = GENES ? 1 : 2;   // D2 accumulators in tensor memory

  // ---- prologue ----
  if (threadIdx.x == 0) {
    mbar_init(&mi->z_full, 1);
    mbar_init(&mi->pz_full, 1);
    mbar_init(&mi->z_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mi->w_full[i], 1);
      mbar_init(&mi->pw_full[i], 1);
      // GLOBAL: a stage is read by the pair's rate GEMM (leader's commit, multicast) and by this CTA's second GEMMs
      mbar_init(&mi->w_empty[i], GENES ? 2 : 1);
      mbar_init(&mi->r_full[i], 1);
      mbar_init(&mi->r_empty[i], 2 * E_WARPS);
      mbar_init(&mi->dz_full[i], 1);
      mbar_init(&mi->dz_empty[i], GENES ? E_WARPS : 2 * E_WARPS);   // GLOBAL: D2 and the Q hand-off are per CTA
    }
    mbar_init(&mi->q_full, GENES ? E_WARPS : 2 * E_WARPS);
    mbar_init(&mi->q_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&mi->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = mi->tmem_base;
  const PipeSmem IL = plan_pipe_smem(KP);   // layout of one W_0 tile image in global memory (k_tc_wgen)
  // Everything above touched shared and tensor memory only: with programmatic dependent launch it ran while the previous
  // kernel's last blocks were retiring.  The operands in global memory are complete from here on.
  pdl_grid_wait();
  // All CTAs of this persistent grid are resident from the start, so the next kernel's blocks cannot take a slot this
  // grid still needs: let them be scheduled now (they block in their own griddepcontrol.wait until this grid is done).
  pdl_launch_dependents();
  auto first_in_seg = [&](int oi, const FlatPos& c) { return c.pt == 0 || oi == 0; };
  auto last_in_seg = [&](int oi, const FlatPos& c) { return c.pt == n_pt - 1 || oi == n_outer - 1; };

  if (warp == COPY_WARP) {
    // =========================================== COPY (TMA), each CTA for itself ===========================
    if (lane == 0) {
      FlatPos c = cur0;
      for (int oi = 0; oi < n_outer; ++oi, c.next(n_pt, S)) {
        const int st = oi & 1;
        PT_WAIT(0, pair_wait(&mi->w_empty[st], ((oi >> 1) & 1) ^ 1, 1));
        mbar_arrive_expect_tx(&mi->w_full[st], L.stage);
        const uint32_t dst = sbase + L.st[st];
        if (GENES) {
          bulk_g2s(dst, ws.zimg_g + (((long long)c.s * n_pt + c.pt) * 2 + rank) * L.stage, L.stage, &mi->w_full[st]);
        } else {
          const uint8_t* img0 = ws.wimg + ((long long)c.s * (2 * n_pt) + 2 * c.pt) * IL.w_unit;
          const uint8_t* img1 = img0 + IL.w_unit;
          bulk_g2s(dst, rank ? img1 : img0, 2u * L.b1_plane, &mi->w_full[st]);                      // B1: hi | lo of the own image
          const uint32_t row_off = rank * L.b2_piece;                                                // k0 rows [KP/2 rank, +KP/2)
          bulk_g2s(dst + L.b2_off + 0u * L.b2_piece, img0 + row_off, L.b2_piece, &mi->w_full[st]);
          bulk_g2s(dst + L.b2_off + 1u * L.b2_piece, img0 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]);
          bulk_g2s(dst + L.b2_off + 2u * L.b2_piece, img1 + row_off, L.b2_piece, &mi->w_full[st]);
          bulk_g2s(dst + L.b2_off + 3u * L.b2_piece, img1 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]);
        }
        if (first_in_seg(oi, c)) {
          // the ROW unit of the new segment replaces the old one once the segment's last rate GEMM has read it
          PT_WAIT(1, pair_wait(&mi->z_empty, ((uint32_t)c.g & 1u) ^ 1u, 17));
          const uint32_t unit = (uint32_t)(2 * c.rb) + rank;
          if (GENES) {
            mbar_arrive_expect_tx(&mi->z_full, 2u * L.a1_plane);
            const int nkg = KP / 8, n_img = a.n_wimg;
            for (int hsel = 0; hsel < 2; ++hsel) {
              const int img = (int)unit * 2 + hsel;   // 64-gene tile image of this half of the unit
              const uint8_t* src = ws.wimg + ((long long)c.s * n_img + img) * IL.w_unit;
              for (int pl = 0; pl < 2; ++pl)
                for (int kg = 0; kg < nkg; ++kg)
                  bulk_g2s(sbase + L.a1 + (uint32_t)pl * L.a1_plane + (uint32_t)kg * L.u_rs + (uint32_t)hsel * 1024u,
                           src + (uint32_t)pl * IL.w_plane + (uint32_t)kg * W_RS, 1024u, &mi->z_full);
            }
          } else {
            mbar_arrive_expect_tx(&mi->z_full, 2u * L.a1_plane);
            bulk_g2s(sbase + L.a1, ws.zimg + ((long long)c.s * gm.n_units + unit) * (2u * L.a1_plane), 2u * L.a1_plane, &mi->z_full);
          }
        }
      }
      // see the last multicast commits land before this CTA can exit
      for (int st = 0; st < 2; ++st) {
        const int uses = (n_outer + 1 - st) / 2;
        if (uses > 0) pair_wait(&mi->w_empty[st], (uint32_t)(uses - 1) & 1u, 15);
      }
      if (U > 0) pair_wait(&mi->q_empty, (uint32_t)(U - 1) & 1u, 16);
      if (n_seg > 0) pair_wait(&mi->z_empty, (uint32_t)(n_seg - 1) & 1u, 18);
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    // GLOBAL: the second GEMM of a sub-unit, by THIS CTA's issuing thread (cta_group::1):
    //   D2[k0 (M = 128), the CTA's 128 genes] (+)= z_h^T Q^T     A = z of half h (MN-major, M = k0), B = Q (K-major, N = genes)
    const uint32_t idesc2g = umma_idesc(128, 128, 1, 0);
    const uint64_t dqb = umma_desc(sbase + L.q, 128u, Q_RS);
    auto issue_mma2g = [&](int u, const FlatPos& c2, int n_outer_) {
      const int oi = u >> 1, h = u & 1, st = oi & 1;
      const bool seg_first = (c2.pt == 0 || oi == 0) && h == 0, seg_last = (c2.pt == n_pt - 1 || oi == n_outer_ - 1) && h == 1;
      if (seg_first) PT_WAIT(7, pair_wait(&mi->dz_empty[0], ((uint32_t)c2.g & 1u) ^ 1u, 19));   // the fold has read D2
      tc_fence_after();
      PT_MARK();
      const uint32_t d = tmem + FLAT_TMEM_D2_GENES;
      uint64_t da = umma_desc(sbase + L.st[st] + ((uint32_t)h == rank ? 0u : 2u * L.b1_plane), L.z_rs, 128u), db = dqb;
      const uint64_t apl = L.b1_plane >> 4, bpl = L.q_plane >> 4, a_k = (uint64_t)(2u * (L.z_rs >> 4));
      umma_bf16(d, da + apl, db, idesc2g, seg_first ? 0u : 1u);
      umma_bf16(d, da, db + bpl, idesc2g, 1u);
      umma_bf16(d, da, db, idesc2g, 1u);
#pragma unroll
      for (int ks = 1; ks < TG / 16; ++ks) {
        da += a_k;
        db += 16;
        umma_bf16(d, da + apl, db, idesc2g, 1u);
        umma_bf16(d, da, db + bpl, idesc2g, 1u);
        umma_bf16(d, da, db, idesc2g, 1u);
      }
      umma_commit(&mi->q_empty);
      if (h == 1) umma_commit(&mi->w_empty[st]);
      if (seg_last) umma_commit(&mi->dz_full[0]);
      PT_LAP(5);
    };
    if (lane == 0 && !leader) {
      // =========================================== RELAY (peer) ==========================================
      // forwards "my TMA landed" to the leader; GLOBAL: also issues this CTA's second GEMMs
      const uint32_t r_pz = mapa_rank(&mi->pz_full, 0), r_pw0 = mapa_rank(&mi->pw_full[0], 0), r_pw1 = mapa_rank(&mi->pw_full[1], 0);
      if (!GENES) {
        FlatPos c = cur0;
        for (int oi = 0; oi < n_outer; ++oi, c.next(n_pt, S)) {
          PT_WAIT(0, pair_wait(&mi->w_full[oi & 1], (oi >> 1) & 1, 3));
          mbar_arrive_remote((oi & 1) ? r_pw1 : r_pw0);
          if (first_in_seg(oi, c)) {
            PT_WAIT(1, pair_wait(&mi->z_full, (uint32_t)c.g & 1u, 2));
            mbar_arrive_remote(r_pz);
          }
        }
      } else {
        FlatPos cz = cur0, c2 = cur0;   // cursors of the next ROW unit to announce and of the next second GEMM
        int rw = 0, rz = 0, u = 0;      // stages / ROW units announced, sub-units issued
        const long long t_begin = clock64();
        while (rw < n_outer || rz < n_seg || u < U) {
          bool moved = false;
          if (rw < n_outer && mbar_test(&mi->w_full[rw & 1], (uint32_t)(rw >> 1) & 1u)) {
            mbar_arrive_remote((rw & 1) ? r_pw1 : r_pw0);
            ++rw; moved = true;
          }
          if (rz < n_seg && mbar_test(&mi->z_full, (uint32_t)rz & 1u)) {
            mbar_arrive_remote(r_pz);
            ++rz; moved = true;
          }
          if (u < U && mbar_test(&mi->q_full, (uint32_t)u & 1u)) {
            issue_mma2g(u, c2, n_outer);
            if (u & 1) c2.next(n_pt, S);
            ++u; moved = true;
          }
          if (!moved && clock64() - t_begin > 4 * PAIR_DEADLINE) pair_stuck(22, (int)blockIdx.x, u, rw);
        }
        (void)cz;
      }
    } else if (lane == 0) {
      // =========================================== MMA ISSUER (leader) ===================================
      // descriptor: lbo = byte stride between core matrices along K, sbo = along M / N (see lik_tc.cu)
      const uint32_t idesc1 = GENES ? umma_idesc(256, 2 * TG, 1, 0) : umma_idesc(256, 2 * TG, 0, 1);
      const uint32_t idesc2 = umma_idesc(256, KP, 0, 0);
      const uint64_t d_a1 = GENES ? umma_desc(sbase + L.a1, L.u_rs, 128u) : umma_desc(sbase + L.a1, 128u, L.z_rs);
      const uint64_t a1_k = GENES ? (uint64_t)(2u * (L.u_rs >> 4)) : 16ull;
      uint64_t d_b1[2], d_b2[2];
      for (int i = 0; i < 2; ++i) {
        d_b1[i] = GENES ? umma_desc(sbase + L.st[i], 128u, L.z_rs) : umma_desc(sbase + L.st[i], W_RS, 128u);
        d_b2[i] = umma_desc(sbase + L.st[i] + L.b2_off, 128u, W_RS);
      }
      const uint64_t b1_k = GENES ? 16ull : (uint64_t)(2u * (W_RS >> 4));
      const uint64_t dq_k = umma_desc(sbase + L.q, 128u, Q_RS);
      const uint64_t a1pl = L.a1_plane >> 4, b1pl = L.b1_plane >> 4, qpl = L.q_plane >> 4, b2pl = L.b2_piece >> 4;
      const int nk1 = KP / 16;
      FlatPos c1 = cur0, c2 = cur0;
      auto ready1 = [&](int oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        if (first_in_seg(oi, c1) && !(mbar_test(&mi->z_full, (uint32_t)c1.g & 1u) && mbar_test_remote(&mi->pz_full, (uint32_t)c1.g & 1u)))
          return false;
        return mbar_test(&mi->w_full[st], par) && mbar_test_remote(&mi->pw_full[st], par) &&
               mbar_test_remote(&mi->r_empty[GENES ? 0 : st], (GENES ? ((uint32_t)oi & 1u) : par) ^ 1u);
      };
      // The rate GEMM of a pair-tile can be issued in chunks of k-steps (state ks1 / da1 / db1), see MMA1_CHUNK below.
      int ks1 = 0;                 // k-steps of the pending rate GEMM (pair-tile next1) issued so far
      uint64_t da1 = 0, db1 = 0;
      auto issue_mma1 = [&](int oi, int max_steps) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        // R: double-buffered with the stage (LOCAL) or a single buffer (GLOBAL: barrier 0, one phase per pair-tile)
        const int rs = GENES ? 0 : st;
        const uint32_t d = tmem + (uint32_t)rs * (2 * TG);
        if (ks1 == 0) {
          if (first_in_seg(oi, c1)) {
            PT_WAIT(6, pair_wait(&mi->z_full, (uint32_t)c1.g & 1u, 7));
            PT_WAIT(6, pair_wait_remote(&mi->pz_full, (uint32_t)c1.g & 1u, 8));
          }
          PT_WAIT(0, pair_wait(&mi->w_full[st], par, 4));
          PT_WAIT(1, pair_wait_remote(&mi->pw_full[st], par, 5));
          const uint32_t rpar = GENES ? ((uint32_t)oi & 1u) : par;
          PT_WAIT(2, pair_wait_remote(&mi->r_empty[rs], rpar ^ 1u, 6));
          tc_fence_after();
          da1 = d_a1;
          db1 = d_b1[st];
        }
        PT_MARK();
        for (int n = 0; n < max_steps && ks1 < nk1; ++n, ++ks1) {
          umma2_bf16(d, da1 + a1pl, db1, idesc1, ks1 > 0 ? 1u : 0u);
          umma2_bf16(d, da1, db1 + b1pl, idesc1, 1u);
          umma2_bf16(d, da1, db1, idesc1, 1u);
          da1 += a1_k;
          db1 += b1_k;
        }
        PT_LAP(4);
        if (ks1 < nk1) return false;
        umma2_commit(&mi->r_full[rs]);
        if (GENES) umma2_commit(&mi->w_empty[st]);               // the pair's rate GEMM has read both CTAs' stage
        if (last_in_seg(oi, c1)) umma2_commit(&mi->z_empty);   // the segment's ROW unit has been read by its last rate GEMM
        c1.next(n_pt, S);
        ks1 = 0;
        return true;
      };
      // k-steps per slipped-in chunk.  Measured on a B200 (round 2): chunks of 2 k-steps (so that a Q tile arriving
      // meanwhile waits ~0.6k instead of ~2k cycles of tensor pipe) made both passes 3 % SLOWER (0.259 -> 0.268 ms LOCAL,
      // 0.441 -> 0.453 GLOBAL): the pipe is not the resource the epilogue waits for.  So: the whole rate GEMM at once.
      constexpr int MMA1_CHUNK = 64;
      int next1 = 0;
      for (int u = 0; u < U; ++u) {
        const int oi = u >> 1, h = u & 1, st = oi & 1;
        long long t_poll = 0;
#ifdef SCDEF_TC_TIMING
        const long long t_loop0 = clock64(), t_in1 = tacc[0] + tacc[1] + tacc[2] + tacc[4] + tacc[6];
#endif
        for (;;) {
          // this pair-tile's own rate GEMM: nothing else can happen before it
          if (next1 == oi) { if (issue_mma1(next1, nk1)) ++next1; continue; }
          if (GENES ? mbar_test(&mi->q_full, (uint32_t)u & 1u) : mbar_test_remote(&mi->q_full, (uint32_t)u & 1u)) break;
          // the next pair-tile's, a chunk at a time while Q is not there
          if (next1 < n_outer && next1 == oi + 1 && (ks1 > 0 || ready1(next1))) { if (issue_mma1(next1, MMA1_CHUNK)) ++next1; continue; }
          if (next1 > oi + 1 || next1 >= n_outer) {
            if (GENES) pair_wait(&mi->q_full, (uint32_t)u & 1u, 9);
            else pair_wait_remote(&mi->q_full, (uint32_t)u & 1u, 9);
            break;
          }
          if (t_poll == 0) t_poll = clock64();
          else if (clock64() - t_poll > PAIR_DEADLINE) pair_stuck(14, (int)blockIdx.x, u, next1);
        }
#ifdef SCDEF_TC_TIMING
        tacc[3] += (clock64() - t_loop0) - (tacc[0] + tacc[1] + tacc[2] + tacc[4] + tacc[6] - t_in1);
#endif
        if (GENES) {
          issue_mma2g(u, c2, n_outer);
          if (h == 1) c2.next(n_pt, S);
          continue;
        }
        const bool seg_first = first_in_seg(oi, c2) && h == 0, seg_last = last_in_seg(oi, c2) && h == 1;
        const uint32_t dzb = (uint32_t)c2.g & 1u;
        if (seg_first) PT_WAIT(7, pair_wait_remote(&mi->dz_empty[dzb], (((uint32_t)c2.g >> 1) & 1u) ^ 1u, 19));   // drained two segments ago
        tc_fence_after();
        PT_MARK();
        const uint32_t d = tmem + FLAT_TMEM_D2_LOCAL + dzb * FLAT_DZ_STRIDE;
        uint64_t da = dq_k, db = d_b2[st] + (uint64_t)(2 * h) * b2pl;
        umma2_bf16(d, da + qpl, db, idesc2, seg_first ? 0u : 1u);
        umma2_bf16(d, da, db + b2pl, idesc2, 1u);
        umma2_bf16(d, da, db, idesc2, 1u);
#pragma unroll
        for (int ks = 1; ks < TG / 16; ++ks) {
          da += 16;
          db += 16;
          umma2_bf16(d, da + qpl, db, idesc2, 1u);
          umma2_bf16(d, da, db + b2pl, idesc2, 1u);
          umma2_bf16(d, da, db, idesc2, 1u);
        }
        umma2_commit(&mi->q_empty);
        if (h == 1) umma2_commit(&mi->w_empty[st]);
        if (seg_last) umma2_commit(&mi->dz_full[dzb]);
        PT_LAP(5);
        if (h == 1) c2.next(n_pt, S);
      }
    }
    __syncwarp();
  } else {
    // =========================================== EPILOGUE (both CTAs) ====================================
    const int e = warp;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int r = (int)lane_base + lane;   // row of this CTA's unit: cell (LOCAL) or gene (GLOBAL)
    const int cq = e >> 2;                 // 16-column quarter of a 64-column half
    const uint32_t r_rempty[2] = {mapa_rank(&mi->r_empty[0], 0), mapa_rank(&mi->r_empty[1], 0)};
    const uint32_t r_dzempty[2] = {mapa_rank(&mi->dz_empty[0], 0), mapa_rank(&mi->dz_empty[1], 0)};
    const uint32_t r_qfull = mapa_rank(&mi->q_full, 0);
    float ll = 0.f;
    // one 64-column half: R (16 columns of this thread's row) -> q = x / R where x > 0; sum x log R
    auto half_math = [&](const float4 (&xv4)[4], const uint32_t (&Rr)[16], float (&q)[16]) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float xs[4] = {xv4[v].x, xv4[v].y, xv4[v].z, xv4[v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          // R >= 1e-30 wherever a count can be non-zero (samples are clipped at 1e-15); the clamp only keeps the padded
          // rows / columns (R = 0, x = 0) finite, so that x = 0 gives q = 0 and adds 0 to the sum without a select.
          // One FMNMX, one MUFU.RCP, one FMUL (+ one MUFU.LG2, one FFMA): the raw .ftz approximations, because
          // __fdividef / __logf wrap the same MUFU in ~6 / ~4 instructions of denormal handling that R never needs.
          constexpr int MV = GENES ? SCDEF_MATH_GLOBAL : SCDEF_MATH_LOCAL;
          if (MV > 0) {
            const float Rv = __uint_as_float(Rr[v * 4 + i]);
            const bool nz = xs[i] > 0.f;
            q[v * 4 + i] = nz ? (MV == 2 ? __fdividef(xs[i], Rv) : xs[i] * rcp_ftz(Rv)) : 0.f;
            if (WANT_LL) ll += nz ? xs[i] * lg2_ftz(Rv) : 0.f;
            continue;
          }
          const float Rc = fmaxf(__uint_as_float(Rr[v * 4 + i]), 1e-35f);
          q[v * 4 + i] = xs[i] * rcp_ftz(Rc);
          if (WANT_LL) ll = fmaf(xs[i], lg2_ftz(Rc), ll);   // in units of log 2: scaled once at the end
        }
      }
    };
    // split of 8 values into bf16 hi (round to nearest: F2FP, the conversion pipe is the epilogue's bottleneck together
    // with MUFU.RCP) and lo = the upper 16 bits of (x - hi) (truncation: one PRMT on the integer pipe).  hi + lo carries
    // 15 bits beyond hi's 8; the truncation error has the sign of lo, which is random because hi is rounded to nearest.
    auto split8t = [](const float* x, uint4& hi, uint4& lo) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
        const float r0 = x[2 * i] - __uint_as_float(h[i] << 16);
        const float r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
        l[i] = __byte_perm(__float_as_uint(r0), __float_as_uint(r1), 0x7632);   // {r1[31:16], r0[31:16]}
      }
      hi = make_uint4(h[0], h[1], h[2], h[3]);
      lo = make_uint4(l[0], l[1], l[2], l[3]);
    };
    auto store_q = [&](const float (&q)[16]) {
      uint4 hi, lo;
      const uint32_t off = L.q + (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u + (uint32_t)(cq * 2) * 128u;
      split8t(q, hi, lo);
      *reinterpret_cast<uint4*>(sm + off) = hi;
      *reinterpret_cast<uint4*>(sm + off + L.q_plane) = lo;
      split8t(q + 8, hi, lo);
      *reinterpret_cast<uint4*>(sm + off + 128u) = hi;
      *reinterpret_cast<uint4*>(sm + off + 128u + L.q_plane) = lo;
      fence_async_q();
      __syncwarp();
      if (lane == 0) {
        if (GENES) mbar_arrive(&mi->q_full);   // the Q hand-off is per CTA
        else mbar_arrive_remote(r_qfull);
      }
    };

    // ---- LOCAL: the segment's dz0 is complete: drain its accumulator (cell row r, every 4th 16-column chunk of k0) ----
    auto drain_segment = [&](const FlatPos& p) {
      const uint32_t dzb = (uint32_t)p.g & 1u;
      PT_WAIT(7, pair_wait(&mi->dz_full[dzb], ((uint32_t)p.g >> 1) & 1u, 13));
      tc_fence_after();
      PT_MARK();
      const int j = p.rb * CB + (int)rank * UC + r;
      const long long fseg = ((long long)p.rb * S + p.s) * n_pt;
      const int slot = chunk - flat_chunk_of(fseg, T, P);
      float* dst = ws.part + (((long long)p.s * gm.slots + slot) * a.B + j) * a.K0;
      const uint32_t tdz = tmem + FLAT_TMEM_D2_LOCAL + dzb * FLAT_DZ_STRIDE + (lane_base << 16);
      const bool vec4 = (a.K0 & 3) == 0, valid = j < a.B;
      for (int ch = cq; ch < KP / 16; ch += 4) {
        float v[16];
        tmem_ld16(tdz + (uint32_t)(ch * 16), v);
        if (valid) {
          if (vec4) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              if (ch * 16 + i < a.K0) *reinterpret_cast<float4*>(dst + ch * 16 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (ch * 16 + i < a.K0) dst[ch * 16 + i] = v[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_dzempty[dzb]);
      PT_LAP(8);
    };

    // ---- GLOBAL: fold dW of the segment into the two accumulators of the gene block (tensor memory) ----
    // thread = k0 row r (tensor-memory lane), the 32 genes [32 cq, 32 cq + 32) of this CTA's unit
    bool acc_live = false;   // the accumulators hold a partial sum of the current gene block
    const uint32_t tacc_mu = tmem + FLAT_TMEM_ACC_MU + (lane_base << 16), tacc_lw = tmem + FLAT_TMEM_ACC_LW + (lane_base << 16);
    auto fold_segment = [&](const FlatPos& p) {
      // W^s[k][genes] comes from the tile image in global memory (L2: the COPY thread streamed it a segment ago), NOT
      // from the ROW unit in shared memory: the unit is released by the segment's last rate GEMM, so the next segment's
      // unit loads while this one is still being processed and folded (no bubble in the tensor pipeline).  Four steps of
      // 8 genes: one 16 B load per plane and step, the next step's in flight, the first requested before the wait.
      const int g0 = (p.rb * 2 + (int)rank) * UC + cq * 32;
      const bool kok = r < a.K0;
      const float* tab = ws.foldtab + (long long)p.s * (2 + a.nb) * KP;
      const int rr = r < KP ? r : 0;
      // image of the 64-gene tile that holds these 32 genes; + plane * w_plane + step * 128
      const uint8_t* wline = ws.wimg + ((long long)p.s * a.n_wimg + (p.rb * 2 + (int)rank) * 2 + (cq >> 1)) * IL.w_unit +
                             (uint32_t)(rr >> 3) * W_RS + (uint32_t)((cq & 1) * 4) * 128u + (uint32_t)(rr & 7) * 16u;
      uint4 nh = __ldg(reinterpret_cast<const uint4*>(wline)), nl = __ldg(reinterpret_cast<const uint4*>(wline + IL.w_plane));
      // gene-scale samples of the step's 8 genes (batch 0; the same 32 B for every lane), fetched one step ahead as well
      const float* sline = a.smp_s + (long long)p.s * a.nb * a.G + g0;
      const bool s_ok = a.lik_on && g0 < a.G;   // G is a multiple of 8 and g0 of 32: a step is inside or outside
      float4 ns0 = make_float4(0.f, 0.f, 0.f, 0.f), ns1 = ns0;
      if (s_ok) { ns0 = __ldg(reinterpret_cast<const float4*>(sline)); ns1 = __ldg(reinterpret_cast<const float4*>(sline + 4)); }
      // A chunk boundary may cut a segment in two: the terms that do not depend on dW' (the W_0 prior gradient and the
      // rank-1 part) belong to the piece that holds the segment's first pair-tile only.
      const float own = (p.g > 0 || cur0.pt == 0) ? 1.f : 0.f;
      const float t_a1 = kok ? own * __ldg(tab + r) : 0.f, t_bm = kok ? own * __ldg(tab + KP + r) : 0.f;
      float t_ct[MAX_NB];
#pragma unroll
      for (int b = 0; b < MAX_NB; ++b) t_ct[b] = (b < a.nb && kok && a.lik_on) ? own * __ldg(tab + (2 + b) * KP + r) : 0.f;
      float* sg_sm = reinterpret_cast<float*>(sm + L.fold);
      if (a.nb > 1) {
        // gene-scale samples of the unit's 128 genes, batches 1 .. nb-1 (the same for every k0 row): one cooperative copy
        // per segment.  No warp can be a sub-unit ahead of another (the Q hand-off needs all 16), so nobody still reads
        // the previous segment's copy here.
        const int gu = (p.rb * 2 + (int)rank) * UC;
        for (int i = (int)threadIdx.x; i < a.nb * 128; i += E_THREADS) {
          const int b = i >> 7, gl = i & 127;
          sg_sm[i] = (a.lik_on && gu + gl < a.G) ? a.smp_s[((long long)p.s * a.nb + b) * a.G + gu + gl] : 0.f;
        }
        named_bar_sync(3, E_THREADS);
      }
      PT_WAIT(7, pair_wait(&mi->dz_full[0], (uint32_t)p.g & 1u, 13));
      tc_fence_after();
      PT_MARK();
      const uint32_t td2 = tmem + FLAT_TMEM_D2_GENES + (lane_base << 16);
#pragma unroll FOLD_UNROLL_N
      for (int t = 0; t < 4; ++t) {
        const uint4 ch = nh, cl = nl;
        const float4 cs0 = ns0, cs1 = ns1;
        if (t + 1 < 4) {
          nh = __ldg(reinterpret_cast<const uint4*>(wline + (t + 1) * 128));
          nl = __ldg(reinterpret_cast<const uint4*>(wline + IL.w_plane + (t + 1) * 128));
          if (s_ok && g0 + (t + 1) * 8 < a.G) {
            ns0 = __ldg(reinterpret_cast<const float4*>(sline + (t + 1) * 8));
            ns1 = __ldg(reinterpret_cast<const float4*>(sline + (t + 1) * 8 + 4));
          }
        }
        const int gb = g0 + t * 8;            // first of this step's 8 genes
        const uint32_t col = (uint32_t)(cq * 32 + t * 8);
        uint32_t dW[8], am[8], al[8];
        tmem_ld8_issue(td2 + col, dW);
        if (acc_live) {
          tmem_ld8_issue(tacc_mu + col, am);
          tmem_ld8_issue(tacc_lw + col, al);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { am[i] = 0u; al[i] = 0u; }
        }
        // sum_b ctS_b[k] s_b[g] for the 8 genes (the gene-scale samples: the same 32 B for every lane)
        float cs[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[i] = t_bm;
        if (gb < a.G) {
          {
            const float sv[8] = {cs0.x, cs0.y, cs0.z, cs0.w, cs1.x, cs1.y, cs1.z, cs1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = fmaf(t_ct[0], sv[i], cs[i]);
          }
#pragma unroll 1
          for (int b = 1; b < a.nb; ++b) {   // further batches: from the shared-memory copy made at the top of the fold
            const float* sp = sg_sm + b * 128 + cq * 32 + t * 8;
            const float4 s0 = *reinterpret_cast<const float4*>(sp), s1 = *reinterpret_cast<const float4*>(sp + 4);
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            float tc = 0.f;
#pragma unroll
            for (int bb = 1; bb < MAX_NB; ++bb) tc = (bb == b) ? t_ct[bb] : tc;
#pragma unroll
            for (int i = 0; i < 8; ++i) cs[i] = fmaf(tc, sv[i], cs[i]);
          }
        }
        float Pv[8], Rv[8];
        if (MAT) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            Pv[i] = (a.P && kok && gb + i < a.G) ? __ldg(a.P + (long long)r * a.G + gb + i) : a.P_scalar;
            Rv[i] = (a.R && kok && gb + i < a.G) ? __ldg(a.R + (long long)r * a.G + gb + i) : a.R_scalar;
          }
        }
        tmem_ld8_wait3(dW, am, al);
        const bool st_ok = kok && gb < a.G;   // G is a multiple of 8: a step is inside or outside
        const uint32_t hw[4] = {ch.x, ch.y, ch.z, ch.w}, lw[4] = {cl.x, cl.y, cl.z, cl.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t hv = hw[i >> 1], lv = lw[i >> 1];
          const float W = (i & 1) ? (bf16hi(hv) + bf16hi(lv)) : (bf16lo(hv) + bf16lo(lv));
          float a1 = t_a1, c1 = cs[i];
          if (MAT) {
            c1 = cs[i] - t_bm + a.global_weight * Rv[i] * (float)a.K0 * t_bm;   // gw B, B = R K0 / (tau omega); table: own / (tau omega)
            a1 = a.global_weight * (Pv[i] * t_a1 - own);                       // gw (A - 1), A = P / tau; table: own / tau
          }
          const float tt = fmaf(W, fmaf(a.scale, __uint_as_float(dW[i]), -c1), a1);
          // inside the sample clip <=> |log2 W| < log2(0.9999e15) (the clip is symmetric in the log; W = 0 of the padding
          // gives -inf); the second accumulator is kept in units of log 2 and scaled when the block is flushed
          const float lgW = lg2_ftz(W);
          if (fabsf(lgW) < LG2_V_HI_M && st_ok) {
            am[i] = __float_as_uint(__uint_as_float(am[i]) + tt);
            al[i] = __float_as_uint(fmaf(tt, lgW, __uint_as_float(al[i])));
          }
        }
        tmem_st8(tacc_mu + col, am);
        tmem_st8(tacc_lw + col, al);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      acc_live = true;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&mi->dz_empty[0]);   // D2 may be overwritten by this CTA's next segment
      PT_LAP(8);
    };
    auto flush_block = [&](const FlatPos& p) {
      // partial sums of gene block p.rb held by this pair -> slot of the (slots, 2, K0, G) buffer
      const int g0 = (p.rb * 2 + (int)rank) * UC + cq * 32;
      const long long per_rb = (long long)S * n_pt;
      const int slot = chunk - flat_chunk_of((long long)p.rb * per_rb, T, P);
      const long long n = (long long)a.K0 * a.G;
      float* pm = ws.part + ((long long)slot * 2 + 0) * n + (long long)r * a.G + g0;
      float* pl = ws.part + ((long long)slot * 2 + 1) * n + (long long)r * a.G + g0;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        float am[16], al[16];
        tmem_ld16(tacc_mu + (uint32_t)(cq * 32 + t * 16), am);
        tmem_ld16(tacc_lw + (uint32_t)(cq * 32 + t * 16), al);
        if (r < a.K0) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            if (g0 + t * 16 + i < a.G) {   // G is a multiple of 8
              *reinterpret_cast<float4*>(pm + t * 16 + i) = make_float4(am[i], am[i + 1], am[i + 2], am[i + 3]);
              *reinterpret_cast<float4*>(pl + t * 16 + i) = make_float4(LN2_F * al[i], LN2_F * al[i + 1], LN2_F * al[i + 2], LN2_F * al[i + 3]);
            }
          }
        }
      }
      acc_live = false;
    };

    // this thread's first float4 of 64-column tile t of the count image sits at x_base + t * x_tile
    const long long x_tile = (long long)gm.n_units * (4 * 16 * 128);
    const float* x_img = GENES ? ws.ximg_t : ws.ximg;
#ifdef SCDEF_LOCAL_XB_AHEAD
    constexpr bool XB_AHEAD = !GENES;
#else
    constexpr bool XB_AHEAD = false;
#endif
    float4 xa[4], xb[4];   // counts of half 0 / half 1, loaded one pair-tile ahead
    FlatPos cc = cur0;
    auto x_ptr = [&](const FlatPos& p) {
      const int unit = p.rb * 2 + (int)rank;
      return x_img + (((long long)unit * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4 + (long long)(2 * p.pt) * x_tile;
    };
    // Half 0 a pair-tile ahead, half 1 at the top of its own pair-tile (it has the first half's math and Q store, ~2k
    // cycles, to arrive): holding both halves a pair-tile ahead costs 16 more live registers and spills in the math.
    if (n_outer > 0) {
      const float* xi = x_ptr(cc);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        xa[v] = *reinterpret_cast<const float4*>(xi + v * 128);
        if (XB_AHEAD) xb[v] = *reinterpret_cast<const float4*>(xi + x_tile + v * 128);
      }
    }
    for (int oi = 0; oi < n_outer; ++oi) {
      const int st = oi & 1;
      const uint32_t par = (uint32_t)(oi >> 1) & 1u;
      const bool more = oi + 1 < n_outer;
      FlatPos cn = cc;
      cn.next(n_pt, S);
      const float* xnext = x_ptr(cn);
      const int rs = GENES ? 0 : st;
      const uint32_t rpar = GENES ? ((uint32_t)oi & 1u) : par;
      const uint32_t taddr = tmem + (lane_base << 16) + (uint32_t)(rs * (2 * TG) + cq * 16);
      PT_MARK();
      PT_WAIT(0, pair_wait(&mi->r_full[rs], rpar, 11));
      tc_fence_after();
      PT_MARK();
      uint32_t R0[16], R1[16];
      float q[16];
      tmem_ld16_issue(taddr, R0);
      if (!XB_AHEAD) {
        const float* xi = x_ptr(cc) + x_tile;
#pragma unroll
        for (int v = 0; v < 4; ++v) xb[v] = *reinterpret_cast<const float4*>(xi + v * 128);
      }
      tmem_ld16_wait(R0);
      PT_LAP(3);
      half_math(xa, R0, q);
      // the next pair-tile's counts are fetched while this one is processed -- except across a fold (GLOBAL), whose
      // working set needs the registers: there they are fetched after it
      const bool x_early = more && !(GENES && last_in_seg(oi, cc));
      if (x_early) {
#pragma unroll
        for (int v = 0; v < 4; ++v) xa[v] = *reinterpret_cast<const float4*>(xnext + v * 128);
      }
      tmem_ld16_issue(taddr + (uint32_t)TG, R1);   // in flight while the first half's Q tile is stored
      PT_LAP(4);
      PT_WAIT(1, pair_wait(&mi->q_empty, 1u, 12));  // sub-unit 2 oi: parity (u & 1) ^ 1 with u even
      PT_MARK();
      store_q(q);
      tmem_ld16_wait(R1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_rempty[rs]);   // both halves of R are in registers
      PT_LAP(5);
      half_math(xb, R1, q);
      if (x_early && XB_AHEAD) {
#pragma unroll
        for (int v = 0; v < 4; ++v) xb[v] = *reinterpret_cast<const float4*>(xnext + x_tile + v * 128);
      }
      PT_LAP(4);
      PT_WAIT(1, pair_wait(&mi->q_empty, 0u, 12));  // sub-unit 2 oi + 1
      PT_MARK();
      store_q(q);
      PT_LAP(5);
      if (last_in_seg(oi, cc)) {
        if (GENES) {
          fold_segment(cc);
          if (!more || cn.rb != cc.rb) flush_block(cc);
          if (more) {
#pragma unroll
            for (int v = 0; v < 4; ++v) xa[v] = *reinterpret_cast<const float4*>(xnext + v * 128);
          }
        } else {
          drain_segment(cc);
        }
      }
      cc = cn;
    }
    if (WANT_LL) {
      double v = warp_sum_d((double)ll * 0.69314718055994530942);
      if (lane == 0) mi->red[e] = v;
      named_bar_sync(2, E_THREADS);
      if (e == 0 && lane == 0) {
        double tot = 0.0;
        for (int i = 0; i < E_WARPS; ++i) tot += mi->red[i];
        atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
      }
    }
  }

#ifdef SCDEF_TC_TIMING
  if ((int)(blockIdx.x >> 1) == (int)(gridDim.x >> 2) && lane == 0 &&
      (warp == COPY_WARP || warp == MMA_WARP || warp == 0 || warp == 15)) {
    const int role = warp == COPY_WARP ? 0 : (warp == MMA_WARP ? 1 : (warp == 0 ? 2 : 3));
    for (int i = 0; i < 15; ++i) g_pair_timing[rank][role][i] = (unsigned long long)tacc[i];
    g_pair_timing[rank][role][15] = (unsigned long long)(clock64() - t_start);
  }
#endif
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
