// Warp-specialised, mbarrier-pipelined version of the fused likelihood kernel (included by
// lik_tc.cu inside its anonymous namespace; uses its PTX wrappers and layouts).
//
// One CTA = 13 warps:
//   warps 0-3  PRODUCERS : sample the W^s tile (Philox / explicit eps -> bf16 hi/lo, UMMA layout) and,
//                          in the GLOBAL pass, stage z0^s blocks; side sums for the W_0 prior / gene scale
//   warps 4-11 EPILOGUE  : TMEM -> registers, Poisson terms, Q tile back to smem; GLOBAL: fold dW
//   warp  12   MMA ISSUER: one thread issues every tcgen05.mma and tcgen05.commit
//   warp  13   COPY      : one thread streams pre-split z0^s blocks into smem with cp.async.bulk (TMA)
// Work is cut into UNITS of 128 cells x 64 genes.  W tiles, z blocks and the R accumulators are
// double-buffered, so generating tile/sample n+1, the tensor-core work of unit u+1 and the
// epilogue of unit u overlap.  All hand-offs are mbarriers (full/empty pairs).
//
//   LOCAL  CTA = (MC sample s, gene range, 256-cell block): units = (tile, half); z static in smem;
//          dz0[half] += Q W^T accumulates in TMEM over the tiles of the range.
//   GLOBAL CTA = (gene tile, chunk of MC samples): units = (s, 128-cell block); dW[s&1] += z0^T Q over the
//          blocks of s; epilogue 2 folds W.dW into per-thread d mu / d rho accumulators over the chunk.
#pragma once

constexpr int PIPE_NT = 448;           // 14 warps
constexpr int P_THREADS = 128, E_THREADS = 256;

struct PipeMisc {
  uint64_t w_full[2], w_empty[2], z_full[2], z_empty[2], r_full[2], r_empty[2], q_full, q_empty, dw_full[2],
      dw_empty[2], dz_full;
  uint32_t tmem_base, pad;
  double red[32];
  float rowU[2][128], rowW[2][128];
  float colacc[2][MAX_NB][TG];
  float ct_sm[MAX_NB][128];     // ctilde of the sample being generated (GLOBAL)
};

#ifdef SCDEF_TC_TIMING
__device__ unsigned long long g_tc_timing[2][4][16];  // [kernel variant][role P/MMA/E/COPY][slot]
#define TWAIT(slot, bar, par) do { long long _t0 = clock64(); mbar_wait(bar, par); tacc[slot] += clock64() - _t0; } while (0)
#else
#define TWAIT(slot, bar, par) mbar_wait(bar, par)
#endif

struct PipeSmem {
  uint32_t z[2], w[2], q, misc, total;   // each z / w / q region = hi plane followed by lo plane
  uint32_t z_rs, z_plane, w_plane, q_plane;
  uint32_t w_unit;                       // bytes of one W image unit: hi | lo | gene-scale tile (MAX_NB x TG floats)
};
__host__ __device__ inline PipeSmem plan_pipe_smem(int KP) {
  PipeSmem s;
  s.z_rs = (uint32_t)(KP / 8) * 128u;
  s.z_plane = 16u * s.z_rs;                 // 128 cells
  s.w_plane = (uint32_t)(KP / 8) * W_RS;    // KP rows x 64 genes
  s.q_plane = 16u * Q_RS;                   // 128 cells x 64 genes
  uint32_t o = 0;
  s.z[0] = o; o += 2 * s.z_plane;
  s.z[1] = o; o += 2 * s.z_plane;
  s.w_unit = 2 * s.w_plane + (uint32_t)(MAX_NB * TG * 4);
  s.w[0] = o; o += s.w_unit;
  s.w[1] = o; o += s.w_unit;
  s.q = o; o += 2 * s.q_plane;
  s.misc = o;
  s.total = o + (uint32_t)((sizeof(PipeMisc) + 1023) / 1024 * 1024);
  return s;
}


__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D), completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- producers -------------------------------------------------------------------------------------
// GLOBAL side work on a landed W tile: colacc[b][g] = sum_k ctilde[b][k] W[k][g]  (W = hi + lo from smem)
__device__ __forceinline__ void pipe_col_w(const LikTcArgs& a, const uint8_t* wbuf, uint32_t w_plane, PipeMisc* mi, int stage,
                                           int KP, int tid) {
  const int nkg = KP / 8;
  const int kk = tid & 7, cg = (tid >> 3) & 7;
  float col[MAX_NB][8];
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b)
#pragma unroll
    for (int e = 0; e < 8; ++e) col[b][e] = 0.f;
  for (int kg = tid >> 6; kg < nkg; kg += P_THREADS / 64) {
    const int k = kg * 8 + kk;
    if (k >= a.K0) continue;
    const uint32_t off = (uint32_t)kg * W_RS + (uint32_t)cg * 128u + (uint32_t)kk * 16u;
    const uint4 hv = *reinterpret_cast<const uint4*>(wbuf + off), lv = *reinterpret_cast<const uint4*>(wbuf + w_plane + off);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
    float w[8];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      w[2 * p] = bf16lo(hw[p]) + bf16lo(lw[p]);
      w[2 * p + 1] = bf16hi(hw[p]) + bf16hi(lw[p]);
    }
#pragma unroll
    for (int b = 0; b < MAX_NB; ++b) {
      if (b < a.nb) {
        const float ct = mi->ct_sm[b][k];
#pragma unroll
        for (int e = 0; e < 8; ++e) col[b][e] = fmaf(ct, w[e], col[b][e]);
      }
    }
  }
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b) {
    if (b < a.nb) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = col[b][e];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (kk == 0) atomicAdd(&mi->colacc[stage][b][cg * 8 + e], v);
      }
    }
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------
template <bool GLOBAL, bool WANT_LL>
__global__ void __launch_bounds__(PIPE_NT, 1) k_lik_pipe(const LikTcArgs a, const Geom gm, const TcScratch ws, double* loss_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const PipeSmem L = plan_pipe_smem(KP);
  PipeMisc* mi = reinterpret_cast<PipeMisc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef SCDEF_TC_TIMING
  long long tacc[16];
  for (int i = 0; i < 16; ++i) tacc[i] = 0;
  const long long t_start = clock64();
#endif

  // ---- work of this CTA ----
  int s_fixed = 0, tile_fixed = 0, outer_begin, outer_end, j_block0 = 0, units_per_outer;
  if (GLOBAL) {
    tile_fixed = blockIdx.x % gm.n_tiles;
    const int chunk = blockIdx.x / gm.n_tiles;
    outer_begin = (int)((long long)chunk * a.S / gm.n_chunks);
    outer_end = (int)((long long)(chunk + 1) * a.S / gm.n_chunks);
    units_per_outer = (a.B + UC - 1) / UC;
  } else {
    s_fixed = blockIdx.x / gm.NG;
    const int range = blockIdx.x % gm.NG;
    outer_begin = range * gm.tiles_per_range;
    outer_end = min(gm.n_tiles, outer_begin + gm.tiles_per_range);
    j_block0 = blockIdx.y * CB;
    units_per_outer = (min(CB, a.B - j_block0) > UC) ? 2 : 1;
  }
  const int n_outer = outer_end - outer_begin;
  const int U = n_outer * units_per_outer;

  // ---- prologue ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mi->w_full[i], 1);
      mbar_init(&mi->w_empty[i], GLOBAL ? E_THREADS + P_THREADS : 1);
      mbar_init(&mi->z_full[i], 1);
      mbar_init(&mi->z_empty[i], 1);
      mbar_init(&mi->r_full[i], 1);
      mbar_init(&mi->r_empty[i], E_THREADS);
      mbar_init(&mi->dw_full[i], 1);
      mbar_init(&mi->dw_empty[i], E_THREADS);
    }
    mbar_init(&mi->q_full, E_THREADS);
    mbar_init(&mi->q_empty, 1);
    mbar_init(&mi->dz_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * 128; i += PIPE_NT) { (&mi->rowU[0][0])[i] = 0.f; (&mi->rowW[0][0])[i] = 0.f; }
  for (int i = threadIdx.x; i < 2 * MAX_NB * TG; i += PIPE_NT) (&mi->colacc[0][0][0])[i] = 0.f;
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&mi->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int units_total_B = (a.B + UC - 1) / UC;  // z image: (S, units_total_B, 2 planes)
  const uint32_t z_bytes = 2u * L.z_plane;
  const uint32_t tmem = mi->tmem_base;
  const uint32_t tmem_acc2 = tmem + 2 * TG;  // LOCAL: dz[half] (KP cols each); GLOBAL: dW[stage] (TG cols each)

  if (warp < 4) {
    // =========================================== SIDE WARPS ===========================================
    // GLOBAL: gene-scale likelihood gradient  G_s += scale * (colsumX / s - ctilde^T W)  per (tile, sample)
    if (GLOBAL) {
      const int tid = threadIdx.x;
      const int g0 = tile_fixed * TG;
      for (int oi = 0; oi < n_outer; ++oi) {
        const int st = oi & 1;
        const int s = outer_begin + oi;
        const float* ct = ws.ctilde + (long long)s * a.nb * a.K0;
        for (int i = tid; i < a.nb * a.K0; i += P_THREADS) mi->ct_sm[i / a.K0][i % a.K0] = ct[i];
        named_bar_sync(1, P_THREADS);
        TWAIT(0, &mi->w_full[st], (oi >> 1) & 1);
        pipe_col_w(a, sm + L.w[st], L.w_plane, mi, st, KP, tid);
        named_bar_sync(1, P_THREADS);
        if (tid < TG) {
          const int g = g0 + tid;
          const float* stile = reinterpret_cast<const float*>(sm + L.w[st] + 2 * L.w_plane);
          for (int b = 0; b < a.nb; ++b) {
            if (g < a.G && a.G_s) {
              const float sv = stile[b * TG + tid];
              a.G_s[((long long)s * a.nb + b) * a.G + g] += a.scale * (ws.colsumX[(long long)b * a.G + g] / sv - mi->colacc[st][b][tid]);
            }
            mi->colacc[st][b][tid] = 0.f;
          }
        }
        mbar_arrive(&mi->w_empty[st]);
      }
    }
  } else if (warp == 13) {
    // =========================================== COPY (TMA) ===========================================
    if (lane == 0) {
      const uint32_t w_bytes = 2u * L.w_plane + (uint32_t)(a.nb * TG * 4);
      if (!GLOBAL) {
        // z0^s of the 256-cell block is static: both halves land once, on z_full[0]
        mbar_arrive_expect_tx(&mi->z_full[0], z_bytes * (uint32_t)units_per_outer);
        for (int h = 0; h < units_per_outer; ++h)
          bulk_g2s(sbase + L.z[h], ws.zimg + ((long long)s_fixed * units_total_B + (j_block0 / UC + h)) * z_bytes, z_bytes,
                   &mi->z_full[0]);
      }
      int u = 0;
      for (int oi = 0; oi < n_outer; ++oi) {
        const int st = oi & 1;
        const int s = GLOBAL ? outer_begin + oi : s_fixed;
        const int tile = GLOBAL ? tile_fixed : outer_begin + oi;
        TWAIT(1, &mi->w_empty[st], ((oi >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&mi->w_full[st], w_bytes);
        bulk_g2s(sbase + L.w[st], ws.wimg + ((long long)s * gm.n_tiles + tile) * L.w_unit, w_bytes, &mi->w_full[st]);
        if (GLOBAL) {
          for (int cb = 0; cb < units_per_outer; ++cb, ++u) {
            const int zs = u & 1;
            TWAIT(0, &mi->z_empty[zs], ((u >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&mi->z_full[zs], z_bytes);
            bulk_g2s(sbase + L.z[zs], ws.zimg + ((long long)s * units_total_B + cb) * z_bytes, z_bytes, &mi->z_full[zs]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 12) {
    // =========================================== MMA ISSUER ===========================================
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc(128, TG, 0, 1);
      const uint32_t idesc2 = GLOBAL ? umma_idesc(128, TG, 1, 1) : umma_idesc(128, KP, 0, 0);
      auto issue_mma1 = [&](int u) {
        const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
        const int st = oi & 1, rs = u & 1, zb = GLOBAL ? (u & 1) : cb;
        if (cb == 0) TWAIT(0, &mi->w_full[st], (oi >> 1) & 1);
        if (GLOBAL) TWAIT(1, &mi->z_full[zb], (u >> 1) & 1);
        TWAIT(2, &mi->r_empty[rs], ((u >> 1) & 1) ^ 1);
        tc_fence_after();
        uint32_t acc = 0;
        for (int ks = 0; ks < KP / 16; ++ks) {
          const uint32_t za = sbase + L.z[zb] + (uint32_t)ks * 256u, wb = sbase + L.w[st] + (uint32_t)ks * 2u * W_RS;
          const uint64_t a_hi = umma_desc(za, 128u, L.z_rs), a_lo = umma_desc(za + L.z_plane, 128u, L.z_rs);
          const uint64_t b_hi = umma_desc(wb, W_RS, 128u), b_lo = umma_desc(wb + L.w_plane, W_RS, 128u);
          umma_bf16(tmem + rs * TG, a_lo, b_hi, idesc1, acc);
          umma_bf16(tmem + rs * TG, a_hi, b_lo, idesc1, 1u);
          umma_bf16(tmem + rs * TG, a_hi, b_hi, idesc1, 1u);
          acc = 1u;
        }
        umma_commit(&mi->r_full[rs]);
      };
      if (!GLOBAL) mbar_wait(&mi->z_full[0], 0);
      if (U > 0) issue_mma1(0);
      for (int u = 0; u < U; ++u) {
        if (u + 1 < U) issue_mma1(u + 1);
        const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
        const int st = oi & 1;
        TWAIT(3, &mi->q_full, u & 1);
        if (GLOBAL && cb == 0) TWAIT(4, &mi->dw_empty[st], ((oi >> 1) & 1) ^ 1);
        tc_fence_after();
        if (GLOBAL) {
          // dW[st] (+)= z0^T Q : A = z (MN-major, M = k0), B = Q (MN-major, N = genes), K = cells of the unit
          const int rows = min(UC, a.B - cb * UC), nks = (rows + 15) / 16;
          const int zb = u & 1;
          uint32_t acc = cb > 0 ? 1u : 0u;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t za = sbase + L.z[zb] + (uint32_t)ks * 2u * L.z_rs, qb = sbase + L.q + (uint32_t)ks * 2u * Q_RS;
            const uint64_t a_hi = umma_desc(za, L.z_rs, 128u), a_lo = umma_desc(za + L.z_plane, L.z_rs, 128u);
            const uint64_t b_hi = umma_desc(qb, Q_RS, 128u), b_lo = umma_desc(qb + L.q_plane, Q_RS, 128u);
            umma_bf16(tmem_acc2 + st * TG, a_lo, b_hi, idesc2, acc);
            umma_bf16(tmem_acc2 + st * TG, a_hi, b_lo, idesc2, 1u);
            umma_bf16(tmem_acc2 + st * TG, a_hi, b_hi, idesc2, 1u);
            acc = 1u;
          }
          umma_commit(&mi->q_empty);
          umma_commit(&mi->z_empty[zb]);
          if (cb == units_per_outer - 1) umma_commit(&mi->dw_full[st]);
        } else {
          // dz[cb] (+)= Q W^T : A = Q (K-major), B = W (K-major, N = k0), K = genes of the tile
          uint32_t acc = oi > 0 ? 1u : 0u;
          for (int ks = 0; ks < TG / 16; ++ks) {
            const uint32_t qa = sbase + L.q + (uint32_t)ks * 256u, wb = sbase + L.w[st] + (uint32_t)ks * 256u;
            const uint64_t a_hi = umma_desc(qa, 128u, Q_RS), a_lo = umma_desc(qa + L.q_plane, 128u, Q_RS);
            const uint64_t b_hi = umma_desc(wb, 128u, W_RS), b_lo = umma_desc(wb + L.w_plane, 128u, W_RS);
            umma_bf16(tmem_acc2 + cb * KP, a_lo, b_hi, idesc2, acc);
            umma_bf16(tmem_acc2 + cb * KP, a_hi, b_lo, idesc2, 1u);
            umma_bf16(tmem_acc2 + cb * KP, a_hi, b_hi, idesc2, 1u);
            acc = 1u;
          }
          umma_commit(&mi->q_empty);
          if (cb == units_per_outer - 1) umma_commit(&mi->w_empty[st]);
        }
      }
      if (!GLOBAL) umma_commit(&mi->dz_full);
    }
    __syncwarp();
  } else {
    // =========================================== EPILOGUE ===========================================
    const int e = warp - 4;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int r = (int)lane_base + lane;  // row of the unit (cell) / k0 row in epilogue 2
    const int hc = e >> 2;                // 32-gene half of the tile
    float ll = 0.f;
    float esum0 = 0.f, esum1 = 0.f;
    const uint32_t tmem_x = tmem + 4 * TG;  // GLOBAL: the X tile of the first two 128-cell units (2 x TG columns)
    if (GLOBAL) {
      for (int cb = 0; cb < min(units_per_outer, 2); ++cb) {
        const int j = cb * UC + r;
        const int gq = tile_fixed * TG + hc * 32;
        float xv[32];
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < a.B && gq + v * 4 < a.G) t4 = *reinterpret_cast<const float4*>(a.X + ws.xoff[j] + gq + v * 4);
          xv[v * 4] = t4.x; xv[v * 4 + 1] = t4.y; xv[v * 4 + 2] = t4.z; xv[v * 4 + 3] = t4.w;
        }
        tmem_st16(tmem_x + (lane_base << 16) + (uint32_t)(cb * TG + hc * 32), xv);
        tmem_st16(tmem_x + (lane_base << 16) + (uint32_t)(cb * TG + hc * 32 + 16), xv + 16);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    float acc_mu[32], acc_lw[32];
    if (GLOBAL) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { acc_mu[i] = 0.f; acc_lw[i] = 0.f; }
    }
    // row context of a unit: (c, batch row of the gene-scale sample, row of X); the next unit's
    // context is fetched while the current one is processed (no dependent-load chains in the loop)
    struct Ctx { float c; int b; long long xoff; bool valid; };
    auto fetch_ctx = [&](int u) {
      Ctx cx; cx.c = 0.f; cx.b = 0; cx.xoff = 0; cx.valid = false;
      if (u < U) {
        const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
        const int s = GLOBAL ? outer_begin + oi : s_fixed;
        const int j = (GLOBAL ? cb * UC : j_block0 + cb * UC) + r;
        if (j < a.B) {
          cx.valid = true;
          cx.c = a.smp_c[(long long)s * a.B + j];
          cx.b = ws.brow[j];
          cx.xoff = ws.xoff[j];
        }
      }
      return cx;
    };
    Ctx nxt = fetch_ctx(0);
    for (int u = 0; u < U; ++u) {
      const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
      const int st = oi & 1, rs = u & 1;
      const int s = GLOBAL ? outer_begin + oi : s_fixed;
      const int g0 = (GLOBAL ? tile_fixed : outer_begin + oi) * TG + hc * 32;
      const int j = (GLOBAL ? cb * UC : j_block0 + cb * UC) + r;
      const Ctx cx = nxt;
      const bool valid = cx.valid;
      const float c = cx.c;
      const float* xrow = a.X + cx.xoff;
      const bool x_in_tmem = GLOBAL && cb < 2;
      float es = 0.f;
      float4 x4[8];
      if (!x_in_tmem) {
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int g = g0 + v * 4;
          x4[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid && g < a.G) x4[v] = *reinterpret_cast<const float4*>(xrow + g);
        }
      }
      nxt = fetch_ctx(u + 1);
      if (cb == 0) mbar_wait(&mi->w_full[st], (oi >> 1) & 1);  // the gene-scale tile rides with the W image
      TWAIT(0, &mi->r_full[rs], (u >> 1) & 1);
      tc_fence_after();
      TWAIT(1, &mi->q_empty, (u & 1) ^ 1);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float R[16];
        tmem_ld16(tmem + (lane_base << 16) + (uint32_t)(rs * TG + hc * 32 + cc * 16), R);
        float q[16];
        float xs16[16];
        if (x_in_tmem) {
          tmem_ld16(tmem_x + (lane_base << 16) + (uint32_t)(cb * TG + hc * 32 + cc * 16), xs16);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            xs16[v * 4 + 0] = x4[cc * 4 + v].x; xs16[v * 4 + 1] = x4[cc * 4 + v].y;
            xs16[v * 4 + 2] = x4[cc * 4 + v].z; xs16[v * 4 + 3] = x4[cc * 4 + v].w;
          }
        }
        const float* srow_sm = reinterpret_cast<const float*>(sm + L.w[st] + 2 * L.w_plane) + cx.b * TG + hc * 32 + cc * 16;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          // rows / genes outside the problem carry c = 0 / s = 0 and x = 0: Lambda = E = Q = 0 without masks
          const float4 s4v = *reinterpret_cast<const float4*>(srow_sm + v * 4);
          const float ss[4] = {s4v.x, s4v.y, s4v.z, s4v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float Rv = R[v * 4 + i];
            const float xv = xs16[v * 4 + i];
            const float lam = (c * ss[i]) * Rv;
            const float E = xv - lam;
            q[v * 4 + i] = __fdividef(E, fmaxf(Rv, 1e-37f));
            if (WANT_LL) {
              const float t = xv * __logf(lam);   // lam = 0 only where x = 0: the select discards -inf * 0
              ll += (xv > 0.f ? t : 0.f) - lam;
            }
            es += E;
          }
        }
        if (g_dbg_R && valid) {
          for (int i = 0; i < 16; ++i)
            if (g0 + cc * 16 + i < a.G) g_dbg_R[((long long)s * a.B + j) * a.G + g0 + cc * 16 + i] = R[i];
        }
        uint4 hi, lo;
        const uint32_t off = L.q + (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u + (uint32_t)(hc * 4 + cc * 2) * 128u;
        split8(q, hi, lo);
        *reinterpret_cast<uint4*>(sm + off) = hi;
        *reinterpret_cast<uint4*>(sm + off + L.q_plane) = lo;
        split8(q + 8, hi, lo);
        *reinterpret_cast<uint4*>(sm + off + 128u) = hi;
        *reinterpret_cast<uint4*>(sm + off + 128u + L.q_plane) = lo;
      }
      tc_fence_before();
      mbar_arrive(&mi->r_empty[rs]);
      fence_async_smem();
      mbar_arrive(&mi->q_full);
      if (cb & 1) esum1 += es; else esum0 += es;

      if (GLOBAL && cb == units_per_outer - 1) {
        // ---- epilogue 2: fold dW^s (and the W_0 prior gradient) into d mu, sum t log W ----
        TWAIT(2, &mi->dw_full[st], (oi >> 1) & 1);
        tc_fence_after();
        float dW[32];
        tmem_ld16(tmem_acc2 + (lane_base << 16) + (uint32_t)(st * TG + hc * 32), dW);
        tmem_ld16(tmem_acc2 + (lane_base << 16) + (uint32_t)(st * TG + hc * 32 + 16), dW + 16);
        tc_fence_before();
        mbar_arrive(&mi->dw_empty[st]);
        if (r < a.K0 && !a.w0_stopped) {
          float A = a.P_scalar, Bm = a.R_scalar * (float)a.K0;
          if (a.use_brd) {
            const float tau = a.smp_tau[(long long)s * a.K0 + r], om = a.smp_om[(long long)s * a.K0 + r];
            A = a.P_scalar / tau;
            Bm = a.R_scalar * (float)a.K0 / (tau * om);
          }
          const uint32_t woff = L.w[st] + (uint32_t)(r >> 3) * W_RS + (uint32_t)(r & 7) * 16u;
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const uint4 hv = *reinterpret_cast<const uint4*>(sm + woff + (uint32_t)(hc * 4 + c8) * 128u);
            const uint4 lv = *reinterpret_cast<const uint4*>(sm + woff + L.w_plane + (uint32_t)(hc * 4 + c8) * 128u);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                const int i = c8 * 8 + p * 2 + q2;
                const float W = q2 ? (bf16hi(hw[p]) + bf16hi(lw[p])) : (bf16lo(hw[p]) + bf16lo(lw[p]));
                if (W > V_LO_M && W < V_HI_M) {
                  const float t = a.scale * dW[i] * W + a.global_weight * (A - 1.f - Bm * W);
                  acc_mu[i] += t;
                  acc_lw[i] = fmaf(t, __logf(W), acc_lw[i]);
                }
              }
            }
          }
        }
        mbar_arrive(&mi->w_empty[st]);
      }
    }

    if (GLOBAL) {
      const int chunk = blockIdx.x / gm.n_tiles;
      const int gq = tile_fixed * TG + hc * 32;
      if (r < a.K0 && !a.w0_stopped) {
        float* pm = ws.part + (((long long)chunk * 2 + 0) * a.K0 + r) * a.G + gq;
        float* pr = ws.part + (((long long)chunk * 2 + 1) * a.K0 + r) * a.G + gq;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (gq + i < a.G) {
            *reinterpret_cast<float4*>(pm + i) = make_float4(acc_mu[i], acc_mu[i + 1], acc_mu[i + 2], acc_mu[i + 3]);
            *reinterpret_cast<float4*>(pr + i) = make_float4(acc_lw[i], acc_lw[i + 1], acc_lw[i + 2], acc_lw[i + 3]);
          }
        }
      }
    } else {
      // ---- drain dz0[half] (this thread: cell row r, k0 columns of its half hc) ----
      mbar_wait(&mi->dz_full, 0);
      tc_fence_after();
      const int range = blockIdx.x % gm.NG;
      for (int h = 0; h < units_per_outer; ++h) {
        const int j = j_block0 + h * UC + r;
        float* dst = ws.part + (((long long)s_fixed * gm.NG + range) * a.B + j) * a.K0;
        for (int cc = hc; cc < KP / 16; cc += 2) {
          float v[16];
          tmem_ld16(tmem_acc2 + (lane_base << 16) + (uint32_t)(h * KP + cc * 16), v);
          if (j < a.B) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cc * 16 + i < a.K0) dst[cc * 16 + i] = v[i];
          }
        }
        const float eh = h ? esum1 : esum0;
        if (j < a.B && eh != 0.f) atomicAdd(&a.G_c[(long long)s_fixed * a.B + j], a.scale * eh / a.smp_c[(long long)s_fixed * a.B + j]);
      }
    }
    // loss: reduce over the 256 epilogue threads
    double v = warp_sum_d((double)ll);
    if (lane == 0) mi->red[e] = v;
    named_bar_sync(2, E_THREADS);
    if (e == 0 && lane == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += mi->red[i];
      atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
    }
  }

#ifdef SCDEF_TC_TIMING
  if (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == 12 || warp == 4 || warp == 13)) {
    const int role = warp == 0 ? 0 : (warp == 12 ? 1 : (warp == 4 ? 2 : 3));
    for (int i = 0; i < 15; ++i) g_tc_timing[GLOBAL ? 1 : 0][role][i] = (unsigned long long)tacc[i];
    g_tc_timing[GLOBAL ? 1 : 0][role][15] = (unsigned long long)(clock64() - t_start);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
