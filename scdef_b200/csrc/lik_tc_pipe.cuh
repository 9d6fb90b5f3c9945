// Warp-specialised, mbarrier-pipelined version of the fused likelihood kernel (included by
// lik_tc.cu inside its anonymous namespace; uses its PTX wrappers and layouts).
//
// One CTA = 18 warps:
//   warps 0-15 EPILOGUE  : TMEM -> registers, Poisson terms, Q tile back to smem; GLOBAL: fold dW.
//                          warp w owns TMEM lanes 32*(w%4).. (cell rows / k0 rows) and 16 gene columns (w/4)
//   warp  16   MMA ISSUER: one thread issues every tcgen05.mma and tcgen05.commit
//   warp  17   COPY      : one thread streams the pre-sampled W_0 tile images (k_tc_wgen) and the pre-split
//                          z0^s blocks (k_tc_zprep) into smem with cp.async.bulk (TMA)
// Work is cut into UNITS of 128 cells x 64 genes.  W tiles, z blocks and the R accumulators are
// double-buffered, so generating tile/sample n+1, the tensor-core work of unit u+1 and the
// epilogue of unit u overlap.  All hand-offs are mbarriers (full/empty pairs).
//
//   LOCAL  CTA = (MC sample s, gene range, 256-cell block): units = (tile, half); z static in smem;
//          dz0[half] += Q W^T accumulates in TMEM over the tiles of the range.
//   GLOBAL CTA = (gene tile, chunk of MC samples): units = (s, 128-cell block); dW[s&1] += z0^T Q over the
//          blocks of s; epilogue 2 folds W.dW into per-thread d mu / d rho accumulators over the chunk.
#pragma once

constexpr int PIPE_NT = 576;           // 18 warps
constexpr int E_THREADS = 512, E_WARPS = 16, MMA_WARP = 16, COPY_WARP = 17;

struct PipeMisc {
  uint64_t w_full[2], w_empty[2], z_full[2], z_empty[2], r_full[2], r_empty[2], q_full, q_empty, dw_full[2],
      dw_empty[2], dz_full;
  uint32_t tmem_base, pad;
  double red[32];
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
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- producers -------------------------------------------------------------------------------------
// ---- the kernel ----------------------------------------------------------------------------------------
template <bool GLOBAL, bool WANT_LL, bool MAT = false>   // MAT: matrix-valued W_0 prior (iscDEF), GLOBAL only
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
      mbar_init(&mi->w_empty[i], GLOBAL ? E_THREADS : 1);
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
  if (warp == MMA_WARP) {
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

  if (warp == COPY_WARP) {
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
        if (a.l2_prefetch && oi + 2 < n_outer) {
          // opt-in (SCDEF_TC_L2PF=1, not measured yet): pull the W image this stage will receive NEXT (two outer steps
          // ahead) into L2 now; the stage is only refilled once epilogue 2 has released it, which leaves about one unit
          // of time for the copy (the MMA thread waits ~3k cycles per sample on w_full in the GLOBAL kernel)
          const uint8_t* nxt = GLOBAL ? ws.wimg + ((long long)(s + 2) * gm.n_tiles + tile) * L.w_unit
                                      : ws.wimg + ((long long)s * gm.n_tiles + tile + 2) * L.w_unit;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nxt), "r"(w_bytes) : "memory");
        }
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
  } else if (warp == MMA_WARP) {
    // =========================================== MMA ISSUER ===========================================
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc(128, TG, 0, 1);
      const uint32_t idesc2 = GLOBAL ? umma_idesc(128, TG, 1, 1) : umma_idesc(128, KP, 0, 0);
      // descriptors differ only in the 14-bit start-address field (>>4): build them once, then add constants
      const uint64_t dz_k[2] = {umma_desc(sbase + L.z[0], 128u, L.z_rs), umma_desc(sbase + L.z[1], 128u, L.z_rs)};      // z, K-major
      const uint64_t dz_m[2] = {umma_desc(sbase + L.z[0], L.z_rs, 128u), umma_desc(sbase + L.z[1], L.z_rs, 128u)};      // z, MN-major
      const uint64_t dw_m[2] = {umma_desc(sbase + L.w[0], W_RS, 128u), umma_desc(sbase + L.w[1], W_RS, 128u)};          // W, MN-major
      const uint64_t dw_k[2] = {umma_desc(sbase + L.w[0], 128u, W_RS), umma_desc(sbase + L.w[1], 128u, W_RS)};          // W, K-major
      const uint64_t dq_k = umma_desc(sbase + L.q, 128u, Q_RS), dq_m = umma_desc(sbase + L.q, Q_RS, 128u);              // Q
      const uint64_t zpl = L.z_plane >> 4, wpl = L.w_plane >> 4, qpl = L.q_plane >> 4;
      const int nk1 = KP / 16;
      auto issue_mma1 = [&](int u) {
        const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
        const int st = oi & 1, rs = u & 1, zb = GLOBAL ? (u & 1) : cb;
        if (cb == 0) TWAIT(0, &mi->w_full[st], (oi >> 1) & 1);
        if (GLOBAL) TWAIT(1, &mi->z_full[zb], (u >> 1) & 1);
        TWAIT(2, &mi->r_empty[rs], ((u >> 1) & 1) ^ 1);
        tc_fence_after();
        uint64_t da = dz_k[zb], db = dw_m[st];
        const uint32_t d = tmem + rs * TG;
        const bool full = a.dbg_nsplit != 1;
        if (full) { umma_bf16(d, da + zpl, db, idesc1, 0u); umma_bf16(d, da, db + wpl, idesc1, 1u); }
        umma_bf16(d, da, db, idesc1, full ? 1u : 0u);
        for (int ks = 1; ks < nk1; ++ks) {
          da += 16;              // 256 B along K
          db += 2 * (W_RS >> 4); // two k0 row groups
          if (full) { umma_bf16(d, da + zpl, db, idesc1, 1u); umma_bf16(d, da, db + wpl, idesc1, 1u); }
          umma_bf16(d, da, db, idesc1, 1u);
        }
        umma_commit(&mi->r_full[rs]);
      };
      if (!GLOBAL) mbar_wait(&mi->z_full[0], 0);
      // Units are issued in order; whenever Q of unit u is not ready yet, the rate GEMM of unit u+1 is slipped in
      // front of the second GEMM of unit u (R is double-buffered), otherwise the second GEMM goes first so that the
      // single Q buffer is released as early as possible.
      int next1 = 0;  // next unit whose MMA1 has not been issued
      if (U > 0) { issue_mma1(0); next1 = 1; }
      // opt-in (SCDEF_TC_POLL=1, not measured yet): slip the next rate GEMM in only when its operands have landed and its
      // R buffer is free, and keep polling for that while waiting for Q — issue_mma1 blocks on w_full / z_full / r_empty,
      // and in the GLOBAL kernel the MMA thread was seen waiting ~30 % of its time there while Q tiles were ready
      auto ready1 = [&](int un) {
        const int oi1 = un / units_per_outer, cb1 = un - oi1 * units_per_outer;
        if (cb1 == 0 && !mbar_test(&mi->w_full[oi1 & 1], (oi1 >> 1) & 1)) return false;
        if (GLOBAL && !mbar_test(&mi->z_full[un & 1], (un >> 1) & 1)) return false;
        return mbar_test(&mi->r_empty[un & 1], ((un >> 1) & 1) ^ 1);
      };
      for (int u = 0; u < U; ++u) {
        if (a.mma_poll) {
          while (!mbar_test(&mi->q_full, u & 1)) {
            if (next1 == u + 1 && next1 < U && ready1(next1)) { issue_mma1(next1); ++next1; }
          }
        } else
        if (next1 == u + 1 && next1 < U && !mbar_test(&mi->q_full, u & 1)) { issue_mma1(next1); ++next1; }
        const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
        const int st = oi & 1;
        TWAIT(3, &mi->q_full, u & 1);
        if (GLOBAL && cb == 0) TWAIT(4, &mi->dw_empty[st], ((oi >> 1) & 1) ^ 1);
        tc_fence_after();
        if (GLOBAL) {
          // dW[st] (+)= z0^T Q : A = z (MN-major, M = k0), B = Q (MN-major, N = genes), K = cells of the unit
          const int rows = min(UC, a.B - cb * UC), nks = (rows + 15) / 16;
          const int zb = u & 1;
          uint64_t da = dz_m[zb], db = dq_m;
          const uint32_t d = tmem_acc2 + st * TG;
          const bool full = a.dbg_nsplit != 1;
          if (full) { umma_bf16(d, da + zpl, db, idesc2, cb > 0 ? 1u : 0u); umma_bf16(d, da, db + qpl, idesc2, 1u); }
          umma_bf16(d, da, db, idesc2, (full || cb > 0) ? 1u : 0u);
          for (int ks = 1; ks < nks; ++ks) {
            da += 2 * (L.z_rs >> 4);
            db += 2 * (Q_RS >> 4);
            if (full) { umma_bf16(d, da + zpl, db, idesc2, 1u); umma_bf16(d, da, db + qpl, idesc2, 1u); }
            umma_bf16(d, da, db, idesc2, 1u);
          }
          umma_commit(&mi->q_empty);
          umma_commit(&mi->z_empty[zb]);
          if (cb == units_per_outer - 1) umma_commit(&mi->dw_full[st]);
        } else {
          // dz[cb] (+)= Q W^T : A = Q (K-major), B = W (K-major, N = k0), K = genes of the tile
          uint64_t da = dq_k, db = dw_k[st];
          const uint32_t d = tmem_acc2 + cb * KP;
          const bool full = a.dbg_nsplit != 1;
          if (full) { umma_bf16(d, da + qpl, db, idesc2, oi > 0 ? 1u : 0u); umma_bf16(d, da, db + wpl, idesc2, 1u); }
          umma_bf16(d, da, db, idesc2, (full || oi > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < TG / 16; ++ks) {
            da += 16;
            db += 16;
            if (full) { umma_bf16(d, da + qpl, db, idesc2, 1u); umma_bf16(d, da, db + wpl, idesc2, 1u); }
            umma_bf16(d, da, db, idesc2, 1u);
          }
          umma_commit(&mi->q_empty);
          if (cb == units_per_outer - 1) umma_commit(&mi->w_empty[st]);
        }
        if (next1 == u + 1 && next1 < U) { issue_mma1(next1); ++next1; }
      }
      if (!GLOBAL) umma_commit(&mi->dz_full);
    }
    __syncwarp();
  } else {
    // =========================================== EPILOGUE ===========================================
    const int e = warp;                   // 0..15
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int r = (int)lane_base + lane;  // row of the unit (cell) / k0 row in epilogue 2
    const int cq = e >> 2;                // 16-gene quarter of the tile
    float ll = 0.f;
    float esum0 = 0.f, esum1 = 0.f;
    const uint32_t tmem_x = tmem + 4 * TG;  // GLOBAL: the X tile of the first two 128-cell units (2 x TG columns)
    if (GLOBAL) {
      for (int cb = 0; cb < min(units_per_outer, 2); ++cb) {
        const int j = cb * UC + r;
        const int gq = tile_fixed * TG + cq * 16;
        float xv[16];
        (void)j; (void)gq;
        const float* xi = ws.ximg + ((((long long)tile_fixed * units_total_B + cb) * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 t4 = *reinterpret_cast<const float4*>(xi + v * 128);
          xv[v * 4] = t4.x; xv[v * 4 + 1] = t4.y; xv[v * 4 + 2] = t4.z; xv[v * 4 + 3] = t4.w;
        }
        tmem_st16(tmem_x + (lane_base << 16) + (uint32_t)(cb * TG + cq * 16), xv);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    float acc_mu[16], acc_lw[16];
    if (GLOBAL) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { acc_mu[i] = 0.f; acc_lw[i] = 0.f; }
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
    auto epilogue2 = [&](int oi2) {
      // ---- fold dW^s (and the W_0 prior gradient) into d mu, sum t log W ----
      const int st2 = oi2 & 1;
      const int s2 = outer_begin + oi2;
        TWAIT(2, &mi->dw_full[st2], (oi2 >> 1) & 1);
        tc_fence_after();
        float dW[16];
        tmem_ld16(tmem_acc2 + (lane_base << 16) + (uint32_t)(st2 * TG + cq * 16), dW);
        tc_fence_before();
        mbar_arrive(&mi->dw_empty[st2]);
        if (r < a.K0 && !a.w0_stopped) {
          float A = a.P_scalar, Bm = a.R_scalar * (float)a.K0;
          float inv_tau = 1.f, inv_tau_om = 1.f;
          if (a.use_brd) {
            const float tau = a.smp_tau[(long long)s2 * a.K0 + r], om = a.smp_om[(long long)s2 * a.K0 + r];
            inv_tau = 1.f / tau;
            inv_tau_om = 1.f / (tau * om);
            A = a.P_scalar * inv_tau;
            Bm = a.R_scalar * (float)a.K0 * inv_tau_om;
          }
          // iscDEF: matrix-valued Gamma shape / rate of W_0, one (k, g) entry per element of this thread's 16 genes
          const int gbase = tile_fixed * TG + cq * 16;
          const float* Prow = (MAT && a.P) ? a.P + (long long)r * a.G + gbase : nullptr;
          const float* Rrow = (MAT && a.R) ? a.R + (long long)r * a.G + gbase : nullptr;
          const uint32_t woff = L.w[st2] + (uint32_t)(r >> 3) * W_RS + (uint32_t)(r & 7) * 16u;
#pragma unroll
          for (int c8 = 0; c8 < 2; ++c8) {
            const uint4 hv = *reinterpret_cast<const uint4*>(sm + woff + (uint32_t)(cq * 2 + c8) * 128u);
            const uint4 lv = *reinterpret_cast<const uint4*>(sm + woff + L.w_plane + (uint32_t)(cq * 2 + c8) * 128u);
            const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
            for (int p = 0; p < 4; ++p) {
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                const int i = c8 * 8 + p * 2 + q2;
                const float W = q2 ? (bf16hi(hw[p]) + bf16hi(lw[p])) : (bf16lo(hw[p]) + bf16lo(lw[p]));
                if (MAT && gbase + i < a.G) {
                  A = (Prow ? __ldg(Prow + i) : a.P_scalar) * inv_tau;
                  Bm = (Rrow ? __ldg(Rrow + i) : a.R_scalar) * (float)a.K0 * inv_tau_om;
                }
                if (W > V_LO_M && W < V_HI_M) {
                  const float t = a.scale * dW[i] * W + a.global_weight * (A - 1.f - Bm * W);
                  acc_mu[i] += t;
                  acc_lw[i] = fmaf(t, __logf(W), acc_lw[i]);
                }
              }
            }
          }
        }
        mbar_arrive(&mi->w_empty[st2]);
    };
    for (int u = 0; u < U; ++u) {
#ifdef SCDEF_TC_TIMING
      long long tq0 = clock64();
#endif
      const int oi = u / units_per_outer, cb = u - oi * units_per_outer;
      const int st = oi & 1, rs = u & 1;
      const int s = GLOBAL ? outer_begin + oi : s_fixed;
      const int g0 = (GLOBAL ? tile_fixed : outer_begin + oi) * TG + cq * 16;
      const int j = (GLOBAL ? cb * UC : j_block0 + cb * UC) + r;
      const Ctx cx = nxt;
      const bool valid = cx.valid;
      const float c = cx.c;
      const bool x_in_tmem = GLOBAL && cb < 2;
      float4 x4[4];
      if (!x_in_tmem) {
        const int tile_u = GLOBAL ? tile_fixed : outer_begin + oi;
        const int unit_u = GLOBAL ? cb : j_block0 / UC + cb;
        const float* xi = ws.ximg + ((((long long)tile_u * units_total_B + unit_u) * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4;
#pragma unroll
        for (int v = 0; v < 4; ++v) x4[v] = *reinterpret_cast<const float4*>(xi + v * 128);
      }
      nxt = fetch_ctx(u + 1);
#ifdef SCDEF_TC_TIMING
      { long long tq1 = clock64(); tacc[6] += tq1 - tq0; tq0 = tq1; }
#endif
      if (cb == 0) mbar_wait(&mi->w_full[st], (oi >> 1) & 1);  // the gene-scale tile rides with the W image
#ifdef SCDEF_TC_TIMING
      tacc[7] += clock64() - tq0;
#endif
      TWAIT(0, &mi->r_full[rs], (u >> 1) & 1);
      tc_fence_after();
      float R[16], xs16[16], q[16];
#ifdef SCDEF_TC_TIMING
      long long tp0 = clock64();
#endif
      tmem_ld16(tmem + (lane_base << 16) + (uint32_t)(rs * TG + cq * 16), R);
      if (x_in_tmem) {
        tmem_ld16(tmem_x + (lane_base << 16) + (uint32_t)(cb * TG + cq * 16), xs16);
      } else {
#pragma unroll
        for (int v = 0; v < 4; ++v) { xs16[v * 4] = x4[v].x; xs16[v * 4 + 1] = x4[v].y; xs16[v * 4 + 2] = x4[v].z; xs16[v * 4 + 3] = x4[v].w; }
      }
      const float* srow_sm = reinterpret_cast<const float*>(sm + L.w[st] + 2 * L.w_plane) + cx.b * TG + cq * 16;
      float es = 0.f;
#ifdef SCDEF_TC_TIMING
      long long tp1 = clock64(); tacc[3] += tp1 - tp0;
#endif
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
            const float t = xv * __logf(lam);  // lam = 0 only where x = 0: the select discards -inf * 0
            ll += (xv > 0.f ? t : 0.f) - lam;
          }
          es += E;
        }
      }
      if (g_dbg_R && valid) {
        for (int i = 0; i < 16; ++i)
          if (g0 + i < a.G) g_dbg_R[((long long)s * a.B + j) * a.G + g0 + i] = R[i];
      }
#ifdef SCDEF_TC_TIMING
      { long long tp2 = clock64(); tacc[4] += tp2 - tp1; }
#endif
      TWAIT(1, &mi->q_empty, (u & 1) ^ 1);
#ifdef SCDEF_TC_TIMING
      long long tp3 = clock64();
#endif
      {
        uint4 hi, lo;
        const uint32_t off = L.q + (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u + (uint32_t)(cq * 2) * 128u;
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
#ifdef SCDEF_TC_TIMING
      tacc[5] += clock64() - tp3;
#endif
      if (cb & 1) esum1 += es; else esum0 += es;

      // epilogue 2 of sample oi-1 runs one unit late: its dW MMA has certainly drained by then (dW is double-buffered),
      // so the epilogue warps never idle on dw_full
      if (GLOBAL && cb == 0 && oi > 0) epilogue2(oi - 1);
    }
    if (GLOBAL && n_outer > 0) epilogue2(n_outer - 1);

    if (GLOBAL) {
      const int chunk = blockIdx.x / gm.n_tiles;
      const int gq = tile_fixed * TG + cq * 16;
      if (r < a.K0 && !a.w0_stopped) {
        float* pm = ws.part + (((long long)chunk * 2 + 0) * a.K0 + r) * a.G + gq;
        float* pr = ws.part + (((long long)chunk * 2 + 1) * a.K0 + r) * a.G + gq;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (gq + i < a.G) {
            *reinterpret_cast<float4*>(pm + i) = make_float4(acc_mu[i], acc_mu[i + 1], acc_mu[i + 2], acc_mu[i + 3]);
            *reinterpret_cast<float4*>(pr + i) = make_float4(acc_lw[i], acc_lw[i + 1], acc_lw[i + 2], acc_lw[i + 3]);
          }
        }
      }
    } else {
      // ---- drain dz0[half] (this thread: cell row r, every 4th 16-column chunk of k0) ----
      mbar_wait(&mi->dz_full, 0);
      tc_fence_after();
      const int range = blockIdx.x % gm.NG;
      for (int h = 0; h < units_per_outer; ++h) {
        const int j = j_block0 + h * UC + r;
        float* dst = ws.part + (((long long)s_fixed * gm.NG + range) * a.B + j) * a.K0;
        for (int cc = cq; cc < KP / 16; cc += 4) {
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
    // loss: reduce over the epilogue threads
    double v = warp_sum_d((double)ll);
    if (lane == 0) mi->red[e] = v;
    named_bar_sync(2, E_THREADS);
    if (e == 0 && lane == 0) {
      double tot = 0.0;
      for (int i = 0; i < E_WARPS; ++i) tot += mi->red[i];
      atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
    }
  }

#ifdef SCDEF_TC_TIMING
  if (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && lane == 0 && (warp == 1 || warp == MMA_WARP || warp == 0 || warp == COPY_WARP)) {
    const int role = warp == 1 ? 0 : (warp == MMA_WARP ? 1 : (warp == 0 ? 2 : 3));
    for (int i = 0; i < 15; ++i) g_tc_timing[GLOBAL ? 1 : 0][role][i] = (unsigned long long)tacc[i];
    g_tc_timing[GLOBAL ? 1 : 0][role][15] = (unsigned long long)(clock64() - t_start);
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
