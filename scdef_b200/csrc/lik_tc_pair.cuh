// CTA-PAIR version of the LOCAL likelihood kernel (included by lik_tc.cu inside its anonymous namespace, after
// lik_tc_pipe.cuh; uses its PTX wrappers, layouts and the pre-kernels' images).
//
// STATUS: compiled for sm_100a and SASS-checked, NOT YET RUN ON A GPU (written when the round's GPU minutes were
// spent).  Opt-in: SCDEF_TC_PAIR=1.  Every barrier wait in this file is bounded (trap after ~1 s) so that a protocol
// error ends in a CUDA error, not a hung device.  Bring-up: tests/tc_bringup.py with SCDEF_TC_PAIR=1.
//
// Why (profiles/r01_mma_probe.txt): a cta_group::1 tcgen05.mma occupies the tensor pipe for ~86 cycles whatever
// N <= 128 is; the pair's 256 x 128 x 16 MMA costs 64.  k_lik_pipe<LOCAL> issues 33 MMAs per 128-cell x 64-gene
// unit (2.8k cycles); here the same work is 45 pair MMAs per FOUR units (0.72k cycles per unit).
//
// Work of one pair = (MC sample s, range of 128-gene pair-tiles, 256-cell block); CTA `rank` of the pair owns the
// cells [j0 + 128 rank, j0 + 128 rank + 128).  Per pair-tile pt (W_0 tile images 2pt and 2pt+1 of k_tc_wgen):
//   MMA1  R[256 cells, 128 genes] = z0 W           M = 256 (128 rows per CTA), N = 128, K = KP; 3 split products
//         A = the CTA's own z unit (K-major); B = N split over the pair: CTA `rank` supplies genes
//         [64 rank, 64 rank + 64) = ALL k0 rows of image 2pt+rank (MN-major) -> region WA of the stage
//   epilogue (both CTAs, 16 warps each), two sub-units h = 0, 1 of 64 genes: R -> Q = (x - Lambda) / R, split to
//         bf16 hi/lo into the CTA's single Q buffer (128 cells x 64 genes), exactly as k_lik_pipe does
//   MMA2  dz0[256 cells, KP] += Q_h W_h^T          M = 256, N = KP, K = 64 genes (4 k-steps); per sub-unit
//         A = the CTA's own Q (K-major); B = N split over the pair: CTA `rank` supplies the k0 rows
//         [KP/2 rank, KP/2 rank + KP/2) of image 2pt+h (K-major) -> region B2[h] of the stage
//   So a stage holds: [WA hi | WA lo | gene-scale tile of image 2pt | of image 2pt+1 | B2[0] hi | B2[0] lo |
//   B2[1] hi | B2[1] lo]; the same smem offsets in both CTAs (one descriptor serves both: the issuing thread's
//   descriptors are CTA-relative).  7 bulk copies per stage and CTA, all contiguous pieces of the tile images.
//
// Barrier protocol.  Only the leader (rank 0) issues MMAs.  tcgen05.commit is multicast to both CTAs, so
// "tensor core is done with ..." barriers (r_full, q_empty, w_empty, dz_full) are waited locally in each CTA.
// "... is ready for the tensor core" barriers live in the LEADER's shared memory and are arrived on remotely
// (mapa + mbarrier.arrive on the shared::cluster address) by one elected lane per epilogue warp of both CTAs (r_empty, q_full:
// 32 arrivals) or by the peer's relay thread (the peer's otherwise idle MMA warp forwards "my TMA landed":
// pz_full, pw_full).
#pragma once

constexpr int PAIR_TMEM_DZ = 256;            // TMEM columns: R[2] at 0 / 128, dz at 256
constexpr long long PAIR_DEADLINE = 2000000000ll;  // cycles (~1 s) before a stuck wait traps

struct PairMisc {
  uint64_t z_full, pz_full, w_full[2], pw_full[2], w_empty[2], r_full[2], r_empty[2], q_full, q_empty, dz_full;
  uint32_t tmem_base, pad;
  double red[32];
};

struct PairSmem {
  uint32_t z, w[2], q, misc, total;
  uint32_t z_rs, z_plane, w_plane, q_plane;
  uint32_t sg_off, b2_off, b2_piece, w_stage;   // offsets inside a W stage
};
__host__ __device__ inline PairSmem plan_pair_smem(int KP) {
  PairSmem s;
  s.z_rs = (uint32_t)(KP / 8) * 128u;
  s.z_plane = 16u * s.z_rs;                  // 128 cells
  s.w_plane = (uint32_t)(KP / 8) * W_RS;     // KP rows x 64 genes
  s.q_plane = 16u * Q_RS;                    // 128 cells x 64 genes
  s.b2_piece = (uint32_t)(KP / 16) * W_RS;   // KP/2 rows x 64 genes
  s.sg_off = 2u * s.w_plane;
  s.b2_off = s.sg_off + 2u * (uint32_t)(MAX_NB * TG * 4);
  s.w_stage = s.b2_off + 4u * s.b2_piece;
  uint32_t o = 0;
  s.z = o; o += 2u * s.z_plane;
  s.w[0] = o; o += s.w_stage;
  s.w[1] = o; o += s.w_stage;
  s.q = o; o += 2u * s.q_plane;
  s.misc = o;
  s.total = o + (uint32_t)((sizeof(PairMisc) + 1023) / 1024 * 1024);
  return s;
}

// ---- cluster-scope PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `bar` in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// Scope of the cross-CTA hand-offs.  Default: the PTX defaults (.release / .acquire at .cta scope), which is what the
// CUTLASS 2-SM kernels use for the same hand-offs (ClusterBarrier::arrive(cta_id) / ::wait) -- the data they guard is
// ordered by its own fences (fence.proxy.async for the Q tile, tcgen05.fence for tensor memory).  -DSCDEF_PAIR_STRICT
// makes them cluster-scoped; ptxas then brackets every such wait with CCTL.IVALL and every arrive with
// MEMBAR.ALL.GPU (seen in the SASS), so it is a debugging aid, not the production setting.
#ifdef SCDEF_PAIR_STRICT
#define PAIR_REL ".release.cluster"
#define PAIR_ACQ ".acquire.cluster"
#else
#define PAIR_REL ""
#define PAIR_ACQ ""
#endif
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive" PAIR_REL ".shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// try_wait with a short suspend hint; REMOTE: the barrier is arrived on from the peer CTA
template <bool REMOTE>
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  if (REMOTE)
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity" PAIR_ACQ ".shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x10000u)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x10000u)
        : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_remote(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity" PAIR_ACQ ".shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// {code of the wait that timed out, blockIdx.x, warp, parity}: points into mapped host memory (lik_tc_launch sets it
// before the first pair launch), so the record survives the trap that follows it
__device__ volatile int* g_pair_stuck = nullptr;
__device__ __noinline__ void pair_stuck(int code, int x, int y, int z) {
  volatile int* p = g_pair_stuck;
  if (p) { p[0] = code; p[1] = x; p[2] = y; p[3] = z; }
  __threadfence_system();
  __trap();
}
// bounded waits: a protocol error ends in a trap (and a record of which wait it was), not in a hung device
template <bool REMOTE>
__device__ __forceinline__ void pair_wait_t(uint64_t* bar, uint32_t parity, int code) {
  if (mbar_try<REMOTE>(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try<REMOTE>(bar, parity)) {
    if (clock64() - t0 > PAIR_DEADLINE) pair_stuck(code, (int)blockIdx.x, (int)(threadIdx.x >> 5), (int)parity);
  }
}
__device__ __forceinline__ void pair_wait(uint64_t* bar, uint32_t parity, int code) { pair_wait_t<false>(bar, parity, code); }
__device__ __forceinline__ void pair_wait_remote(uint64_t* bar, uint32_t parity, int code) { pair_wait_t<true>(bar, parity, code); }
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// completion of all MMAs issued so far by this thread -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// generic-proxy stores of the Q tile -> visible to the tensor core's (async proxy) operand reads.  The unqualified
// fence costs MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + FENCE.VIEW.ASYNC in SASS; the shared::cta form is the last one only.
__device__ __forceinline__ void fence_async_q() {
#ifdef SCDEF_PAIR_STRICT
  asm volatile("fence.proxy.async;" ::: "memory");
#else
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

#ifdef SCDEF_TC_TIMING
// [rank][role: COPY, MMA (leader) / RELAY (peer), epilogue warp 0, epilogue warp 15][slot]; slot 15 = total cycles
__device__ unsigned long long g_pair_timing[2][4][16];
#define PT_DECL long long tacc[16]; for (int _i = 0; _i < 16; ++_i) tacc[_i] = 0; const long long t_start = clock64(); long long t_mark = t_start
#define PT_WAIT(slot, call) do { const long long _t0 = clock64(); call; tacc[slot] += clock64() - _t0; } while (0)
#define PT_LAP(slot) do { const long long _t1 = clock64(); tacc[slot] += _t1 - t_mark; t_mark = _t1; } while (0)
#define PT_MARK() do { t_mark = clock64(); } while (0)
#else
#define PT_DECL do { } while (0)
#define PT_WAIT(slot, call) call
#define PT_LAP(slot) do { } while (0)
#define PT_MARK() do { } while (0)
#endif

// ---- the kernel ------------------------------------------------------------------------------------------
// grid = (2 * S * NGp, B / 256), cluster (2,1,1): blockIdx.x >> 1 = (s, range); requires B % 256 == 0 and an even
// number of tile images per MC sample (make_geom pads n_tiles in pair mode; the phantom image is all zeros:
// R = 0, x = s = 0 -> Q = 0).
template <bool WANT_LL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PIPE_NT, 1)
    k_lik_pair_local(const LikTcArgs a, const Geom gm, const TcScratch ws, double* loss_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const PairSmem L = plan_pair_smem(KP);
  PairMisc* mi = reinterpret_cast<PairMisc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  PT_DECL;

  // ---- work of this pair ----
  const int item = blockIdx.x >> 1;
  const int s_fixed = item / gm.NGp, range = item % gm.NGp;
  const int pt_begin = range * gm.ptiles_per_range;
  const int pt_end = min(gm.n_tiles / 2, pt_begin + gm.ptiles_per_range);
  const int n_outer = pt_end - pt_begin;          // pair-tiles of this range
  const int U = 2 * n_outer;                      // sub-units (pair-tile, gene half)
  const int j_block0 = blockIdx.y * CB;
  const int units_total_B = (a.B + UC - 1) / UC;
  const int unit_u = j_block0 / UC + (int)rank;   // this CTA's 128-cell unit in the z / x images

  // ---- prologue ----
  if (threadIdx.x == 0) {
    mbar_init(&mi->z_full, 1);
    mbar_init(&mi->pz_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mi->w_full[i], 1);
      mbar_init(&mi->pw_full[i], 1);
      mbar_init(&mi->w_empty[i], 1);
      mbar_init(&mi->r_full[i], 1);
      mbar_init(&mi->r_empty[i], 2 * 2 * E_WARPS);   // (2 sub-units) x (16 warps) x (2 CTAs) elected arrivals
    }
    mbar_init(&mi->q_full, 2 * E_WARPS);             // 16 warps x 2 CTAs
    mbar_init(&mi->q_empty, 1);
    mbar_init(&mi->dz_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                                // the peer's barriers exist before anything arrives on them
  if (warp == MMA_WARP) {                            // same warp in both CTAs (Allocator2Sm contract)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&mi->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = mi->tmem_base;
  const uint32_t tmem_dz = tmem + PAIR_TMEM_DZ;
  const uint32_t z_bytes = 2u * L.z_plane;
  const uint32_t sg_bytes = (uint32_t)(a.nb * TG * 4);
  const uint32_t w_bytes = 2u * L.w_plane + 2u * sg_bytes + 4u * L.b2_piece;
  const PipeSmem IL = plan_pipe_smem(KP);            // layout of one tile image in global memory (k_tc_wgen)

  if (warp == COPY_WARP) {
    // =========================================== COPY (TMA), each CTA for itself ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(&mi->z_full, z_bytes);
      bulk_g2s(sbase + L.z, ws.zimg + ((long long)s_fixed * units_total_B + unit_u) * z_bytes, z_bytes, &mi->z_full);
      for (int oi = 0; oi < n_outer; ++oi) {
        const int st = oi & 1;
        const int pt = pt_begin + oi;
        const uint8_t* img0 = ws.wimg + ((long long)s_fixed * gm.n_tiles + 2 * pt) * IL.w_unit;
        const uint8_t* img1 = img0 + IL.w_unit;
        const uint8_t* own = rank ? img1 : img0;
        PT_WAIT(0, pair_wait(&mi->w_empty[st], ((oi >> 1) & 1) ^ 1, 1));
        mbar_arrive_expect_tx(&mi->w_full[st], w_bytes);
        const uint32_t dst = sbase + L.w[st];
        bulk_g2s(dst, own, 2u * L.w_plane, &mi->w_full[st]);                                        // WA: hi | lo
        bulk_g2s(dst + L.sg_off, img0 + 2u * IL.w_plane, sg_bytes, &mi->w_full[st]);                // gene scale, half 0
        bulk_g2s(dst + L.sg_off + (uint32_t)(MAX_NB * TG * 4), img1 + 2u * IL.w_plane, sg_bytes, &mi->w_full[st]);
        const uint32_t row_off = rank * L.b2_piece;                                                 // k0 rows [KP/2 rank, +KP/2)
        bulk_g2s(dst + L.b2_off + 0u * L.b2_piece, img0 + row_off, L.b2_piece, &mi->w_full[st]);             // B2[0] hi
        bulk_g2s(dst + L.b2_off + 1u * L.b2_piece, img0 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]); // B2[0] lo
        bulk_g2s(dst + L.b2_off + 2u * L.b2_piece, img1 + row_off, L.b2_piece, &mi->w_full[st]);             // B2[1] hi
        bulk_g2s(dst + L.b2_off + 3u * L.b2_piece, img1 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]); // B2[1] lo
      }
      // the multicast commits of the last two pair-tiles are waited for by nobody else: see them land before this
      // CTA can exit (a late arrival would hit the shared memory of whatever CTA runs here next)
      for (int st = 0; st < 2; ++st) {
        const int uses = (n_outer + 1 - st) / 2;   // pair-tiles that went through stage st
        if (uses > 0) pair_wait(&mi->w_empty[st], (uint32_t)(uses - 1) & 1u, 15);
      }
      if (U > 0) pair_wait(&mi->q_empty, (uint32_t)(U - 1) & 1u, 16);
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (lane == 0 && !leader) {
      // =========================================== RELAY (peer) ==========================================
      // forwards "my TMA landed" to the leader, whose MMAs read this CTA's shared memory as well
      const uint32_t r_pz = mapa_rank(&mi->pz_full, 0), r_pw0 = mapa_rank(&mi->pw_full[0], 0), r_pw1 = mapa_rank(&mi->pw_full[1], 0);
      pair_wait(&mi->z_full, 0, 2);
      mbar_arrive_remote(r_pz);
      for (int oi = 0; oi < n_outer; ++oi) {
        PT_WAIT(0, pair_wait(&mi->w_full[oi & 1], (oi >> 1) & 1, 3));
        mbar_arrive_remote((oi & 1) ? r_pw1 : r_pw0);
      }
    } else if (lane == 0) {
      // =========================================== MMA ISSUER (leader) ===================================
      const uint32_t idesc1 = umma_idesc(256, 2 * TG, 0, 1);   // R : A K-major, B MN-major
      const uint32_t idesc2 = umma_idesc(256, KP, 0, 0);       // dz: A K-major, B K-major
      const uint64_t dz_k = umma_desc(sbase + L.z, 128u, L.z_rs);
      const uint64_t dwa_m[2] = {umma_desc(sbase + L.w[0], W_RS, 128u), umma_desc(sbase + L.w[1], W_RS, 128u)};
      const uint64_t db2_k[2] = {umma_desc(sbase + L.w[0] + L.b2_off, 128u, W_RS), umma_desc(sbase + L.w[1] + L.b2_off, 128u, W_RS)};
      const uint64_t dq_k = umma_desc(sbase + L.q, 128u, Q_RS);
      const uint64_t zpl = L.z_plane >> 4, wpl = L.w_plane >> 4, qpl = L.q_plane >> 4, b2pl = L.b2_piece >> 4;
      const int nk1 = KP / 16;
      auto ready1 = [&](int oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        return mbar_test(&mi->w_full[st], par) && mbar_test_remote(&mi->pw_full[st], par) &&
               mbar_test_remote(&mi->r_empty[st], par ^ 1u);
      };
      auto issue_mma1 = [&](int oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        PT_WAIT(0, pair_wait(&mi->w_full[st], par, 4));
        PT_WAIT(1, pair_wait_remote(&mi->pw_full[st], par, 5));
        PT_WAIT(2, pair_wait_remote(&mi->r_empty[st], par ^ 1u, 6));
        tc_fence_after();
        PT_MARK();
        uint64_t da = dz_k, db = dwa_m[st];
        const uint32_t d = tmem + (uint32_t)st * (2 * TG);
        umma2_bf16(d, da + zpl, db, idesc1, 0u);
        umma2_bf16(d, da, db + wpl, idesc1, 1u);
        umma2_bf16(d, da, db, idesc1, 1u);
        for (int ks = 1; ks < nk1; ++ks) {
          da += 16;                // 256 B along K (two k0 column groups of the z unit)
          db += 2 * (W_RS >> 4);   // two k0 row groups of the W image
          umma2_bf16(d, da + zpl, db, idesc1, 1u);
          umma2_bf16(d, da, db + wpl, idesc1, 1u);
          umma2_bf16(d, da, db, idesc1, 1u);
        }
        umma2_commit(&mi->r_full[st]);
        PT_LAP(4);
      };
      PT_WAIT(6, pair_wait(&mi->z_full, 0, 7));
      PT_WAIT(6, pair_wait_remote(&mi->pz_full, 0, 8));
      int next1 = 0;
      for (int u = 0; u < U; ++u) {
        const int oi = u >> 1, h = u & 1, st = oi & 1;
        // the rate GEMM of the next pair-tile is slipped in as soon as its operands and its R buffer are ready; the
        // one of THIS pair-tile must have been issued before its Q can ever arrive
        long long t_poll = 0;
#ifdef SCDEF_TC_TIMING
        const long long t_loop0 = clock64(), t_in1 = tacc[0] + tacc[1] + tacc[2] + tacc[4];
#endif
        for (;;) {
          if (next1 < n_outer && next1 <= oi + 1 && (next1 == oi || ready1(next1))) { issue_mma1(next1); ++next1; continue; }
          if (mbar_test_remote(&mi->q_full, (uint32_t)u & 1u)) break;
          if (next1 > oi + 1 || next1 >= n_outer) { pair_wait_remote(&mi->q_full, (uint32_t)u & 1u, 9); break; }
          if (t_poll == 0) t_poll = clock64();
          else if (clock64() - t_poll > PAIR_DEADLINE) pair_stuck(14, (int)blockIdx.x, u, next1);
        }
#ifdef SCDEF_TC_TIMING
        tacc[3] += (clock64() - t_loop0) - (tacc[0] + tacc[1] + tacc[2] + tacc[4] - t_in1);   // waiting for Q, net of slipped-in MMA1 work
#endif
        tc_fence_after();
        PT_MARK();
        // dz (+)= Q_h W_h^T : A = Q (K-major), B = B2[h] (K-major, N = k0), K = the 64 genes of the half
        uint64_t da = dq_k, db = db2_k[st] + (uint64_t)(2 * h) * b2pl;
        umma2_bf16(tmem_dz, da + qpl, db, idesc2, u > 0 ? 1u : 0u);
        umma2_bf16(tmem_dz, da, db + b2pl, idesc2, 1u);
        umma2_bf16(tmem_dz, da, db, idesc2, 1u);
#pragma unroll
        for (int ks = 1; ks < TG / 16; ++ks) {
          da += 16;
          db += 16;
          umma2_bf16(tmem_dz, da + qpl, db, idesc2, 1u);
          umma2_bf16(tmem_dz, da, db + b2pl, idesc2, 1u);
          umma2_bf16(tmem_dz, da, db, idesc2, 1u);
        }
        umma2_commit(&mi->q_empty);
        if (h == 1) umma2_commit(&mi->w_empty[st]);
        PT_LAP(5);
      }
      umma2_commit(&mi->dz_full);
    }
    __syncwarp();
  } else {
    // =========================================== EPILOGUE (both CTAs) ====================================
    const int e = warp;                   // 0..15
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int r = (int)lane_base + lane;  // cell row of this CTA's unit
    const int cq = e >> 2;                // 16-gene quarter of the 64-gene half
    const int j = j_block0 + (int)rank * UC + r;
    const bool valid = j < a.B;
    const float c = valid ? a.smp_c[(long long)s_fixed * a.B + j] : 0.f;
    const int brow = valid ? ws.brow[j] : 0;
    const uint32_t r_rempty[2] = {mapa_rank(&mi->r_empty[0], 0), mapa_rank(&mi->r_empty[1], 0)};
    const uint32_t r_qfull = mapa_rank(&mi->q_full, 0);
    float ll = 0.f, esum = 0.f;
    auto x_ptr = [&](int u) {
      const int tile_u = 2 * (pt_begin + (u >> 1)) + (u & 1);
      return ws.ximg + ((((long long)tile_u * units_total_B + unit_u) * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4;
    };
    float4 x4[4];
    if (U > 0) {
      const float* xi = x_ptr(0);
#pragma unroll
      for (int v = 0; v < 4; ++v) x4[v] = *reinterpret_cast<const float4*>(xi + v * 128);
    }
    for (int u = 0; u < U; ++u) {
      const int oi = u >> 1, h = u & 1, st = oi & 1;
      const uint32_t par = (uint32_t)(oi >> 1) & 1u;
      PT_MARK();
      float xs16[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) { xs16[v * 4] = x4[v].x; xs16[v * 4 + 1] = x4[v].y; xs16[v * 4 + 2] = x4[v].z; xs16[v * 4 + 3] = x4[v].w; }
      if (u + 1 < U) {   // the next sub-unit's counts are in flight while this one is processed
        const float* xi = x_ptr(u + 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) x4[v] = *reinterpret_cast<const float4*>(xi + v * 128);
      }
      PT_LAP(6);
      if (h == 0) PT_WAIT(2, pair_wait(&mi->w_full[st], par, 10));   // the gene-scale tiles ride with the W stage
      PT_WAIT(0, pair_wait(&mi->r_full[st], par, 11));
      tc_fence_after();
      PT_MARK();
      float R[16], q[16];
      tmem_ld16(tmem + (lane_base << 16) + (uint32_t)(st * (2 * TG) + h * TG + cq * 16), R);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_rempty[st]);   // this warp has read its part of R[st] for sub-unit h
      PT_LAP(3);
      const float* srow_sm =
          reinterpret_cast<const float*>(sm + L.w[st] + L.sg_off + (uint32_t)h * (uint32_t)(MAX_NB * TG * 4)) + brow * TG + cq * 16;
      float es = 0.f;
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
        const int g0 = (2 * (pt_begin + oi) + h) * TG + cq * 16;
        for (int i = 0; i < 16; ++i)
          if (g0 + i < a.G) g_dbg_R[((long long)s_fixed * a.B + j) * a.G + g0 + i] = R[i];
      }
      PT_LAP(4);
      PT_WAIT(1, pair_wait(&mi->q_empty, ((uint32_t)u & 1u) ^ 1u, 12));
      PT_MARK();
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
      fence_async_q();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_qfull);
      PT_LAP(5);
      esum += es;
    }

    // ---- drain dz0 (this thread: cell row r, every 4th 16-column chunk of k0) ----
    PT_WAIT(7, pair_wait(&mi->dz_full, 0, 13));
    tc_fence_after();
    PT_MARK();
    if (U > 0) {
      float* dst = ws.part + (((long long)s_fixed * gm.NGp + range) * a.B + j) * a.K0;
      for (int cc = cq; cc < KP / 16; cc += 4) {
        float v[16];
        tmem_ld16(tmem_dz + (lane_base << 16) + (uint32_t)(cc * 16), v);
        if (valid) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cc * 16 + i < a.K0) dst[cc * 16 + i] = v[i];
        }
      }
      if (valid && esum != 0.f) atomicAdd(&a.G_c[(long long)s_fixed * a.B + j], a.scale * esum / c);
    }
    PT_LAP(8);
    // loss: reduce over the epilogue threads
    double v = warp_sum_d((double)ll);
    if (lane == 0) mi->red[e] = v;
    named_bar_sync(2, E_THREADS);
    if (WANT_LL && e == 0 && lane == 0) {
      double tot = 0.0;
      for (int i = 0; i < E_WARPS; ++i) tot += mi->red[i];
      atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
    }
  }

#ifdef SCDEF_TC_TIMING
  if ((int)(blockIdx.x >> 1) == (int)(gridDim.x >> 2) && blockIdx.y == 0 && lane == 0 &&
      (warp == COPY_WARP || warp == MMA_WARP || warp == 0 || warp == 15)) {
    const int role = warp == COPY_WARP ? 0 : (warp == MMA_WARP ? 1 : (warp == 0 ? 2 : 3));
    for (int i = 0; i < 15; ++i) g_pair_timing[rank][role][i] = (unsigned long long)tacc[i];
    g_pair_timing[rank][role][15] = (unsigned long long)(clock64() - t_start);
  }
#endif
  // nobody leaves while the peer may still signal into this CTA's shared memory or read its tensor memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// =============================================================================================================
// Persistent variant (SCDEF_TC_PAIR=2): the (MC sample, pair-tile) space of one 256-cell block is flattened,
// f = s * n_pt + pt, and cut into P contiguous chunks, one per CTA pair (P = 74 / number of cell blocks), so every pair
// slot gets the same number of pair-tiles (+-1) and pays the prologue (barrier init, TMEM allocation, two cluster
// syncs) once.  Measured on the item version above at C5 (tests/pair_timing.py dump): 45k of the 178k cycles of a
// 20-pair-tile item are per-item overhead (prologue, dz drain with scalar stores) and three items run back to back per slot.
// A chunk crosses sample boundaries ("segments"): at a boundary the z unit is replaced (z_empty: committed after the
// LAST rate GEMM of the segment, which is issued one pair-tile early, so the reload hides behind that pair-tile's
// epilogue), dz0 switches to the other of two TMEM accumulators (dz_full / dz_empty), and the epilogue drains the
// finished one into its slot of the partial buffer: slot = chunk - chunk_of(first pair-tile of the sample).
// =============================================================================================================
constexpr int FLAT_TMEM_DZ = 256, FLAT_DZ_STRIDE = 128;   // R[2] at 0 / 128; dz[2] at 256 / 384

// tcgen05.ld split into issue and wait, so that other work can sit between them.  The wait names the destination
// registers as in/out operands: the compiler cannot move a use of them above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

struct FlatMisc {
  uint64_t z_full, pz_full, z_empty, w_full[2], pw_full[2], w_empty[2], r_full[2], r_empty[2], q_full, q_empty, dz_full[2], dz_empty[2];
  uint64_t q_full4[4], q_empty4[4];   // FINE: the Q hand-off per k-step of the second GEMM (= per 16-gene quarter cq)
  uint32_t tmem_base, pad;
  double red[32];
};
static_assert(sizeof(FlatMisc) <= 1024, "FlatMisc must fit the misc block of plan_pair_smem");

// chunk that contains flat index f when [0, T) is cut at floor(c * T / P)
__host__ __device__ inline int flat_chunk_of(long long f, long long T, int P) { return (int)(((f + 1) * P - 1) / T); }

struct FlatCur {   // position in the flattened space
  int s, pt, g;    // MC sample, pair-tile of the sample, segment index inside the chunk
  __device__ __forceinline__ void next(int n_pt) { if (++pt == n_pt) { pt = 0; ++s; ++g; } }
};

// WIDE (SCDEF_TC_PAIR=3, never run on a device yet): the epilogue handles a whole pair-tile (both 64-gene halves) per
// barrier round instead of one half: per-element overhead (barrier fast paths, address arithmetic, tcgen05.ld issue,
// r_empty arrival) halves, and the second half's tcgen05.ld is in flight while the first half's Q tile is stored.
// FINE (SCDEF_TC_PAIR=4 = WIDE + FINE, never run on a device yet): the Q tile is handed over per k-step of the second
// GEMM.  The four warps that own the 16-gene quarter cq arrive on q_full4[cq]; the MMA thread issues that k-step's three
// MMAs as soon as they have, and commits q_empty4[cq].  The quarters stop moving in lockstep (each is released ~0.3k
// cycles after the previous one) and a warp waits only for its own slice.  Modelled in tests/test_pair_protocol.py.
template <bool WANT_LL, bool WIDE, bool FINE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PIPE_NT, 1)
    k_lik_pair_flat(const LikTcArgs a, const Geom gm, const TcScratch ws, double* loss_out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const PairSmem L = plan_pair_smem(KP);
  FlatMisc* mi = reinterpret_cast<FlatMisc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  PT_DECL;

  // ---- work of this pair ----
  const int n_pt = gm.n_tiles / 2, P = gm.Pflat;
  const long long T = (long long)a.S * n_pt;
  const int chunk = blockIdx.x >> 1;
  const long long f0 = (long long)chunk * T / P, f1 = (long long)(chunk + 1) * T / P;
  const int n_outer = (int)(f1 - f0);
  const int U = 2 * n_outer;
  const FlatCur cur0 = {(int)(f0 / n_pt), (int)(f0 % n_pt), 0};
  const int n_seg = n_outer > 0 ? (int)((f1 - 1) / n_pt) - cur0.s + 1 : 0;
  const int j_block0 = blockIdx.y * CB;
  const int units_total_B = (a.B + UC - 1) / UC;
  const int unit_u = j_block0 / UC + (int)rank;

  // ---- prologue ----
  if (threadIdx.x == 0) {
    mbar_init(&mi->z_full, 1);
    mbar_init(&mi->pz_full, 1);
    mbar_init(&mi->z_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&mi->w_full[i], 1);
      mbar_init(&mi->pw_full[i], 1);
      mbar_init(&mi->w_empty[i], 1);
      mbar_init(&mi->r_full[i], 1);
      mbar_init(&mi->r_empty[i], (WIDE ? 2 : 4) * E_WARPS);   // elected arrivals per pair-tile: 16 warps x 2 CTAs (x 2 halves)
      mbar_init(&mi->dz_full[i], 1);
      mbar_init(&mi->dz_empty[i], 2 * E_WARPS);
    }
    mbar_init(&mi->q_full, 2 * E_WARPS);
    mbar_init(&mi->q_empty, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&mi->q_full4[i], 2 * E_WARPS / 4);   // the quarter's 4 warps in each CTA
      mbar_init(&mi->q_empty4[i], 1);
    }
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
  const uint32_t z_bytes = 2u * L.z_plane;
  const uint32_t sg_bytes = (uint32_t)(a.nb * TG * 4);
  const uint32_t w_bytes = 2u * L.w_plane + 2u * sg_bytes + 4u * L.b2_piece;
  const PipeSmem IL = plan_pipe_smem(KP);
  auto first_in_seg = [&](int oi, const FlatCur& c) { return c.pt == 0 || oi == 0; };
  auto last_in_seg = [&](int oi, const FlatCur& c) { return c.pt == n_pt - 1 || oi == n_outer - 1; };

  if (warp == COPY_WARP) {
    // =========================================== COPY (TMA), each CTA for itself ===========================
    if (lane == 0) {
      FlatCur c = cur0;
      for (int oi = 0; oi < n_outer; ++oi, c.next(n_pt)) {
        const int st = oi & 1;
        const uint8_t* img0 = ws.wimg + ((long long)c.s * gm.n_tiles + 2 * c.pt) * IL.w_unit;
        const uint8_t* img1 = img0 + IL.w_unit;
        const uint8_t* own = rank ? img1 : img0;
        PT_WAIT(0, pair_wait(&mi->w_empty[st], ((oi >> 1) & 1) ^ 1, 1));
        mbar_arrive_expect_tx(&mi->w_full[st], w_bytes);
        const uint32_t dst = sbase + L.w[st];
        bulk_g2s(dst, own, 2u * L.w_plane, &mi->w_full[st]);
        bulk_g2s(dst + L.sg_off, img0 + 2u * IL.w_plane, sg_bytes, &mi->w_full[st]);
        bulk_g2s(dst + L.sg_off + (uint32_t)(MAX_NB * TG * 4), img1 + 2u * IL.w_plane, sg_bytes, &mi->w_full[st]);
        const uint32_t row_off = rank * L.b2_piece;
        bulk_g2s(dst + L.b2_off + 0u * L.b2_piece, img0 + row_off, L.b2_piece, &mi->w_full[st]);
        bulk_g2s(dst + L.b2_off + 1u * L.b2_piece, img0 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]);
        bulk_g2s(dst + L.b2_off + 2u * L.b2_piece, img1 + row_off, L.b2_piece, &mi->w_full[st]);
        bulk_g2s(dst + L.b2_off + 3u * L.b2_piece, img1 + IL.w_plane + row_off, L.b2_piece, &mi->w_full[st]);
        if (first_in_seg(oi, c)) {
          // the z unit of the new sample replaces the old one once the old sample's last rate GEMM has read it
          PT_WAIT(1, pair_wait(&mi->z_empty, ((uint32_t)c.g & 1u) ^ 1u, 17));
          mbar_arrive_expect_tx(&mi->z_full, z_bytes);
          bulk_g2s(sbase + L.z, ws.zimg + ((long long)c.s * units_total_B + unit_u) * z_bytes, z_bytes, &mi->z_full);
        }
      }
      // see the last multicast commits land before this CTA can exit
      for (int st = 0; st < 2; ++st) {
        const int uses = (n_outer + 1 - st) / 2;
        if (uses > 0) pair_wait(&mi->w_empty[st], (uint32_t)(uses - 1) & 1u, 15);
      }
      if (U > 0) {
        if (FINE) { for (int i = 0; i < 4; ++i) pair_wait(&mi->q_empty4[i], (uint32_t)(U - 1) & 1u, 16); }
        else pair_wait(&mi->q_empty, (uint32_t)(U - 1) & 1u, 16);
      }
      if (n_seg > 0) pair_wait(&mi->z_empty, (uint32_t)(n_seg - 1) & 1u, 18);
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (lane == 0 && !leader) {
      // =========================================== RELAY (peer) ==========================================
      const uint32_t r_pz = mapa_rank(&mi->pz_full, 0), r_pw0 = mapa_rank(&mi->pw_full[0], 0), r_pw1 = mapa_rank(&mi->pw_full[1], 0);
      FlatCur c = cur0;
      for (int oi = 0; oi < n_outer; ++oi, c.next(n_pt)) {
        PT_WAIT(0, pair_wait(&mi->w_full[oi & 1], (oi >> 1) & 1, 3));
        mbar_arrive_remote((oi & 1) ? r_pw1 : r_pw0);
        if (first_in_seg(oi, c)) {
          PT_WAIT(1, pair_wait(&mi->z_full, (uint32_t)c.g & 1u, 2));
          mbar_arrive_remote(r_pz);
        }
      }
    } else if (lane == 0) {
      // =========================================== MMA ISSUER (leader) ===================================
      const uint32_t idesc1 = umma_idesc(256, 2 * TG, 0, 1);
      const uint32_t idesc2 = umma_idesc(256, KP, 0, 0);
      const uint64_t dz_k = umma_desc(sbase + L.z, 128u, L.z_rs);
      const uint64_t dwa_m[2] = {umma_desc(sbase + L.w[0], W_RS, 128u), umma_desc(sbase + L.w[1], W_RS, 128u)};
      const uint64_t db2_k[2] = {umma_desc(sbase + L.w[0] + L.b2_off, 128u, W_RS), umma_desc(sbase + L.w[1] + L.b2_off, 128u, W_RS)};
      const uint64_t dq_k = umma_desc(sbase + L.q, 128u, Q_RS);
      const uint64_t zpl = L.z_plane >> 4, wpl = L.w_plane >> 4, qpl = L.q_plane >> 4, b2pl = L.b2_piece >> 4;
      const int nk1 = KP / 16;
      FlatCur c1 = cur0, c2 = cur0;   // cursors of the next rate GEMM (oi = next1) and of the second GEMM (oi = u >> 1)
      auto ready1 = [&](int oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        if (first_in_seg(oi, c1) && !(mbar_test(&mi->z_full, (uint32_t)c1.g & 1u) && mbar_test_remote(&mi->pz_full, (uint32_t)c1.g & 1u)))
          return false;
        return mbar_test(&mi->w_full[st], par) && mbar_test_remote(&mi->pw_full[st], par) &&
               mbar_test_remote(&mi->r_empty[st], par ^ 1u);
      };
      auto issue_mma1 = [&](int oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        if (first_in_seg(oi, c1)) {
          PT_WAIT(6, pair_wait(&mi->z_full, (uint32_t)c1.g & 1u, 7));
          PT_WAIT(6, pair_wait_remote(&mi->pz_full, (uint32_t)c1.g & 1u, 8));
        }
        PT_WAIT(0, pair_wait(&mi->w_full[st], par, 4));
        PT_WAIT(1, pair_wait_remote(&mi->pw_full[st], par, 5));
        PT_WAIT(2, pair_wait_remote(&mi->r_empty[st], par ^ 1u, 6));
        tc_fence_after();
        PT_MARK();
        uint64_t da = dz_k, db = dwa_m[st];
        const uint32_t d = tmem + (uint32_t)st * (2 * TG);
        umma2_bf16(d, da + zpl, db, idesc1, 0u);
        umma2_bf16(d, da, db + wpl, idesc1, 1u);
        umma2_bf16(d, da, db, idesc1, 1u);
        for (int ks = 1; ks < nk1; ++ks) {
          da += 16;
          db += 2 * (W_RS >> 4);
          umma2_bf16(d, da + zpl, db, idesc1, 1u);
          umma2_bf16(d, da, db + wpl, idesc1, 1u);
          umma2_bf16(d, da, db, idesc1, 1u);
        }
        umma2_commit(&mi->r_full[st]);
        if (last_in_seg(oi, c1)) umma2_commit(&mi->z_empty);   // the sample's z unit has been read for the last time
        PT_LAP(4);
        c1.next(n_pt);
      };
      int next1 = 0;
      uint64_t* const q_gate = FINE ? &mi->q_full4[0] : &mi->q_full;
      for (int u = 0; u < U; ++u) {
        const int oi = u >> 1, h = u & 1, st = oi & 1;
        long long t_poll = 0;
#ifdef SCDEF_TC_TIMING
        const long long t_loop0 = clock64(), t_in1 = tacc[0] + tacc[1] + tacc[2] + tacc[4] + tacc[6];
#endif
        for (;;) {
          if (next1 < n_outer && next1 <= oi + 1 && (next1 == oi || ready1(next1))) { issue_mma1(next1); ++next1; continue; }
          if (mbar_test_remote(q_gate, (uint32_t)u & 1u)) break;
          if (next1 > oi + 1 || next1 >= n_outer) { pair_wait_remote(q_gate, (uint32_t)u & 1u, 9); break; }
          if (t_poll == 0) t_poll = clock64();
          else if (clock64() - t_poll > PAIR_DEADLINE) pair_stuck(14, (int)blockIdx.x, u, next1);
        }
#ifdef SCDEF_TC_TIMING
        tacc[3] += (clock64() - t_loop0) - (tacc[0] + tacc[1] + tacc[2] + tacc[4] + tacc[6] - t_in1);
#endif
        const bool seg_first = first_in_seg(oi, c2) && h == 0, seg_last = last_in_seg(oi, c2) && h == 1;
        const uint32_t dzb = (uint32_t)c2.g & 1u;
        if (seg_first) PT_WAIT(7, pair_wait_remote(&mi->dz_empty[dzb], (((uint32_t)c2.g >> 1) & 1u) ^ 1u, 19));   // drained two segments ago
        tc_fence_after();
        PT_MARK();
        const uint32_t d = tmem + FLAT_TMEM_DZ + dzb * FLAT_DZ_STRIDE;
        uint64_t da = dq_k, db = db2_k[st] + (uint64_t)(2 * h) * b2pl;
        if (FINE) {
#pragma unroll
          for (int ks = 0; ks < TG / 16; ++ks) {
            if (ks > 0) {   // the gate above was quarter 0
              PT_WAIT(3, pair_wait_remote(&mi->q_full4[ks], (uint32_t)u & 1u, 20));
              tc_fence_after();
            }
            umma2_bf16(d, da + qpl, db, idesc2, (seg_first && ks == 0) ? 0u : 1u);
            umma2_bf16(d, da, db + b2pl, idesc2, 1u);
            umma2_bf16(d, da, db, idesc2, 1u);
            umma2_commit(&mi->q_empty4[ks]);
            da += 16;
            db += 16;
          }
        } else {
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
        }
        if (h == 1) umma2_commit(&mi->w_empty[st]);
        if (seg_last) umma2_commit(&mi->dz_full[dzb]);
        PT_LAP(5);
        if (h == 1) c2.next(n_pt);
      }
    }
    __syncwarp();
  } else {
    // =========================================== EPILOGUE (both CTAs) ====================================
    const int e = warp;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int r = (int)lane_base + lane;
    const int cq = e >> 2;
    const int j = j_block0 + (int)rank * UC + r;
    const bool valid = j < a.B;
    const int brow = valid ? ws.brow[j] : 0;
    const uint32_t r_rempty[2] = {mapa_rank(&mi->r_empty[0], 0), mapa_rank(&mi->r_empty[1], 0)};
    const uint32_t r_dzempty[2] = {mapa_rank(&mi->dz_empty[0], 0), mapa_rank(&mi->dz_empty[1], 0)};
    const uint32_t r_qfull = mapa_rank(FINE ? &mi->q_full4[cq] : &mi->q_full, 0);
    uint64_t* const q_empty_bar = FINE ? &mi->q_empty4[cq] : &mi->q_empty;
    float ll = 0.f, esum = 0.f, c = 0.f;
    // one 64-gene half of a pair-tile: R (16 columns of this thread's cell row) -> Q, log-likelihood, sum of E
    auto half_math = [&](const float4 (&xv4)[4], const uint32_t (&Rr)[16], const float* srow_sm, float (&q)[16], float& es) {
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 s4v = *reinterpret_cast<const float4*>(srow_sm + v * 4);
        const float ss[4] = {s4v.x, s4v.y, s4v.z, s4v.w};
        const float xs[4] = {xv4[v].x, xv4[v].y, xv4[v].z, xv4[v].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float Rv = __uint_as_float(Rr[v * 4 + i]);
          const float lam = (c * ss[i]) * Rv;
          const float E = xs[i] - lam;
          q[v * 4 + i] = __fdividef(E, fmaxf(Rv, 1e-37f));
          if (WANT_LL) {
            const float t = xs[i] * __logf(lam);
            ll += (xs[i] > 0.f ? t : 0.f) - lam;
          }
          es += E;
        }
      }
    };
    auto store_q = [&](const float (&q)[16]) {
      uint4 hi, lo;
      const uint32_t off = L.q + (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u + (uint32_t)(cq * 2) * 128u;
      split8(q, hi, lo);
      *reinterpret_cast<uint4*>(sm + off) = hi;
      *reinterpret_cast<uint4*>(sm + off + L.q_plane) = lo;
      split8(q + 8, hi, lo);
      *reinterpret_cast<uint4*>(sm + off + 128u) = hi;
      *reinterpret_cast<uint4*>(sm + off + 128u + L.q_plane) = lo;
      fence_async_q();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_qfull);
    };
    // the sample's dz0 is complete: drain its accumulator (cell row r, every 4th 16-column chunk of k0)
    auto drain_segment = [&](const FlatCur& p) {
      const uint32_t dzb = (uint32_t)p.g & 1u;
      PT_WAIT(7, pair_wait(&mi->dz_full[dzb], ((uint32_t)p.g >> 1) & 1u, 13));
      tc_fence_after();
      PT_MARK();
      const int slot = chunk - flat_chunk_of((long long)p.s * n_pt, T, P);
      float* dst = ws.part + (((long long)p.s * gm.NGp + slot) * a.B + j) * a.K0;
      const uint32_t tdz = tmem + FLAT_TMEM_DZ + dzb * FLAT_DZ_STRIDE + (lane_base << 16);
      const bool vec4 = (a.K0 & 3) == 0;
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
      if (valid && esum != 0.f) atomicAdd(&a.G_c[(long long)p.s * a.B + j], a.scale * esum / c);
      esum = 0.f;
      PT_LAP(8);
    };
    if (WIDE) {
      // this thread's first float4 of 64-gene tile t of the count image sits at x_base + t * x_tile
      const long long x_tile = (long long)units_total_B * (4 * 16 * 128);
      const float* x_base = ws.ximg + (((long long)unit_u * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4;
      float4 xa[4], xb[4];   // counts of half 0 / half 1, loaded one pair-tile ahead
      FlatCur cc = cur0;
      if (n_outer > 0) {
        const float* xi = x_base + (long long)(2 * cc.pt) * x_tile;
#pragma unroll
        for (int v = 0; v < 4; ++v) { xa[v] = *reinterpret_cast<const float4*>(xi + v * 128); xb[v] = *reinterpret_cast<const float4*>(xi + x_tile + v * 128); }
      }
      for (int oi = 0; oi < n_outer; ++oi) {
        const int st = oi & 1;
        const uint32_t par = (uint32_t)(oi >> 1) & 1u;
        const bool more = oi + 1 < n_outer;
        FlatCur cn = cc;
        cn.next(n_pt);
        const float* xnext = x_base + (long long)(2 * cn.pt) * x_tile;
        if (first_in_seg(oi, cc)) c = valid ? a.smp_c[(long long)cc.s * a.B + j] : 0.f;
        const float* sg0 = reinterpret_cast<const float*>(sm + L.w[st] + L.sg_off) + brow * TG + cq * 16;
        const uint32_t taddr = tmem + (lane_base << 16) + (uint32_t)(st * (2 * TG) + cq * 16);
        PT_MARK();
        PT_WAIT(2, pair_wait(&mi->w_full[st], par, 10));
        PT_WAIT(0, pair_wait(&mi->r_full[st], par, 11));
        tc_fence_after();
        PT_MARK();
        uint32_t R0[16], R1[16];
        float q[16];
        tmem_ld16_issue(taddr, R0);
        tmem_ld16_wait(R0);
        PT_LAP(3);
        float es = 0.f;
        half_math(xa, R0, sg0, q, es);
        if (more) {
#pragma unroll
          for (int v = 0; v < 4; ++v) xa[v] = *reinterpret_cast<const float4*>(xnext + v * 128);
        }
        tmem_ld16_issue(taddr + (uint32_t)TG, R1);   // in flight while the first half's Q tile is stored
        PT_LAP(4);
        PT_WAIT(1, pair_wait(q_empty_bar, 1u, 12));  // sub-unit 2 oi: parity (u & 1) ^ 1 with u even
        PT_MARK();
        store_q(q);
        tmem_ld16_wait(R1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(r_rempty[st]);   // both halves of R[st] are in registers
        PT_LAP(5);
        half_math(xb, R1, sg0 + MAX_NB * TG, q, es);
        if (more) {
#pragma unroll
          for (int v = 0; v < 4; ++v) xb[v] = *reinterpret_cast<const float4*>(xnext + x_tile + v * 128);
        }
        PT_LAP(4);
        PT_WAIT(1, pair_wait(q_empty_bar, 0u, 12));  // sub-unit 2 oi + 1
        PT_MARK();
        store_q(q);
        PT_LAP(5);
        esum += es;
        if (last_in_seg(oi, cc)) drain_segment(cc);
        cc = cn;
      }
    } else {
    FlatCur cc = cur0, cx = cur0;   // cursors of the sub-unit being processed and of the counts being prefetched
    auto x_ptr = [&](const FlatCur& p, int h) {
      const int tile_u = 2 * p.pt + h;
      return ws.ximg + ((((long long)tile_u * units_total_B + unit_u) * 4 + (warp & 3)) * 16 + cq * 4) * 128 + lane * 4;
    };
    float4 x4[4];
    if (U > 0) {
      const float* xi = x_ptr(cx, 0);
#pragma unroll
      for (int v = 0; v < 4; ++v) x4[v] = *reinterpret_cast<const float4*>(xi + v * 128);
    }
    for (int u = 0; u < U; ++u) {
      const int oi = u >> 1, h = u & 1, st = oi & 1;
      const uint32_t par = (uint32_t)(oi >> 1) & 1u;
      PT_MARK();
      float xs16[16];
#pragma unroll
      for (int v = 0; v < 4; ++v) { xs16[v * 4] = x4[v].x; xs16[v * 4 + 1] = x4[v].y; xs16[v * 4 + 2] = x4[v].z; xs16[v * 4 + 3] = x4[v].w; }
      if (u + 1 < U) {
        if (h == 1) cx.next(n_pt);
        const float* xi = x_ptr(cx, h ^ 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) x4[v] = *reinterpret_cast<const float4*>(xi + v * 128);
      }
      if (h == 0 && first_in_seg(oi, cc)) c = valid ? a.smp_c[(long long)cc.s * a.B + j] : 0.f;
      PT_LAP(6);
      if (h == 0) PT_WAIT(2, pair_wait(&mi->w_full[st], par, 10));
      PT_WAIT(0, pair_wait(&mi->r_full[st], par, 11));
      tc_fence_after();
      PT_MARK();
      float R[16], q[16];
      tmem_ld16(tmem + (lane_base << 16) + (uint32_t)(st * (2 * TG) + h * TG + cq * 16), R);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_rempty[st]);
      PT_LAP(3);
      const float* srow_sm =
          reinterpret_cast<const float*>(sm + L.w[st] + L.sg_off + (uint32_t)h * (uint32_t)(MAX_NB * TG * 4)) + brow * TG + cq * 16;
      float es = 0.f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
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
            const float t = xv * __logf(lam);
            ll += (xv > 0.f ? t : 0.f) - lam;
          }
          es += E;
        }
      }
      if (g_dbg_R && valid) {
        const int g0 = (2 * cc.pt + h) * TG + cq * 16;
        for (int i = 0; i < 16; ++i)
          if (g0 + i < a.G) g_dbg_R[((long long)cc.s * a.B + j) * a.G + g0 + i] = R[i];
      }
      PT_LAP(4);
      PT_WAIT(1, pair_wait(q_empty_bar, ((uint32_t)u & 1u) ^ 1u, 12));
      PT_MARK();
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
      fence_async_q();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_qfull);
      PT_LAP(5);
      esum += es;

      if (h == 1) {
        if (last_in_seg(oi, cc)) {
          // ---- the sample's dz0 is complete: drain its accumulator (cell row r, every 4th 16-column chunk of k0) ----
          const uint32_t dzb = (uint32_t)cc.g & 1u;
          PT_WAIT(7, pair_wait(&mi->dz_full[dzb], ((uint32_t)cc.g >> 1) & 1u, 13));
          tc_fence_after();
          PT_MARK();
          const int slot = chunk - flat_chunk_of((long long)cc.s * n_pt, T, P);
          float* dst = ws.part + (((long long)cc.s * gm.NGp + slot) * a.B + j) * a.K0;
          const uint32_t tdz = tmem + FLAT_TMEM_DZ + dzb * FLAT_DZ_STRIDE + (lane_base << 16);
          const bool vec4 = (a.K0 & 3) == 0;
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
          if (valid && esum != 0.f) atomicAdd(&a.G_c[(long long)cc.s * a.B + j], a.scale * esum / c);
          esum = 0.f;
          PT_LAP(8);
        }
        cc.next(n_pt);
      }
    }
    }   // !WIDE
    double v = warp_sum_d((double)ll);
    if (lane == 0) mi->red[e] = v;
    named_bar_sync(2, E_THREADS);
    if (WANT_LL && e == 0 && lane == 0) {
      double tot = 0.0;
      for (int i = 0; i < E_WARPS; ++i) tot += mi->red[i];
      atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
    }
  }

#ifdef SCDEF_TC_TIMING
  if ((int)(blockIdx.x >> 1) == (int)(gridDim.x >> 2) && blockIdx.y == 0 && lane == 0 &&
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

// G_z0[s][j][k] += scale * sum over the slots the sample's chunks wrote (the others hold stale data)
__global__ void __launch_bounds__(256) k_tc_dz_reduce_flat(const LikTcArgs a, const float* part, int NGslots, int P, int n_pt) {
  const long long per = (long long)a.B * a.K0, total = (long long)a.S * per;
  const long long T = (long long)a.S * n_pt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per, r = i - s * per;
    const int nsl = flat_chunk_of((s + 1) * n_pt - 1, T, P) - flat_chunk_of(s * n_pt, T, P) + 1;
    float acc = 0.f;
    for (int q = 0; q < nsl; ++q) acc += part[(s * NGslots + q) * per + r];
    a.G_z0[i] += a.scale * acc;
  }
}
