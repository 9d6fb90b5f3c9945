// Fused Poisson-likelihood kernel on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// Math (reference `_scdef.py:2096-2136` + the W_0 parts of 2012-2045):
//   W^s = clip(exp(mu~ + sigma eps^s))            generated IN the kernel, never written to HBM
//   R   = z0^s W^s      (tcgen05.mma, split-bf16: hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM)
//   Lambda = c_n s_{b(n),g} R,  ll = sum x log Lambda - Lambda,  E = x - Lambda,  Q = E / R
//   LOCAL : dz0 = Q W^T  (second MMA, accumulated in TMEM over the CTA's gene tiles)
//   GLOBAL: dW  = z0^T Q (second MMA), then  d mu += W (.) dW,  d rho += W (.) dW (.) sigma eps
// The cells x genes rate matrix lives only in TMEM / registers.
//
// Shared-memory operand tiles use the no-swizzle UMMA canonical layout: "core matrices" of
// 8 rows x 16 bytes (8 bf16), row r of the matrix at +16*r.  A tile [rows][cols] is a grid of
// core matrices at  (row/8)*RS + (col/8)*128.  The SAME bytes serve as a K-major operand when
// the contiguous dimension is the contraction dimension (LBO = 128, SBO = RS) and as an MN-major
// operand when it is not (SBO = 128, LBO = RS), so z, W and Q are each stored once:
//   z [cell][k0] : A of MMA1 (K-major)   and A of the GLOBAL dW MMA (MN-major, M = k0)
//   W [k0][gene] : B of MMA1 (MN-major)  and B of the LOCAL dz MMA  (K-major,  N = k0)
//   Q [cell][gene]: A of the LOCAL dz MMA (K-major) and B of the GLOBAL dW MMA (MN-major)
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>

#include "lik_tc.cuh"

namespace scdef {

namespace {

constexpr int TG = 64;    // genes per tile (MMA N of the rate GEMM)
constexpr int CB = 256;   // cells per block: two M=128 halves
constexpr int NT = 256;   // threads per CTA (8 warps)
constexpr int MAX_NB = 8; // batches handled by the shared-memory column accumulators
constexpr int SM_TARGET = 148;
constexpr int UC = 128;   // cells per pipeline unit
constexpr uint32_t W_RS = 1024, Q_RS = 1024;  // row-group strides of the W and Q tiles (8 col groups x 128 B)
constexpr float V_LO_M = 1.0001e-15f, V_HI_M = 0.9999e15f;  // "inside the sample clip" with a margin for hi+lo rounding
constexpr float LOG_V_LO = -34.538776394910684f, LOG_V_HI = 34.538776394910684f;

__device__ float* g_dbg_R = nullptr;  // bring-up hook: when set, R is dumped as (S,B,G)

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  // start address [0,14), leading byte offset [16,30), stride byte offset [32,46) (all >>4),
  // version = 1 at [46,48), no swizzle (layout_type = 0 at [61,64))
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  // c_format f32 (1<<4), a/b format bf16 (1<<7, 1<<10), a_major bit 15, b_major bit 16, N>>3 at 17, M>>4 at 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split 8 floats into bf16 hi and lo = bf16(x - hi), packed as two 16-byte vectors.
// Packed conversions (F2FP.BF16.F32.PACK_AB, one instruction per pair, round-to-nearest).
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);  // .x = a (low half), .y = b (high half)
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(x[2 * i], x[2 * i + 1]);
    const float r0 = x[2 * i] - __uint_as_float(h[i] << 16);
    const float r1 = x[2 * i + 1] - __uint_as_float(h[i] & 0xFFFF0000u);
    l[i] = pack_bf16x2(r0, r1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// ---- shared-memory plan -------------------------------------------------------------------------
struct Smem {
  uint32_t z_hi, z_lo, w_hi, w_lo, q_hi, q_lo, misc, total;
  uint32_t z_rs;  // row-group stride of the z tile
};
__host__ __device__ inline Smem plan_smem(int KP) {
  Smem s;
  s.z_rs = (uint32_t)(KP / 8) * 128u;
  uint32_t zp = 32u * s.z_rs, wp = (uint32_t)(KP / 8) * W_RS, qp = 32u * Q_RS;
  s.z_hi = 0; s.z_lo = zp; s.w_hi = 2 * zp; s.w_lo = 2 * zp + wp; s.q_hi = 2 * zp + 2 * wp; s.q_lo = s.q_hi + qp;
  s.misc = s.q_lo + qp;
  s.total = s.misc + 4096;
  return s;
}
struct Misc {                  // lives at Smem.misc
  uint64_t bar;
  uint32_t tmem_base, pad;
  double red[32];
  float colacc[MAX_NB][TG];    // sum_k ctilde[b][k] W[k][g]      (GLOBAL)
  float rowU[128], rowW[128];  // per-k0 sums of log W and W over the tile's genes (prior terms)
};

struct Geom {   // work decomposition of the monolithic kernels (likelihood-off phases, SCDEF_TC_MONO)
  int KP, n_tiles, n_chunks, NG, tiles_per_range, n_cblocks;
};

// ---- (a) z0^s block -> bf16 hi/lo planes --------------------------------------------------------
__device__ __forceinline__ void load_z(const LikTcArgs& a, uint8_t* sm, const Smem& L, int KP, int s, int j0) {
  const int nchunk = KP / 8;
  const float* zs = a.smp_z0 + (long long)s * a.B * a.K0;
  for (int t = threadIdx.x; t < CB * nchunk; t += NT) {
    const int r = t % CB, c = t / CB;
    const int j = j0 + r, k = c * 8;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (j < a.B) {
      const float* row = zs + (long long)j * a.K0;
      if ((a.K0 & 3) == 0) {
        if (k + 4 <= a.K0) { float4 v = *reinterpret_cast<const float4*>(row + k); x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }
        if (k + 8 <= a.K0) { float4 v = *reinterpret_cast<const float4*>(row + k + 4); x[4] = v.x; x[5] = v.y; x[6] = v.z; x[7] = v.w; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (k + e < a.K0) x[e] = row[k + e];
      }
    }
    uint4 hi, lo;
    split8(x, hi, lo);
    const uint32_t off = (uint32_t)(r >> 3) * L.z_rs + (uint32_t)c * 128u + (uint32_t)(r & 7) * 16u;
    *reinterpret_cast<uint4*>(sm + L.z_hi + off) = hi;
    *reinterpret_cast<uint4*>(sm + L.z_lo + off) = lo;
  }
}

// ---- (b) W^s tile: sample, split, and the per-row / per-column side sums ---------------------------
template <bool GLOBAL>
__device__ __forceinline__ void gen_w(const LikTcArgs& a, uint8_t* sm, const Smem& L, Misc* mi, int KP, int s, int g0,
                                      const float* ctilde_s, bool want_stats, bool want_col) {
  const int nkg = KP / 8;
  for (int t = threadIdx.x; t < nkg * 64; t += NT) {
    const int kk = t & 7, cg = (t >> 3) & 7, kg = t >> 6;
    const int k = kg * 8 + kk, g = g0 + cg * 8;
    float w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) w[e] = 0.f;
    if (k < a.K0 && g < a.G) {
      const long long base = (long long)k * a.G + g;
      float4 m0 = *reinterpret_cast<const float4*>(a.mu_W0 + base), m1 = *reinterpret_cast<const float4*>(a.mu_W0 + base + 4);
      float4 r0 = *reinterpret_cast<const float4*>(a.rho_W0 + base), r1 = *reinterpret_cast<const float4*>(a.rho_W0 + base + 4);
      float4 e0, e1;
      if (a.eps_W0) {
        const float* ep = a.eps_W0 + (long long)s * a.K0 * a.G + base;
        e0 = *reinterpret_cast<const float4*>(ep); e1 = *reinterpret_cast<const float4*>(ep + 4);
      } else {
        e0 = philox_normal4(a.key, a.tag, (uint32_t)s, (uint32_t)(base >> 2));
        e1 = philox_normal4(a.key, a.tag, (uint32_t)s, (uint32_t)(base >> 2) + 1u);
      }
      const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      const float rho[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      const float ep[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
      float su = 0.f, sw = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float mut = fminf(fmaxf(mu[e], LOG_1E_10), LOG_1E6);
        float sig = __expf(fminf(fmaxf(rho[e], LOG_1E_10), LOG_1E10));
        float u = fminf(fmaxf(fmaf(sig, ep[e], mut), LOG_V_LO), LOG_V_HI);
        w[e] = __expf(u);
        su += u;
        sw += w[e];
      }
      if (want_stats) {
        atomicAdd(&mi->rowU[k], su);
        atomicAdd(&mi->rowW[k], sw);
      }
      if (GLOBAL && want_col) {
        for (int b = 0; b < a.nb; ++b) {
          const float ct = ctilde_s[b * a.K0 + k];
#pragma unroll
          for (int e = 0; e < 8; ++e) atomicAdd(&mi->colacc[b][cg * 8 + e], ct * w[e]);
        }
      }
    }
    uint4 hi, lo;
    split8(w, hi, lo);
    const uint32_t off = (uint32_t)kg * W_RS + (uint32_t)cg * 128u + (uint32_t)kk * 16u;
    *reinterpret_cast<uint4*>(sm + L.w_hi + off) = hi;
    *reinterpret_cast<uint4*>(sm + L.w_lo + off) = lo;
  }
}

// ---- MMA1: R[half] = z W  (M=128, N=64, K=KP; 3 split products per k-step) ---------------------------
__device__ __forceinline__ void issue_mma1(uint32_t sbase, const Smem& L, int KP, uint32_t tmem, int nh) {
  const uint32_t idesc = umma_idesc(128, TG, /*A K-major*/ 0, /*B MN-major*/ 1);
  for (int h = 0; h < nh; ++h) {
    uint32_t acc = 0;
    for (int ks = 0; ks < KP / 16; ++ks) {
      const uint32_t za = (uint32_t)h * 16u * L.z_rs + (uint32_t)ks * 256u;
      const uint32_t wb = (uint32_t)ks * 2u * W_RS;
      const uint64_t a_hi = umma_desc(sbase + L.z_hi + za, 128u, L.z_rs), a_lo = umma_desc(sbase + L.z_lo + za, 128u, L.z_rs);
      const uint64_t b_hi = umma_desc(sbase + L.w_hi + wb, W_RS, 128u), b_lo = umma_desc(sbase + L.w_lo + wb, W_RS, 128u);
      umma_bf16(tmem + h * TG, a_lo, b_hi, idesc, acc);
      umma_bf16(tmem + h * TG, a_hi, b_lo, idesc, 1u);
      umma_bf16(tmem + h * TG, a_hi, b_hi, idesc, 1u);
      acc = 1u;
    }
  }
}

// ---- epilogue 1: R -> (ll, E sums, Q tile) ---------------------------------------------------------------
struct RowCtx {
  bool valid;
  float c;
  const float* srow;
  const float* xrow;
};

__device__ __forceinline__ void epilogue1(const LikTcArgs& a, uint8_t* sm, const Smem& L, uint32_t tmem, int s, int j0, int g0,
                                          int nh, const RowCtx& rc, float& ll, float& esum) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = warp >> 2;
  if (h >= nh) return;
  const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
  const int r = h * 128 + (int)lane_base + lane;
  const int j = j0 + r;
#pragma unroll 1
  for (int cc = 0; cc < TG / 16; ++cc) {
    float R[16];
    tmem_ld16(tmem + (lane_base << 16) + (uint32_t)(h * TG + cc * 16), R);
    float q[16];
    const int gb = g0 + cc * 16;
#pragma unroll
    for (int v4 = 0; v4 < 4; ++v4) {
      const int g = gb + v4 * 4;
      float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f);
      const bool ok = rc.valid && (g < a.G);
      if (ok) {
        x4 = *reinterpret_cast<const float4*>(rc.xrow + g);
        s4 = *reinterpret_cast<const float4*>(rc.srow + g);
      }
      const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float Rv = R[v4 * 4 + e];
        float qq = 0.f;
        if (ok) {
          const float lam = rc.c * ss[e] * Rv;
          const float E = xs[e] - lam;
          qq = __fdividef(E, Rv);
          ll += (xs[e] > 0.f ? xs[e] * __logf(lam) : 0.f) - lam;
          esum += E;
        }
        q[v4 * 4 + e] = qq;
      }
    }
    if (g_dbg_R && rc.valid) {
      for (int e = 0; e < 16; ++e)
        if (gb + e < a.G) g_dbg_R[((long long)s * a.B + j) * a.G + gb + e] = R[e];
    }
    uint4 hi, lo;
    const uint32_t off = (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u;
    split8(q, hi, lo);
    *reinterpret_cast<uint4*>(sm + L.q_hi + off + (uint32_t)(cc * 2) * 128u) = hi;
    *reinterpret_cast<uint4*>(sm + L.q_lo + off + (uint32_t)(cc * 2) * 128u) = lo;
    split8(q + 8, hi, lo);
    *reinterpret_cast<uint4*>(sm + L.q_hi + off + (uint32_t)(cc * 2 + 1) * 128u) = hi;
    *reinterpret_cast<uint4*>(sm + L.q_lo + off + (uint32_t)(cc * 2 + 1) * 128u) = lo;
  }
}

__device__ __forceinline__ RowCtx make_row(const LikTcArgs& a, int s, int j0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = (warp >> 2) * 128 + (warp & 3) * 32 + lane;
  const int j = j0 + r;
  RowCtx rc;
  rc.valid = j < a.B;
  rc.c = 1.f; rc.srow = a.smp_s; rc.xrow = a.X;
  if (rc.valid) {
    rc.c = a.smp_c[(long long)s * a.B + j];
    const int b = a.batch_id[a.idx[j]];
    rc.srow = a.smp_s + ((long long)s * a.nb + b) * a.G;
    rc.xrow = a.X + (a.idx_x ? a.idx_x[j] : (long long)j) * a.G;
  }
  return rc;
}

// flush the per-tile W row sums (prior terms of W_0) to the (S,K0,2) global accumulators
__device__ __forceinline__ void flush_rowstats(const LikTcArgs& a, Misc* mi, float* rowstats, int s) {
  for (int k = threadIdx.x; k < a.K0; k += NT) {
    atomicAdd(&rowstats[((long long)s * a.K0 + k) * 2 + 0], mi->rowU[k]);
    atomicAdd(&rowstats[((long long)s * a.K0 + k) * 2 + 1], mi->rowW[k]);
    mi->rowU[k] = 0.f;
    mi->rowW[k] = 0.f;
  }
}

struct TcScratch {
  float* rowstats;  // (S,K0,2)
  float* ctilde;    // (S,nb,K0)
  float* colsumX;   // (nb,G)
  float* part;      // LOCAL: (S,NG,B,K0)   GLOBAL: (n_chunks,2,K0,G)
  float* wmu;       // (K0,G) clip(mu) of W_0            } written once per pass by k_tc_wprep
  float* wsig;      // (K0,G) exp(clip(rho)) of W_0      }
  uint8_t* zimg;    // (S, ceil(B/128), 2 planes) z0 samples pre-split to bf16 hi/lo in the UMMA layout
  float* ximg;      // (n_tiles, units, 4 lane quarters, 16 chunks, 32 lanes, 4) minibatch counts, lane-per-row coalesced
  uint8_t* wimg;    // (S, n_tiles) W_0 sample tiles: [hi plane | lo plane | gene-scale tile (MAX_NB x 64 f32)]
  int* brow;        // (B) batch of each minibatch row
  long long* xoff;  // (B) element offset of each minibatch row in X
  // unified pair kernel (lik_flat.cuh)
  float* Ws;        // (S,nb,K0)  sum_g W^s[k][g] s^s[b][g]           } persist from the LOCAL to the GLOBAL pass
  float* rowsumx;   // (B)        sum_g x[j][g]
  uint8_t* zimg_g;  // (S, cell pair-tiles, 2 ranks) z0 column stages of the GLOBAL orientation
  float* ximg_t;    // counts, row = gene (GLOBAL orientation)
  float* foldtab;   // (S, 2 + nb, KP) W_0 prior parameters and rank-1 factors of the fold
};

// =====================================================================================================
// LOCAL pass: CTA = (MC sample s, gene range, cell block).  dz0 accumulates in TMEM over the range.
// =====================================================================================================
__global__ void __launch_bounds__(NT, 1) k_lik_tc_local(const LikTcArgs a, const Geom gm, const TcScratch ws, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const Smem L = plan_smem(KP);
  Misc* mi = reinterpret_cast<Misc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int s = blockIdx.x / gm.NG, range = blockIdx.x % gm.NG;
  const int j0 = blockIdx.y * CB;
  const int rows = min(CB, a.B - j0), nh = rows > 128 ? 2 : 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool want_stats = (blockIdx.y == 0);  // W-row sums are cell-independent: count them once

  if (threadIdx.x == 0) { mbar_init(&mi->bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 128; i += NT) { mi->rowU[i] = 0.f; mi->rowW[i] = 0.f; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&mi->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  load_z(a, sm, L, KP, s, j0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi->tmem_base;
  const uint32_t tmem_dz = tmem + 2 * TG;  // two halves of KP columns each
  const RowCtx rc = make_row(a, s, j0);
  const uint32_t idesc2 = umma_idesc(128, KP, 0, 0);
  float ll = 0.f, esum = 0.f;
  uint32_t phase = 0;
  const int t_begin = range * gm.tiles_per_range, t_end = min(gm.n_tiles, t_begin + gm.tiles_per_range);

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int g0 = tile * TG;
    gen_w<false>(a, sm, L, mi, KP, s, g0, nullptr, want_stats, false);
    fence_async_smem();
    __syncthreads();
    if (want_stats) flush_rowstats(a, mi, ws.rowstats, s);
    if (threadIdx.x == 0) {
      tc_fence_after();
      issue_mma1(sbase, L, KP, tmem, nh);
      umma_commit(&mi->bar);
    }
    mbar_wait(&mi->bar, phase);
    phase ^= 1;
    tc_fence_after();
    epilogue1(a, sm, L, tmem, s, j0, g0, nh, rc, ll, esum);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tc_fence_after();
      for (int h = 0; h < nh; ++h) {
        uint32_t acc = (tile > t_begin) ? 1u : 0u;
        for (int ks = 0; ks < TG / 16; ++ks) {
          const uint32_t qa = (uint32_t)h * 16u * Q_RS + (uint32_t)ks * 256u, wb = (uint32_t)ks * 256u;
          const uint64_t a_hi = umma_desc(sbase + L.q_hi + qa, 128u, Q_RS), a_lo = umma_desc(sbase + L.q_lo + qa, 128u, Q_RS);
          const uint64_t b_hi = umma_desc(sbase + L.w_hi + wb, 128u, W_RS), b_lo = umma_desc(sbase + L.w_lo + wb, 128u, W_RS);
          umma_bf16(tmem_dz + h * KP, a_lo, b_hi, idesc2, acc);
          umma_bf16(tmem_dz + h * KP, a_hi, b_lo, idesc2, 1u);
          umma_bf16(tmem_dz + h * KP, a_hi, b_hi, idesc2, 1u);
          acc = 1u;
        }
      }
      umma_commit(&mi->bar);
    }
    mbar_wait(&mi->bar, phase);  // W and Q tiles are free again, dz is up to date
    phase ^= 1;
    tc_fence_after();
  }

  // dz0 partial of this (s, range): TMEM -> global
  {
    const int h = warp >> 2;
    const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
    const int j = j0 + h * 128 + (int)lane_base + lane;
    if (h < nh) {
      float* dst = ws.part + (((long long)s * gm.NG + range) * a.B + j) * a.K0;
      for (int cc = 0; cc < KP / 16; ++cc) {
        float v[16];
        tmem_ld16(tmem_dz + (lane_base << 16) + (uint32_t)(h * KP + cc * 16), v);
        if (j < a.B) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (cc * 16 + e < a.K0) dst[cc * 16 + e] = v[e];
        }
      }
      if (rc.valid) atomicAdd(&a.G_c[(long long)s * a.B + j], a.scale * esum / rc.c);
    }
  }
  double tot = block_sum_d((double)ll, mi->red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// =====================================================================================================
// GLOBAL pass: CTA = (gene tile, chunk of MC samples).  d mu / d rho of the tile accumulate in
// registers over the chunk; per sample dW = z0^T Q comes from the second MMA.
// =====================================================================================================
__global__ void __launch_bounds__(NT, 1) k_lik_tc_global(const LikTcArgs a, const Geom gm, const TcScratch ws, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ __align__(1024) uint8_t sm[];
  const int KP = gm.KP;
  const Smem L = plan_smem(KP);
  Misc* mi = reinterpret_cast<Misc*>(sm + L.misc);
  const uint32_t sbase = smem_u32(sm);
  const int tile = blockIdx.x % gm.n_tiles, chunk = blockIdx.x / gm.n_tiles;
  const int g0 = tile * TG;
  const int s_begin = (int)((long long)chunk * a.S / gm.n_chunks), s_end = (int)((long long)(chunk + 1) * a.S / gm.n_chunks);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool lik = a.lik_on != 0;

  if (threadIdx.x == 0) { mbar_init(&mi->bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int i = threadIdx.x; i < 128; i += NT) { mi->rowU[i] = 0.f; mi->rowW[i] = 0.f; }
  for (int i = threadIdx.x; i < MAX_NB * TG; i += NT) (&mi->colacc[0][0])[i] = 0.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&mi->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi->tmem_base;
  const uint32_t tmem_dw = tmem + 2 * TG;

  // epilogue-2 ownership: k0 row = lane of the warp's TMEM quarter, 32 genes of the tile
  const uint32_t lane_base = (uint32_t)(warp & 3) * 32u;
  const int k2 = (int)lane_base + lane, ch = warp >> 2;
  const bool own = k2 < a.K0;
  float acc_mu[32], acc_rho[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) { acc_mu[e] = 0.f; acc_rho[e] = 0.f; }
  const uint32_t idesc2 = umma_idesc(128, TG, 1, 1);
  float ll = 0.f;
  uint32_t phase = 0;

  for (int s = s_begin; s < s_end; ++s) {
    gen_w<true>(a, sm, L, mi, KP, s, g0, ws.ctilde + (long long)s * a.nb * a.K0, true, lik);
    if (lik) {
      for (int cb = 0; cb < gm.n_cblocks; ++cb) {
        const int j0 = cb * CB;
        const int rows = min(CB, a.B - j0), nh = rows > 128 ? 2 : 1;
        load_z(a, sm, L, KP, s, j0);
        fence_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
          tc_fence_after();
          issue_mma1(sbase, L, KP, tmem, nh);
          umma_commit(&mi->bar);
        }
        const RowCtx rc = make_row(a, s, j0);
        mbar_wait(&mi->bar, phase);
        phase ^= 1;
        tc_fence_after();
        float esum = 0.f;
        epilogue1(a, sm, L, tmem, s, j0, g0, nh, rc, ll, esum);
        if (nh == 1 && warp >= 4) {  // rows 128..255 of the Q tile feed the dW MMA: keep them zero
          const int r = 128 + (warp & 3) * 32 + lane;
          const uint32_t off = (uint32_t)(r >> 3) * Q_RS + (uint32_t)(r & 7) * 16u;
          for (int c8 = 0; c8 < 8; ++c8) {
            *reinterpret_cast<uint4*>(sm + L.q_hi + off + c8 * 128u) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sm + L.q_lo + off + c8 * 128u) = make_uint4(0, 0, 0, 0);
          }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (threadIdx.x == 0) {
          tc_fence_after();
          uint32_t acc = cb > 0 ? 1u : 0u;
          const int nks = (rows + 15) / 16;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t za = (uint32_t)ks * 2u * L.z_rs, qb = (uint32_t)ks * 2u * Q_RS;
            const uint64_t a_hi = umma_desc(sbase + L.z_hi + za, L.z_rs, 128u), a_lo = umma_desc(sbase + L.z_lo + za, L.z_rs, 128u);
            const uint64_t b_hi = umma_desc(sbase + L.q_hi + qb, Q_RS, 128u), b_lo = umma_desc(sbase + L.q_lo + qb, Q_RS, 128u);
            umma_bf16(tmem_dw, a_lo, b_hi, idesc2, acc);
            umma_bf16(tmem_dw, a_hi, b_lo, idesc2, 1u);
            umma_bf16(tmem_dw, a_hi, b_hi, idesc2, 1u);
            acc = 1u;
          }
          umma_commit(&mi->bar);
        }
        mbar_wait(&mi->bar, phase);  // z and Q tiles are free again
        phase ^= 1;
        tc_fence_after();
      }
    } else {
      __syncthreads();
    }

    // ---- epilogue 2: fold dW (and the W_0 prior gradient) into d mu, d rho ----
    if (!a.w0_stopped) {
      float dW[32];
      if (lik) {
        tmem_ld16(tmem_dw + (lane_base << 16) + (uint32_t)(ch * 32), dW);
        tmem_ld16(tmem_dw + (lane_base << 16) + (uint32_t)(ch * 32 + 16), dW + 16);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) dW[e] = 0.f;
      }
      if (own) {
        float A = a.P_scalar, Bm = a.R_scalar * (float)a.K0;
        if (a.use_brd) {
          const float tau = a.smp_tau[(long long)s * a.K0 + k2], om = a.smp_om[(long long)s * a.K0 + k2];
          A = a.P_scalar / tau;
          Bm = a.R_scalar * (float)a.K0 / (tau * om);
        }
        const uint32_t woff = (uint32_t)(k2 >> 3) * W_RS + (uint32_t)(k2 & 7) * 16u;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const uint4 hv = *reinterpret_cast<const uint4*>(sm + L.w_hi + woff + (uint32_t)(ch * 4 + c8) * 128u);
          const uint4 lv = *reinterpret_cast<const uint4*>(sm + L.w_lo + woff + (uint32_t)(ch * 4 + c8) * 128u);
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
#pragma unroll
            for (int q2 = 0; q2 < 2; ++q2) {
              const int e = c8 * 8 + p * 2 + q2;
              const float W = q2 ? (bf16hi(hw[p]) + bf16hi(lw[p])) : (bf16lo(hw[p]) + bf16lo(lw[p]));
              if (W > V_LO_M && W < V_HI_M) {
                const float t = a.scale * dW[e] * W + a.global_weight * (A - 1.f - Bm * W);
                acc_mu[e] += t;
                acc_rho[e] = fmaf(t, __logf(W), acc_rho[e]);  // minus mu~ * acc_mu, folded in k_tc_w0_final
              }
            }
          }
        }
      }
    }
    // ---- per-(tile, s) side outputs: W_0 row sums, likelihood gradient of the gene scale ----
    tc_fence_before();
    __syncthreads();
    flush_rowstats(a, mi, ws.rowstats, s);
    if (lik && threadIdx.x < TG && g0 + (int)threadIdx.x < a.G && a.G_s) {
      const int g = g0 + threadIdx.x;
      for (int b = 0; b < a.nb; ++b) {
        const float sv = a.smp_s[((long long)s * a.nb + b) * a.G + g];
        a.G_s[((long long)s * a.nb + b) * a.G + g] += a.scale * (ws.colsumX[(long long)b * a.G + g] / sv - mi->colacc[b][threadIdx.x]);
        mi->colacc[b][threadIdx.x] = 0.f;
      }
    }
    __syncthreads();
    tc_fence_after();
  }

  // partial sums of this chunk: (n_chunks, 2, K0, G)
  if (own && !a.w0_stopped) {
    float* pm = ws.part + (((long long)chunk * 2 + 0) * a.K0 + k2) * a.G + g0 + ch * 32;
    float* pr = ws.part + (((long long)chunk * 2 + 1) * a.K0 + k2) * a.G + g0 + ch * 32;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      if (g0 + ch * 32 + e < a.G) {
        *reinterpret_cast<float4*>(pm + e) = make_float4(acc_mu[e], acc_mu[e + 1], acc_mu[e + 2], acc_mu[e + 3]);
        *reinterpret_cast<float4*>(pr + e) = make_float4(acc_rho[e], acc_rho[e + 1], acc_rho[e + 2], acc_rho[e + 3]);
      }
    }
  }
  double tot = block_sum_d((double)ll, mi->red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

#include "lik_tc_pipe.cuh"
#include "lik_tc_pair.cuh"
#include "lik_flat.cuh"

// ---- pre-kernels of the pipelined path -----------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tc_wprep(const LikTcArgs a, float* wmu, float* wsig) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long n = (long long)a.K0 * a.G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    wmu[i] = fminf(fmaxf(a.mu_W0[i], LOG_1E_10), LOG_1E6);
    wsig[i] = expf(fminf(fmaxf(a.rho_W0[i], LOG_1E_10), LOG_1E10));
  }
}
__global__ void __launch_bounds__(256) k_tc_rowprep(const LikTcArgs a, int* brow, long long* xoff) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.B) return;
  brow[j] = a.nb > 1 ? a.batch_id[a.idx[j]] : 0;
  xoff[j] = (a.idx_x ? a.idx_x[j] : (long long)j) * a.G;
}
// the same plus the row sums of the minibatch counts, one 128-thread CTA per row (a warp per row left 32 CTAs walking
// 20 KB rows serially: 10 us per launch; unified pair kernel)
__global__ void __launch_bounds__(128) k_tc_rowprep_sum(const LikTcArgs a, int* brow, long long* xoff, float* rowsum) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ float red[4];
  const int j = blockIdx.x, tid = threadIdx.x;
  if (j >= a.B) return;
  const long long off = (a.idx_x ? a.idx_x[j] : (long long)j) * a.G;
  const float* row = a.X + off;
  float acc = 0.f;
  for (int g = tid * 4; g < a.G; g += 512) {   // G is a multiple of 8, rows are 16 B aligned
    const float4 v = *reinterpret_cast<const float4*>(row + g);
    acc += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    rowsum[j] = (red[0] + red[1]) + (red[2] + red[3]);
    brow[j] = a.nb > 1 ? a.batch_id[a.idx[j]] : 0;
    xoff[j] = off;
  }
}
// minibatch counts -> image in which the epilogue's "lane = cell row, 16 B per lane" loads are contiguous:
// element (row = unit*128 + q*32 + lane, gene = tile*64 + v*4 + i) at ((((tile*units + unit)*4 + q)*16 + v)*32 + lane)*4 + i
__global__ void __launch_bounds__(256) k_tc_xprep(const LikTcArgs a, float* ximg, int n_tiles, int units) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long total = (long long)n_tiles * units * 4 * 16 * 32;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    // decode so that consecutive threads read consecutive genes of one row (coalesced reads)
    const int v = (int)(t % 16);
    long long rest = t / 16;
    const int tile = (int)(rest % n_tiles); rest /= n_tiles;
    const int lane = (int)(rest % 32); rest /= 32;
    const int q = (int)(rest % 4);
    const int unit = (int)(rest / 4);
    const int j = unit * UC + q * 32 + lane, g = tile * TG + v * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < a.B && g < a.G) x = *reinterpret_cast<const float4*>(a.X + (a.idx_x ? a.idx_x[j] : (long long)j) * a.G + g);
    *reinterpret_cast<float4*>(ximg + ((((((long long)tile * units + unit) * 4 + q) * 16 + v) * 32 + lane) * 4)) = x;
  }
}

// z0 samples (S,B,K0) fp32 -> per (s, 128-cell unit) image [hi plane | lo plane] in the canonical layout
__global__ void __launch_bounds__(256) k_tc_zprep(const LikTcArgs a, uint8_t* zimg, int KP, int units) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const int nchunk = KP / 8;
  const uint32_t z_rs = (uint32_t)nchunk * 128u, z_plane = 16u * z_rs;
  const long long per_unit = (long long)UC * nchunk, total = (long long)a.S * units * per_unit;
  const bool vec = (a.K0 & 3) == 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long su = t / per_unit;
    const int i = (int)(t - su * per_unit);
    const int s = (int)(su / units), un = (int)(su - (long long)s * units);
    const int r = i % UC, c = i / UC, j = un * UC + r, k = c * 8;
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
    uint8_t* dst = zimg + su * (2ll * z_plane) + (uint32_t)(r >> 3) * z_rs + (uint32_t)c * 128u + (uint32_t)(r & 7) * 16u;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + z_plane) = lo;
  }
}

// W_0 samples of every (MC sample, gene tile) -> tile images for the TMA-fed pipeline, at full occupancy.
// grid = (ceil(G/256), S), block = 256: thread = (row lane kk, 8-gene chunk), loop over the k0 row groups.
// Also: the per-row sums (sum_g log W, sum_g W) of the W_0 prior, and the gene-scale sample tile.
__global__ void __launch_bounds__(256) k_tc_wgen(const LikTcArgs a, const TcScratch ws, int KP, int n_tiles, uint32_t w_plane,
                                                 uint32_t w_unit) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ float rowacc[128][2 + MAX_NB];   // per k0 row: sum log W, sum W, sum_g W s_b[g] (the rank-1 vector W s_b)
  const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const int kk = tid & 7, cgl = tid >> 3;
  const int g = blockIdx.x * 256 + cgl * 8;
  const int tile = g / TG, cg = (g % TG) / 8;
  const int nkg = KP / 8;
  __shared__ __align__(16) float sgs[MAX_NB - 1][256];   // gene-scale samples of batches 1 .. nb-1 for the block's 256 genes
  for (int i = tid; i < 128 * (2 + MAX_NB); i += 256) (&rowacc[0][0])[i] = 0.f;
  for (int i = tid; i < (a.nb - 1) * 256; i += 256) {
    const int b = 1 + i / 256, gl = i & 255, gg = blockIdx.x * 256 + gl;
    sgs[b - 1][gl] = gg < a.G ? a.smp_s[((long long)s * a.nb + b) * a.G + gg] : 0.f;
  }
  __syncthreads();
  if (tile < n_tiles) {
    uint8_t* unit = ws.wimg + ((long long)s * n_tiles + tile) * w_unit;
    const bool gok = g < a.G;
    float4 sv0 = make_float4(0.f, 0.f, 0.f, 0.f), sv1 = sv0;
    if (gok) {
      const float* sp = a.smp_s + (long long)s * a.nb * a.G + g;
      sv0 = *reinterpret_cast<const float4*>(sp); sv1 = *reinterpret_cast<const float4*>(sp + 4);
    }
    for (int kg = 0; kg < nkg; ++kg) {
      const int k = kg * 8 + kk;
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = 0.f;
      float su = 0.f, sw = 0.f;
      if (k < a.K0 && gok) {
        const long long base = (long long)k * a.G + g;
        const float4 m0 = *reinterpret_cast<const float4*>(ws.wmu + base), m1 = *reinterpret_cast<const float4*>(ws.wmu + base + 4);
        const float4 s0 = *reinterpret_cast<const float4*>(ws.wsig + base), s1 = *reinterpret_cast<const float4*>(ws.wsig + base + 4);
        float4 e0, e1;
        if (a.eps_W0) {
          const float* ep = a.eps_W0 + (long long)s * a.K0 * a.G + base;
          e0 = *reinterpret_cast<const float4*>(ep); e1 = *reinterpret_cast<const float4*>(ep + 4);
        } else {
          e0 = philox_normal4(a.key, a.tag, (uint32_t)s, (uint32_t)(base >> 2));
          e1 = philox_normal4(a.key, a.tag, (uint32_t)s, (uint32_t)(base >> 2) + 1u);
        }
        const float mu[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
        const float sg[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        const float ep[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float u = fminf(fmaxf(fmaf(sg[e], ep[e], mu[e]), LOG_V_LO), LOG_V_HI);
          w[e] = __expf(u);
          su += u;
          sw += w[e];
        }
      }
      uint4 hi, lo;
      split8(w, hi, lo);
      const uint32_t off = (uint32_t)kg * W_RS + (uint32_t)cg * 128u + (uint32_t)kk * 16u;
      *reinterpret_cast<uint4*>(unit + off) = hi;
      *reinterpret_cast<uint4*>(unit + w_plane + off) = lo;
      // lanes {kk, kk+8, kk+16, kk+24} hold 4 chunks of row k
      su += __shfl_xor_sync(0xffffffffu, su, 8); sw += __shfl_xor_sync(0xffffffffu, sw, 8);
      su += __shfl_xor_sync(0xffffffffu, su, 16); sw += __shfl_xor_sync(0xffffffffu, sw, 16);
      if (lane < 8 && k < a.K0) {
        atomicAdd(&rowacc[k][0], su);
        atomicAdd(&rowacc[k][1], sw);
      }
      // (W s_b)[k] += sum over this thread's 8 genes (batch 0's gene-scale samples are held in registers, further
      // batches in shared memory)
      for (int b = 0; b < a.nb; ++b) {
        float d = 0.f;
        if (k < a.K0 && gok) {
          float4 t0 = sv0, t1 = sv1;
          if (b > 0) {
            const float* sp = &sgs[b - 1][cgl * 8];
            t0 = *reinterpret_cast<const float4*>(sp); t1 = *reinterpret_cast<const float4*>(sp + 4);
          }
          d = w[0] * t0.x + w[1] * t0.y + w[2] * t0.z + w[3] * t0.w + w[4] * t1.x + w[5] * t1.y + w[6] * t1.z + w[7] * t1.w;
        }
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        d += __shfl_xor_sync(0xffffffffu, d, 16);
        if (lane < 8 && k < a.K0) atomicAdd(&rowacc[k][2 + b], d);
      }
    }
  }
  // gene-scale sample tile(s) of this block's 256 genes
  for (int i = tid; i < a.nb * 256; i += 256) {
    const int b = i / 256, gl = i - b * 256, gg = blockIdx.x * 256 + gl;
    const int t2 = gg / TG;
    if (t2 < n_tiles) {
      float* st = reinterpret_cast<float*>(ws.wimg + ((long long)s * n_tiles + t2) * w_unit + 2 * w_plane);
      st[b * TG + (gg % TG)] = b > 0 ? sgs[b - 1][gl] : (gg < a.G ? a.smp_s[(long long)s * a.nb * a.G + gg] : 0.f);
    }
  }
  __syncthreads();
  for (int k = tid; k < a.K0; k += 256) {
    if (rowacc[k][0] != 0.f || rowacc[k][1] != 0.f) {
      atomicAdd(&ws.rowstats[((long long)s * a.K0 + k) * 2 + 0], rowacc[k][0]);
      atomicAdd(&ws.rowstats[((long long)s * a.K0 + k) * 2 + 1], rowacc[k][1]);
      for (int b = 0; b < a.nb; ++b) atomicAdd(&ws.Ws[((long long)s * a.nb + b) * a.K0 + k], rowacc[k][2 + b]);
    }
  }
}

// GLOBAL pass: likelihood gradient of the gene scale from the W_0 tile images,
//   G_s[s,b,g] += scale * ( colsumX[b][g] / s_{b,g} - sum_k ctilde[s][b][k] W^s[k][g] ).   grid = (n_tiles, S), block = 256
// The tile (KP x 64, bf16 hi + lo) is expanded to float in shared memory once; then thread = (gene, batch) runs the K0-long
// dot product with conflict-free reads (consecutive genes) and broadcast ctilde rows: any number of batches costs
// nb x 64 x K0 FMAs per tile instead of the shuffle reductions of the first version (0.53 ms per launch at 8 batches).
// With `loss_out` (unified pair kernel) it also adds sum_{b,g} colsumX_b[g] log s_bg, the gene-scale part of sum x log Lambda.
__global__ void __launch_bounds__(256) k_tc_colw(const LikTcArgs a, const TcScratch ws, int KP, int n_tiles, uint32_t w_plane,
                                                 uint32_t w_unit, double* loss_out = nullptr) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  extern __shared__ float csm[];
  __shared__ float red[8];
  float* Wt = csm;                         // [KP][65]
  float* ct_sm = csm + (size_t)KP * 65;    // [nb][KP]
  const int tile = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  const uint8_t* unit = ws.wimg + ((long long)s * n_tiles + tile) * w_unit;
  const float* ct = ws.ctilde + (long long)s * a.nb * a.K0;
  for (int i = tid; i < a.nb * KP; i += 256) {
    const int b = i / KP, k = i - b * KP;
    ct_sm[i] = k < a.K0 ? ct[b * a.K0 + k] : 0.f;
  }
  for (int c = tid; c < (KP / 8) * 64; c += 256) {   // chunk = (k0 row group kg, gene group cg, row kk): 8 genes of one k0 row
    const int kk = c & 7, cg = (c >> 3) & 7, kg = c >> 6;
    const uint32_t off = (uint32_t)kg * W_RS + (uint32_t)cg * 128u + (uint32_t)kk * 16u;
    const uint4 hv = *reinterpret_cast<const uint4*>(unit + off), lv = *reinterpret_cast<const uint4*>(unit + w_plane + off);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, lw[4] = {lv.x, lv.y, lv.z, lv.w};
    float* dst = Wt + (kg * 8 + kk) * 65 + cg * 8;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      dst[2 * p] = bf16lo(hw[p]) + bf16lo(lw[p]);
      dst[2 * p + 1] = bf16hi(hw[p]) + bf16hi(lw[p]);
    }
  }
  __syncthreads();
  const float* stile = reinterpret_cast<const float*>(unit + 2 * w_plane);
  float lgs = 0.f;
  for (int p = tid; p < a.nb * TG; p += 256) {
    const int gl = p & (TG - 1), b = p >> 6, g = tile * TG + gl;
    float acc = 0.f;
    const float* cb = ct_sm + b * KP;
#pragma unroll 4
    for (int k = 0; k < a.K0; ++k) acc = fmaf(cb[k], Wt[k * 65 + gl], acc);
    if (g < a.G) {
      const float cx = ws.colsumX[(long long)b * a.G + g], sv = stile[b * TG + gl];
      if (a.G_s) a.G_s[((long long)s * a.nb + b) * a.G + g] += a.scale * (cx / sv - acc);
      if (loss_out && cx > 0.f) lgs += cx * __logf(sv);
    }
  }
  if (loss_out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lgs += __shfl_xor_sync(0xffffffffu, lgs, o);
    if ((tid & 31) == 0) red[tid >> 5] = lgs;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      if (t != 0.f) atomicAdd(&loss_out[0], -(double)t * (double)a.scale / (double)a.S);
    }
  }
}

// ---- small companions ---------------------------------------------------------------------------------
// ctilde[s][b][k] = sum_{n in batch b} c_n^s z_nk^s ;  grid = S, block = 1024 (k = tid & 127, 8 row groups)
// With `loss_out` (GLOBAL pass of the unified pair kernel) it also adds the cell-scale parts of the log-likelihood of its
// sample:  sum_j rowsumX_j log c_j - sum_{b,k} ctilde_b[k] (W s_b)[k]   (= sum_n c_n z_n . (W s_b), the sum of all rates).
__global__ void __launch_bounds__(1024) k_tc_ctilde(const LikTcArgs a, float* ctilde, float* foldtab = nullptr, int KP = 0, int mat = 0,
                                                    const float* rowsumx = nullptr, const float* Ws = nullptr, double* loss_out = nullptr) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ float part[8][MAX_NB][128];
  __shared__ double red[32];
  const int s = blockIdx.x, k = threadIdx.x & 127, rg = threadIdx.x >> 7;
  double lacc = 0.0;
  float acc[MAX_NB];
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b) acc[b] = 0.f;
  if (k < a.K0) {
    // four rows per round: the loads of a round are independent, so one latency covers four of them (the loop was a
    // chain of 32 dependent-latency rounds on 100 CTAs: 20 us per launch)
    for (int j0 = rg; j0 < a.B; j0 += 32) {
      float v[4];
      int bj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 8 * u;
        const bool ok = j < a.B;
        v[u] = ok ? a.smp_c[(long long)s * a.B + j] * a.smp_z0[((long long)s * a.B + j) * a.K0 + k] : 0.f;
        bj[u] = (ok && a.nb > 1) ? a.batch_id[a.idx[j]] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int b = 0; b < MAX_NB; ++b) acc[b] += (b == bj[u]) ? v[u] : 0.f;
    }
  }
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b) part[rg][b][k] = acc[b];
  __syncthreads();
  for (int i = threadIdx.x; i < a.nb * a.K0; i += 1024) {
    const int b = i / a.K0, kk = i - b * a.K0;
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += part[r][b][kk];
    ctilde[((long long)s * a.nb + b) * a.K0 + kk] = t;
    if (foldtab) foldtab[((long long)s * (2 + a.nb) + 2 + b) * KP + kk] = a.scale * t;   // ctS_b[k]
    if (loss_out) lacc -= (double)(t * Ws[((long long)s * a.nb + b) * a.K0 + kk]);
  }
  if (loss_out) {
    for (int j = threadIdx.x; j < a.B; j += 1024) lacc += (double)(rowsumx[j] * __logf(a.smp_c[(long long)s * a.B + j]));
    const double tot = block_sum_d(lacc, red);
    if (threadIdx.x == 0) atomicAdd(&loss_out[0], -tot * (double)a.scale / (double)a.S);
  }
  // unified pair kernel: the per-(sample, k0) constants of the W_0 fold, rows [a1 | bm | ctS_0 .. ctS_{nb-1}] of KP floats.
  // scalar priors: a1 = gw (A - 1), bm = gw B;  matrix priors: a1 = 1 / tau, bm = 1 / (tau omega)
  if (foldtab) {
    float* t = foldtab + (long long)s * (2 + a.nb) * KP;
    for (int kk = threadIdx.x; kk < KP; kk += 1024) {
      float a1 = 0.f, bm = 0.f;
      if (kk < a.K0) {
        float tau = 1.f, om = 1.f;
        if (a.use_brd) { tau = a.smp_tau[(long long)s * a.K0 + kk]; om = a.smp_om[(long long)s * a.K0 + kk]; }
        if (mat) { a1 = 1.f / tau; bm = 1.f / (tau * om); }
        else {
          const float A = a.use_brd ? a.P_scalar / tau : a.P_scalar;
          a1 = a.global_weight * (A - 1.f);
          bm = a.global_weight * a.R_scalar * (float)a.K0 / (tau * om);
        }
      } else {
        for (int b = 0; b < a.nb; ++b) t[(2 + b) * KP + kk] = 0.f;
      }
      t[kk] = a1;
      t[KP + kk] = bm;
    }
  }
}

// colsumX[b][g] = sum_{n in batch b} x_ng ;  grid = ceil(G/64), block = 1024 (gene = tid & 63, 16 row groups)
__global__ void __launch_bounds__(1024) k_tc_colsum(const LikTcArgs a, float* colsumX) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ float part[16][MAX_NB][64];
  const int gl = threadIdx.x & 63, rg = threadIdx.x >> 6;
  const int g = blockIdx.x * 64 + gl;
  float acc[MAX_NB];
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b) acc[b] = 0.f;
  if (g < a.G) {
    for (int j = rg; j < a.B; j += 16) {
      const float x = a.X[(a.idx_x ? a.idx_x[j] : (long long)j) * a.G + g];
      const int bj = a.nb > 1 ? a.batch_id[a.idx[j]] : 0;
#pragma unroll
      for (int b = 0; b < MAX_NB; ++b) acc[b] += (b == bj) ? x : 0.f;
    }
  }
#pragma unroll
  for (int b = 0; b < MAX_NB; ++b) part[rg][b][gl] = acc[b];
  __syncthreads();
  for (int i = threadIdx.x; i < a.nb * 64; i += 1024) {
    const int b = i / 64, c = i - b * 64;
    if (blockIdx.x * 64 + c < a.G) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += part[r][b][c];
      colsumX[(long long)b * a.G + blockIdx.x * 64 + c] = t;
    }
  }
}

// G_z0[s][j][k] += scale * sum_r part[s][r][j][k]
__global__ void __launch_bounds__(256) k_tc_dz_reduce(const LikTcArgs a, const float* part, int NG) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  const long long per = (long long)a.B * a.K0, total = (long long)a.S * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per, r = i - s * per;
    float acc = 0.f;
    for (int q = 0; q < NG; ++q) acc += part[(s * NG + q) * per + r];
    a.G_z0[i] += a.scale * acc;
  }
}

// W_0 prior: value (both passes) and the brd / ard gradients (GLOBAL) from the row sums
// U = sum_g log W, Ws = sum_g W  (`_scdef.py:2014-2028`).  one thread per (s,k)
__global__ void __launch_bounds__(128) k_tc_w0_prior(const LikTcArgs a, const float* rowstats, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double val = 0.0;
  if (i < a.S * a.K0) {
    const float U = rowstats[2 * (long long)i], Ws = rowstats[2 * (long long)i + 1];
    const float Gn = (float)a.G;
    float A = a.P_scalar, Bm = a.R_scalar * (float)a.K0, tau = 1.f, om = 1.f;
    if (a.use_brd) {
      tau = a.smp_tau[i]; om = a.smp_om[i];
      A = a.P_scalar / tau;
      Bm = a.R_scalar * (float)a.K0 / (tau * om);
    }
    const float lB = logf(Bm);
    val = (double)((A - 1.f) * U - Bm * Ws) + (double)Gn * (double)(A * lB - lgammaf(A));
    if (a.use_brd && a.pass == SCDEF_PASS_GLOBAL) {
      const float dA = U - Gn * (digammaf(A) - lB);
      const float dB = -Ws + Gn * A / Bm;
      a.G_tau[i] += a.global_weight * (dA * (-A / tau) + dB * (-Bm / tau));
      a.G_om[i] += a.global_weight * dB * (-Bm / om);
    }
  }
  double tot = block_sum_d(val, red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[1], -tot * (double)a.global_weight / (double)a.S);
}

// W_0 prior with matrix-valued shape / rate (iscDEF gene-set priors, `_iscdef.py:442-450`): the Gamma parameters differ
// per (k, g), so the row sums are not enough.  One CTA per (64-gene tile, MC sample) walks the tile image
// (W = hi + lo, log W from it) and accumulates the prior value and, in the GLOBAL pass, the brd / ard gradients
// (same arithmetic as k_wprior mode 0/1).  The per-element W gradient of the prior is folded by the main kernel.
// fold_grad (GLOBAL pass of a rank without minibatch rows, W_0 not frozen): the main kernel does not run, so the
// prior's W gradient  t = gw (A - 1 - B W)  is accumulated here into chunk 0 of the partials k_tc_w0_final reduces.
__global__ void __launch_bounds__(256) k_tc_w0_prior_mat(const LikTcArgs a, const TcScratch ws, int KP, int n_tiles,
                                                         uint32_t w_plane, uint32_t w_unit, double* loss_out, int fold_grad) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  __shared__ float acc_t[128], acc_o[128];
  const int tile = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < 128; i += 256) { acc_t[i] = 0.f; acc_o[i] = 0.f; }
  __syncthreads();
  const uint8_t* unit = ws.wimg + ((long long)s * n_tiles + tile) * w_unit;
  const bool want_g = a.use_brd && a.pass == SCDEF_PASS_GLOBAL;
  double val = 0.0;
  // element e of the tile: k = e / 64 (row), gl = e % 64 (gene in tile); image offset (k/8)*W_RS + (gl/8)*128 + (k%8)*16 + (gl%8)*2
  for (int e = tid; e < a.K0 * TG; e += 256) {
    const int k = e >> 6, gl = e & 63, g = tile * TG + gl;
    if (g >= a.G) continue;
    const uint32_t off = (uint32_t)(k >> 3) * W_RS + (uint32_t)(gl >> 3) * 128u + (uint32_t)(k & 7) * 16u + (uint32_t)(gl & 7) * 2u;
    const float W = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(unit + off)) +
                    __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(unit + w_plane + off));
    const float lw = logf(W);
    const float P = a.P ? a.P[(long long)k * a.G + g] : a.P_scalar;
    const float R = a.R ? a.R[(long long)k * a.G + g] : a.R_scalar;
    float tau = 1.f, om = 1.f;
    if (a.use_brd) { tau = a.smp_tau[(long long)s * a.K0 + k]; om = a.smp_om[(long long)s * a.K0 + k]; }
    const float A = a.use_brd ? P / tau : P;
    const float Bm = R * (float)a.K0 / (tau * om);
    const float lB = logf(Bm);
    val += (double)((A - 1.f) * lw - Bm * W - lgammaf(A) + A * lB);
    if (want_g) {
      const float dA = lw - digammaf(A) + lB;
      const float dB = -W + A / Bm;
      atomicAdd(&acc_t[k], dA * (-A / tau) + dB * (-Bm / tau));
      atomicAdd(&acc_o[k], dB * (-Bm / om));
    }
    if (fold_grad && W > V_LO_M && W < V_HI_M) {
      const float t = a.global_weight * (A - 1.f - Bm * W);
      atomicAdd(&ws.part[(long long)k * a.G + g], t);
      atomicAdd(&ws.part[(long long)a.K0 * a.G + (long long)k * a.G + g], t * lw);
    }
  }
  __syncthreads();
  if (want_g)
    for (int k = tid; k < a.K0; k += 256) {
      atomicAdd(&a.G_tau[(long long)s * a.K0 + k], a.global_weight * acc_t[k]);
      atomicAdd(&a.G_om[(long long)s * a.K0 + k], a.global_weight * acc_o[k]);
    }
  double tot = block_sum_d(val, red);
  if (tid == 0) atomicAdd(&loss_out[1], -tot * (double)a.global_weight / (double)a.S);
}

// W_0: entropy value (both passes) and, in the GLOBAL pass, the final (mu, rho) gradients from the
// chunk partials.
__global__ void __launch_bounds__(256) k_tc_w0_final(const LikTcArgs a, const float* part, int n_chunks, double* loss_out) {
  pdl_grid_wait();   // programmatic dependent launch: inputs of the previous kernel are complete from here on
  __shared__ double red[32];
  const long long n = (long long)a.K0 * a.G;
  const float wH = a.anneal * a.global_weight, invS = 1.f / (float)a.S;
  double ent = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float mu = a.mu_W0[i], rho = a.rho_W0[i];
    ent += (double)(fminf(fmaxf(mu, LOG_1E_10), LOG_1E6) + fminf(fmaxf(rho, LOG_1E_10), LOG_1E10) + HALF_LOG_2PI_PLUS_HALF);
    if (a.pass == SCDEF_PASS_GLOBAL && a.gmu_W0) {
      float dmu = 0.f, drho = 0.f;
      if (!a.w0_stopped) {
        float gv = 0.f, glw = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          gv += part[((long long)c * 2 + 0) * n + i];
          glw += part[((long long)c * 2 + 1) * n + i];
        }
        // sum_s t_s * (sigma eps)_s  with  sigma eps = log W - mu~
        const float gve = glw - fminf(fmaxf(mu, LOG_1E_10), LOG_1E6) * gv;
        const bool in_mu = mu >= LOG_1E_10 && mu <= LOG_1E6, in_rho = rho >= LOG_1E_10 && rho <= LOG_1E10;
        dmu = in_mu ? -(gv * invS + wH) : 0.f;
        drho = in_rho ? -(gve * invS + wH) : 0.f;
      }
      a.gmu_W0[i] = dmu;
      a.grho_W0[i] = drho;
    }
  }
  double tot = block_sum_d(ent, red);
  if (threadIdx.x == 0) atomicAdd(&loss_out[1], -tot * (double)wH);
}

int round_up16(int k) { return (k + 15) / 16 * 16; }

// work decomposition: balance (items / 148 SMs rounded up) x (work per item)
Geom make_geom(int B, int S, int K0, int G) {
  Geom g;
  g.KP = round_up16(K0);
  g.n_tiles = (G + TG - 1) / TG;
  g.n_cblocks = B > 0 ? (B + CB - 1) / CB : 1;
  double best = 1e30;
  g.n_chunks = 1;
  for (int c = 1; c <= S; ++c) {
    const long long items = (long long)g.n_tiles * c;
    const double waves = (double)((items + SM_TARGET - 1) / SM_TARGET);
    const double cost = waves * (double)((S + c - 1) / c) * (1.0 + 0.002 * c) + 0.02 * c;  // mild penalty per partial buffer
    if (cost < best - 1e-9) { best = cost; g.n_chunks = c; }
  }
  best = 1e30;
  g.NG = 1;
  for (int ng = 1; ng <= g.n_tiles; ++ng) {
    const long long items = (long long)S * ng * g.n_cblocks;
    const double waves = (double)((items + SM_TARGET - 1) / SM_TARGET);
    const double cost = waves * ((double)((g.n_tiles + ng - 1) / ng) + 1.5) + 0.01 * ng;  // +1.5: per-item z load / dz drain
    if (cost < best - 1e-9) { best = cost; g.NG = ng; }
  }
  g.tiles_per_range = (g.n_tiles + g.NG - 1) / g.NG;
  g.NG = (g.n_tiles + g.tiles_per_range - 1) / g.tiles_per_range;
  return g;
}

size_t a256(size_t x) { return (x + 255) / 256 * 256; }

int g_sm_count = 148;   // set by lik_tc_launch / lik_tc_workspace_bytes callers through lik_tc_set_sm_count

FlatGeom make_flat_geom(bool genes, int B, int S, int K0, int G) {
  FlatGeom g;
  g.KP = round_up16(K0);
  const int n_gb = (G + 255) / 256, Bp = B > 0 ? B : 1;
  if (genes) { g.n_rblk = n_gb; g.n_pt = (Bp + 127) / 128; }
  else { g.n_rblk = (Bp + 255) / 256; g.n_pt = 2 * n_gb; }
  g.n_units = 2 * g.n_rblk;
  const long long T = (long long)g.n_rblk * S * g.n_pt;
  long long P = g_sm_count / 2;
  if (P < 1) P = 1;
  if (P > T) P = T > 0 ? T : 1;
  g.P = (int)P;
  int slots = 1;
  if (genes) {
    const long long per_rb = (long long)S * g.n_pt;
    for (int rb = 0; rb < g.n_rblk; ++rb) {
      const int n = flat_chunk_of((rb + 1) * per_rb - 1, T, g.P) - flat_chunk_of(rb * per_rb, T, g.P) + 1;
      if (n > slots) slots = n;
    }
  } else {
    for (long long sg = 0; sg < (long long)g.n_rblk * S; ++sg) {
      const int n = flat_chunk_of((sg + 1) * g.n_pt - 1, T, g.P) - flat_chunk_of(sg * g.n_pt, T, g.P) + 1;
      if (n > slots) slots = n;
    }
  }
  g.slots = slots;
  return g;
}

struct FlatSizes {
  size_t part, zimg, ximg, zimg_g, ximg_t, foldtab, Ws, rowsumx, wimg;
  int n_tiles;
};
FlatSizes flat_sizes(int B, int S, int K0, int G, int nb) {
  const FlatGeom gl = make_flat_geom(false, B, S, K0, G), gg = make_flat_geom(true, B, S, K0, G);
  const FlatSmem L = plan_flat_smem(gl.KP);
  const size_t Bp = (size_t)(B > 0 ? B : 1);
  FlatSizes z;
  z.n_tiles = 4 * ((G + 255) / 256);
  const size_t pl = (size_t)S * gl.slots * Bp * K0 * 4, pg = (size_t)gg.slots * 2 * K0 * G * 4;
  z.part = pl > pg ? pl : pg;
  z.zimg = (size_t)S * gl.n_units * 2 * L.a1_plane;
  z.ximg = (size_t)z.n_tiles * gl.n_units * 128 * 64 * 4;
  z.zimg_g = (size_t)S * gg.n_pt * 2 * L.stage;
  z.ximg_t = (size_t)(2 * gg.n_pt) * gg.n_units * 128 * 64 * 4;
  z.foldtab = (size_t)S * (2 + nb) * gl.KP * 4;
  z.Ws = (size_t)S * nb * K0 * 4;
  z.rowsumx = Bp * 4;
  z.wimg = (size_t)S * z.n_tiles * plan_pipe_smem(gl.KP).w_unit;
  return z;
}

}  // namespace

bool lik_tc_supported(int K0, int G, int nb) { return K0 >= 1 && K0 <= 112 && (G % 8) == 0 && nb >= 1 && nb <= MAX_NB; }

size_t lik_tc_workspace_bytes(int B, int S, int K0, int G, int nb) {
  const Geom g = make_geom(B, S, K0, G);
  const size_t Bp = (size_t)(B > 0 ? B : 1);
  const FlatSizes f = flat_sizes(B, S, K0, G, nb);
  size_t part = (size_t)S * g.NG * Bp * K0 * 4;                       // monolithic LOCAL
  const size_t part_global = (size_t)g.n_chunks * 2 * K0 * G * 4;     // monolithic GLOBAL
  if (part_global > part) part = part_global;
  if (f.part > part) part = f.part;
  return a256((size_t)S * K0 * 2 * 4) + a256((size_t)S * nb * K0 * 4) + a256((size_t)nb * G * 4) + a256(part) +
         2 * a256((size_t)K0 * G * 4) + a256(f.zimg) + a256(Bp * 4) + a256(Bp * 8) + a256(f.wimg) + a256(f.ximg) +
         a256(f.Ws) + a256(f.rowsumx) + a256(f.zimg_g) + a256(f.ximg_t) + a256(f.foldtab) + 1024;
}

void lik_tc_set_sm_count(int n) { if (n > 1) g_sm_count = n; }

void lik_tc_set_debug(float* p) { cudaMemcpyToSymbol(g_dbg_R, &p, sizeof(p)); }

// after a trapped launch of the pair kernel: {code of the wait that timed out, blockIdx.x, warp / sub-unit, parity}
static int* g_pair_stuck_host = nullptr;
static void pair_stuck_init() {
  if (g_pair_stuck_host) return;
  int* dev = nullptr;
  if (cudaHostAlloc((void**)&g_pair_stuck_host, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) { g_pair_stuck_host = nullptr; return; }
  for (int i = 0; i < 4; ++i) g_pair_stuck_host[i] = 0;
  cudaHostGetDevicePointer((void**)&dev, g_pair_stuck_host, 0);
  cudaMemcpyToSymbol(g_pair_stuck, &dev, sizeof(dev));
}
void lik_tc_pair_stuck(int* out) { for (int i = 0; i < 4; ++i) out[i] = g_pair_stuck_host ? g_pair_stuck_host[i] : 0; }

// host logic under test (tests/test_host.py): the work decomposition of the CTA-pair kernel for a problem shape.
// out = {KP, W_0 tile images per sample, LOCAL: n_rblk, n_pt, P, slots, GLOBAL: n_rblk, n_pt, P, slots}
void lik_tc_debug_geom(int B, int S, int K0, int G, int* out) {
  const FlatGeom l = make_flat_geom(false, B, S, K0, G), g = make_flat_geom(true, B, S, K0, G);
  const int v[10] = {l.KP, 4 * ((G + 255) / 256), l.n_rblk, l.n_pt, l.P, l.slots, g.n_rblk, g.n_pt, g.P, g.slots};
  for (int i = 0; i < 10; ++i) out[i] = v[i];
}
int lik_tc_debug_flat_chunk_of(long long f, long long T, int P) { return flat_chunk_of(f, T, P); }

void lik_tc_read_pair_timing(unsigned long long* out) {
#ifdef SCDEF_TC_TIMING
  cudaMemcpyFromSymbol(out, g_pair_timing, sizeof(unsigned long long) * 2 * 4 * 16);
#else
  for (int i = 0; i < 128; ++i) out[i] = 0;
#endif
}

int lik_tc_launch(const LikTcArgs& a_in, double* loss_out, int sm_count, cudaStream_t st) {
  lik_tc_set_sm_count(sm_count);
  LikTcArgs a = a_in;
  if (!lik_tc_supported(a.K0, a.G, a.nb)) return SCDEF_EUNSUPPORTED;
  const bool matrix_prior = a.P || a.R;   // iscDEF: Gamma shape / rate of W_0 per (k, g)
  if (a.B <= 0) a.lik_on = 0;
  const Geom g = make_geom(a.B, a.S, a.K0, a.G);
  const FlatSizes fz = flat_sizes(a.B, a.S, a.K0, a.G, a.nb);
  const size_t Bp = (size_t)(a.B > 0 ? a.B : 1);
  char* w = (char*)a.ws;
  TcScratch ws;
  ws.rowstats = (float*)w; w += a256((size_t)a.S * a.K0 * 2 * 4);
  ws.ctilde = (float*)w;   w += a256((size_t)a.S * a.nb * a.K0 * 4);
  ws.colsumX = (float*)w;  w += a256((size_t)a.nb * a.G * 4);
  ws.part = (float*)w;
  {
    size_t part = (size_t)a.S * g.NG * Bp * a.K0 * 4;
    const size_t part_global = (size_t)g.n_chunks * 2 * a.K0 * a.G * 4;
    if (part_global > part) part = part_global;
    if (fz.part > part) part = fz.part;
    w += a256(part);
  }
  ws.wmu = (float*)w;      w += a256((size_t)a.K0 * a.G * 4);
  ws.wsig = (float*)w;     w += a256((size_t)a.K0 * a.G * 4);
  ws.zimg = (uint8_t*)w;   w += a256(fz.zimg);
  ws.brow = (int*)w;       w += a256(Bp * 4);
  ws.xoff = (long long*)w; w += a256(Bp * 8);
  ws.wimg = (uint8_t*)w;   w += a256(fz.wimg);
  ws.ximg = (float*)w;     w += a256(fz.ximg);
  ws.Ws = (float*)w;       w += a256(fz.Ws);
  ws.rowsumx = (float*)w;  w += a256(fz.rowsumx);
  ws.zimg_g = (uint8_t*)w; w += a256(fz.zimg_g);
  ws.ximg_t = (float*)w;   w += a256(fz.ximg_t);
  ws.foldtab = (float*)w;  w += a256(fz.foldtab);
  if ((size_t)(w - (char*)a.ws) > a.ws_bytes) return SCDEF_EWORKSPACE;
  const Smem L = plan_smem(g.KP);
  const PipeSmem PL = plan_pipe_smem(g.KP);
  static const bool mono = getenv("SCDEF_TC_MONO") != nullptr;  // A/B switch: the monolithic kernels with the likelihood on
  const bool global_pass = a.pass == SCDEF_PASS_GLOBAL;
  const bool pair = a.lik_on && !mono;   // the CTA-pair kernel runs whenever the likelihood is on
  if (matrix_prior && mono) return SCDEF_EUNSUPPORTED;   // the monolithic kernels only know scalar W_0 priors
  cudaEvent_t ev0 = a.ev0, ev1 = a.ev1;
  a.ev0 = a.ev1 = nullptr;
  const bool reuse = pair && a.reuse_w0;  // W_0 images + row sums of the previous pass (same noise, same W_0) are still valid
  if (!reuse) {
    cudaMemsetAsync(ws.rowstats, 0, (size_t)a.S * a.K0 * 2 * 4, st);
    cudaMemsetAsync(ws.Ws, 0, (size_t)a.S * a.nb * a.K0 * 4, st);
  }
  const int n_tiles = fz.n_tiles;   // W_0 tile images per sample
  a.n_wimg = n_tiles;
  if (pair) {
    // ================= the CTA-pair kernel (lik_flat.cuh) =================
    const FlatGeom fg = make_flat_geom(global_pass, a.B, a.S, a.K0, a.G);
    const FlatSmem FL = plan_flat_smem(fg.KP);
    const int nblk = g_sm_count * 4;
    const long long T = (long long)fg.n_rblk * a.S * fg.n_pt;
    pair_stuck_init();
    SCDEF_LAUNCH(k_tc_rowprep_sum, a.B, 128, 0, st, a, ws.brow, ws.xoff, ws.rowsumx);
    if (!reuse) {
      SCDEF_LAUNCH(k_tc_wprep, g_sm_count * 2, 256, 0, st, a, ws.wmu, ws.wsig);
      SCDEF_LAUNCH(k_tc_wgen, dim3(n_tiles / 4, a.S), 256, 0, st, a, ws, fg.KP, n_tiles, PL.w_plane, PL.w_unit);
    }
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FL.total);
      if (ev0) cudaEventRecord(ev0, st);
      SCDEF_LAUNCH(kern, dim3(2 * fg.P), PIPE_NT, FL.total, st, a, fg, ws, loss_out);
      if (ev1 && ev0) cudaEventRecord(ev1, st);
    };
    const int rank1_blocks = (int)(((long long)a.S * a.B + 7) / 8);
    auto loss_terms = [&](bool gene_scale_term) {
      if (gene_scale_term) SCDEF_LAUNCH(k_tc_loss_gs, g_sm_count, 256, 0, st, a, ws.colsumX, loss_out);
      if (matrix_prior)
        SCDEF_LAUNCH(k_tc_w0_prior_mat, dim3(n_tiles, a.S), 256, 0, st, a, ws, fg.KP, n_tiles, PL.w_plane, PL.w_unit, loss_out, 0);
      else SCDEF_LAUNCH(k_tc_w0_prior, (a.S * a.K0 + 127) / 128, 128, 0, st, a, ws.rowstats, loss_out);
    };
    if (global_pass) {
      SCDEF_LAUNCH(k_tc_zprep_g, nblk, 256, 0, st, a, ws.zimg_g, fg.KP, fg.n_pt);
      SCDEF_LAUNCH(k_tc_xprep_t, nblk, 256, 0, st, a, ws.ximg_t, 2 * fg.n_pt, fg.n_units);
      SCDEF_LAUNCH(k_tc_ctilde, a.S, 1024, 0, st, a, ws.ctilde, ws.foldtab, fg.KP, matrix_prior ? 1 : 0, ws.rowsumx, ws.Ws, loss_out);
      SCDEF_LAUNCH(k_tc_colsum, (a.G + 63) / 64, 1024, 0, st, a, ws.colsumX);
      SCDEF_LAUNCH(k_tc_colw, dim3(n_tiles, a.S), 256, (size_t)(fg.KP * 65 + a.nb * fg.KP) * sizeof(float), st, a, ws, fg.KP, n_tiles, PL.w_plane, PL.w_unit, loss_out);
      if (matrix_prior) launch(k_lik_flat<true, true, true>);
      else launch(k_lik_flat<true, true, false>);
      loss_terms(false);   // the cell-scale terms of the loss came with k_tc_ctilde, the gene-scale term with k_tc_colw
      SCDEF_LAUNCH(k_tc_w0_final_flat, g_sm_count * 4, 256, 0, st, a, ws.part, T, fg.P, a.S * fg.n_pt, loss_out);
    } else {
      SCDEF_LAUNCH(k_tc_zprep, nblk, 256, 0, st, a, ws.zimg, fg.KP, fg.n_units);
      SCDEF_LAUNCH(k_tc_xprep, nblk, 256, 0, st, a, ws.ximg, n_tiles, fg.n_units);
      if (a.want_loss) launch(k_lik_flat<false, true, false>);
      else launch(k_lik_flat<false, false, false>);
      SCDEF_LAUNCH(k_tc_rank1, rank1_blocks, 256, 0, st, a, ws, ws.part, fg.slots, T, fg.P, fg.n_pt, 1, a.want_loss ? 1 : 0, loss_out);
      if (a.want_loss) {
        SCDEF_LAUNCH(k_tc_colsum, (a.G + 63) / 64, 1024, 0, st, a, ws.colsumX);
        loss_terms(true);
        LikTcArgs v = a;
        v.gmu_W0 = nullptr;   // LOCAL pass: the entropy value only
        SCDEF_LAUNCH(k_tc_w0_final_flat, g_sm_count * 4, 256, 0, st, v, nullptr, T, fg.P, a.S * fg.n_pt, loss_out);
      }
    }
    const cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) {
      fprintf(stderr, "scdef_b200: lik_tc_launch (pair kernel): %s (smem %u)\n", cudaGetErrorString(e2), FL.total);
      return SCDEF_ECUDA;
    }
    return SCDEF_OK;
  }

  // ================= likelihood off (layer 0 frozen, or a rank without minibatch rows), or SCDEF_TC_MONO =================
  // W_0 prior and entropy: value, and — GLOBAL pass of a rank without rows — their gradients.  The monolithic kernels
  // sample W_0 in the kernel; matrix-valued priors walk the tile images instead.
  const bool fold_prior_grad = matrix_prior && global_pass && !a.w0_stopped;
  if (matrix_prior) {
    SCDEF_LAUNCH(k_tc_wprep, g_sm_count * 2, 256, 0, st, a, ws.wmu, ws.wsig);
    SCDEF_LAUNCH(k_tc_wgen, dim3(n_tiles / 4, a.S), 256, 0, st, a, ws, g.KP, n_tiles, PL.w_plane, PL.w_unit);
    if (fold_prior_grad) cudaMemsetAsync(ws.part, 0, (size_t)g.n_chunks * 2 * a.K0 * a.G * 4, st);
  }
  if (global_pass) {
    if (a.lik_on) {
      SCDEF_LAUNCH(k_tc_ctilde, a.S, 1024, 0, st, a, ws.ctilde, (float*)nullptr, 0, 0, (const float*)nullptr, (const float*)nullptr, (double*)nullptr);
      SCDEF_LAUNCH(k_tc_colsum, (a.G + 63) / 64, 1024, 0, st, a, ws.colsumX);
    }
    if (!matrix_prior) {
      cudaFuncSetAttribute(k_lik_tc_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
      if (ev0 && a.lik_on) cudaEventRecord(ev0, st);
      SCDEF_LAUNCH(k_lik_tc_global, g.n_tiles * g.n_chunks, NT, L.total, st, a, g, ws, loss_out);
      if (ev1 && ev0 && a.lik_on) cudaEventRecord(ev1, st);
    }
  } else {
    if (a.lik_on) {
      cudaFuncSetAttribute(k_lik_tc_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
      if (ev0) cudaEventRecord(ev0, st);
      SCDEF_LAUNCH(k_lik_tc_local, dim3(a.S * g.NG, g.n_cblocks), NT, L.total, st, a, g, ws, loss_out);
      if (ev1 && ev0) cudaEventRecord(ev1, st);
      const long long total = (long long)a.S * a.B * a.K0;
      const int blocks = (int)((total + 255) / 256 < g_sm_count * 8 ? (total + 255) / 256 : g_sm_count * 8);
      SCDEF_LAUNCH(k_tc_dz_reduce, blocks, 256, 0, st, a, ws.part, g.NG);
    } else if (!matrix_prior) {
      // value only: the GLOBAL kernel without likelihood produces the W_0 row sums
      LikTcArgs v = a;
      v.w0_stopped = 1;
      cudaFuncSetAttribute(k_lik_tc_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
      SCDEF_LAUNCH(k_lik_tc_global, g.n_tiles * g.n_chunks, NT, L.total, st, v, g, ws, loss_out);
    }
  }
  // LOCAL pass with the loss discarded: the W_0 prior and entropy VALUES are all the kernels below would produce
  if (!global_pass && !a.want_loss) {
    const cudaError_t e0 = cudaGetLastError();
    if (e0 != cudaSuccess) {
      fprintf(stderr, "scdef_b200: lik_tc_launch: %s\n", cudaGetErrorString(e0));
      return SCDEF_ECUDA;
    }
    return SCDEF_OK;
  }
  if (matrix_prior)
    SCDEF_LAUNCH(k_tc_w0_prior_mat, dim3(n_tiles, a.S), 256, 0, st, a, ws, g.KP, n_tiles, PL.w_plane, PL.w_unit, loss_out, fold_prior_grad ? 1 : 0);
  else SCDEF_LAUNCH(k_tc_w0_prior, (a.S * a.K0 + 127) / 128, 128, 0, st, a, ws.rowstats, loss_out);
  SCDEF_LAUNCH(k_tc_w0_final, g_sm_count * 4, 256, 0, st, a, ws.part, g.n_chunks, loss_out);
  const cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    fprintf(stderr, "scdef_b200: lik_tc_launch: %s (smem %u)\n", cudaGetErrorString(err), L.total);
    return SCDEF_ECUDA;
  }
  return SCDEF_OK;
}

}  // namespace scdef
