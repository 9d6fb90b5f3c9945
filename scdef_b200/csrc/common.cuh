// Shared device helpers for the scdef_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/scdef_b200.h"

namespace scdef {

// ---------------------------------------------------------------------------------------
// Programmatic dependent launch.  A step is ~29 short kernels in one stream; with plain stream order every boundary
// pays "grid drained -> memory flushed -> next grid scheduled".  Every kernel here starts with `pdl_grid_wait()`
// (griddepcontrol.wait: returns once the grids this one depends on have completed and their writes are visible) and is
// launched with cudaLaunchAttributeProgrammaticStreamSerialization, so its blocks are scheduled while the previous
// grid's last blocks retire.  Nothing is read or written before the wait, so the semantics are those of stream order.
// SCDEF_NO_PDL=1 launches without the attribute (A/B).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_grid_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SCDEF_PDL_EARLY
  // A/B switch: release the dependents as soon as every block of this grid has STARTED (the trigger fires when all blocks
  // have called it, so no block of this grid is still waiting for a slot the dependents could take).  Measured at C5:
  // 1.357 ms per step against 1.348 without — the waiting blocks of the next kernels only take room; off by default.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
inline bool pdl_enabled() {
  static const bool on = getenv("SCDEF_NO_PDL") == nullptr;
  return on;
}
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}
#define SCDEF_LAUNCH(kern, grid, block, smem, st, ...) \
  (void)::scdef::launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(st), __VA_ARGS__)


constexpr float LOG_1E_10 = -23.025850929940457f;  // log(1e-10)  _scdef.py:1835,1837
constexpr float LOG_1E6 = 13.815510557964274f;     // log(1e6)    _scdef.py:1836
constexpr float LOG_1E10 = 23.025850929940457f;    // log(1e10)   _scdef.py:1838
constexpr float V_LO = 1e-15f, V_HI = 1e15f;       // jax_utils.py:8
constexpr float HALF_LOG_2PI_PLUS_HALF = 1.4189385332046727f;

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator; one call -> two N(0,1) (Box-Muller).
// ---------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

struct PhiloxKey {
  uint32_t k0, k1;
};

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, PhiloxKey k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.k0, lo1, hi0 ^ c.w ^ k.k1, lo0);
    k.k0 += W0;
    k.k1 += W1;
  }
  return c;
}

// Identifies the noise stream of one variational block.
struct NoiseTag {
  uint32_t a, b;
};

__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Four standard normals for the elements (4*quad .. 4*quad+3) of block `tag`, MC sample `s`:
// one Philox call, two Box-Muller pairs.
__device__ __forceinline__ float4 philox_normal4(PhiloxKey key, NoiseTag tag, uint32_t s, uint32_t quad) {
  uint4 r = philox4x32_10(make_uint4(quad, tag.a, tag.b, s), key);
  const float k = 2.3283064365386963e-10f;  // 2^-32; u in (0,1): (x + 0.5) * 2^-32
  float u1 = fminf(((float)r.x + 0.5f) * k, 0.99999994f);
  float u2 = ((float)r.y + 0.5f) * k;
  float u3 = fminf(((float)r.z + 0.5f) * k, 0.99999994f);
  float u4 = ((float)r.w + 0.5f) * k;
  float ra = fast_sqrt(-2.0f * __logf(u1)), rb = fast_sqrt(-2.0f * __logf(u3));
  float sa, ca, sb, cb;
  __sincosf(6.283185307179586f * u2, &sa, &ca);
  __sincosf(6.283185307179586f * u4, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__device__ __forceinline__ float philox_normal(PhiloxKey key, NoiseTag tag, uint32_t s, uint32_t elem) {
  float4 n = philox_normal4(key, tag, s, elem >> 2);
  uint32_t c = elem & 3u;
  return c == 0 ? n.x : (c == 1 ? n.y : (c == 2 ? n.z : n.w));
}

// ---------------------------------------------------------------------------------------
// Special functions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float digammaf(float x) {
  float acc = 0.f;
  while (x < 6.f) {
    acc -= 1.f / x;
    x += 1.f;
  }
  float inv = 1.f / x, inv2 = inv * inv;
  // ln x - 1/(2x) - 1/(12x^2) + 1/(120x^4) - 1/(252x^6)
  return acc + logf(x) - 0.5f * inv -
         inv2 * (0.083333333333f - inv2 * (0.0083333333333f - inv2 * 0.003968253968f));
}

// Gamma(a, b) log-density at x with log x = lx supplied (jax_utils.py:35-46)
__device__ __forceinline__ float gamma_lp(float x, float lx, float a, float b, float lgam_a, float log_b) {
  return (a - 1.f) * lx - b * x - lgam_a + a * log_b;
}

// ---------------------------------------------------------------------------------------
// Reductions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide double sum, result valid in thread 0.  `red` = shared double[32].
__device__ __forceinline__ double block_sum_d(double v, double* red) {
  v = warp_sum_d(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum_d(r);
  }
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------
// One LogNormal variational block as the generic kernels see it.
// ---------------------------------------------------------------------------------------
struct BlockDesc {
  const float* mu;   // parameter rows: element (r, c) at mu[row(r) * ld + col0 + c]
  const float* rho;
  const int64_t* gather;  // row(r) = gather[r] (local blocks) or r
  const float* eps;       // explicit noise (S, rows, eps_ld) at column eps_col0, or nullptr
  const float* fixed;     // sscDEF: constant sample rows (n_cells, cols) gathered, or nullptr
  float* smp;             // (S, rows, cols) samples v
  float* G;               // (S, rows, cols) dELBO_s/dv
  float* gmu;             // gradient outputs: element (r, c) at gmu[r * g_ld + g_col0 + c]
  float* grho;
  int rows, cols, ld, col0, eps_ld, eps_col0, g_ld, g_col0;
  NoiseTag tag;
  long long noise_e0;  // flat element offset of row 0 in the philox stream of this block (data-parallel row offset)
  float mu_hi;      // LOG_1E6, or +inf for the gene scale (_scdef.py:1872)
  float ent_weight; // w_H * scale_H * (block weight), 0 for the pinned sscDEF layer
  float g_weight;   // multiplies the sample-gradient part (1; global_weight handled upstream)
  int stopped;      // stop_gradient on (mu, sigma): both gradients are exactly 0
  int active;       // 0: block not part of the model (e.g. brd without use_brd) -> zero grads
};

constexpr int MAX_BLOCKS = 5 + 2 * SCDEF_MAX_LAYERS;

struct BlockTable {
  BlockDesc b[MAX_BLOCKS];
  int n;
};

}  // namespace scdef
