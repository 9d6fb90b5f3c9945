// Semantics probe for the GLOBAL CTA-pair kernel (sm_100a), three questions, each checked NUMERICALLY:
//   (1) cta_group::2 MMA (M = 256, N = 128): which CTA supplies which half of N, and which rows land in which CTA's TMEM?
//       (k_lik_pair_* assumes: rank 0 supplies columns [0, N/2), each CTA gets its own 128 rows.)
//   (2) MIXING: inside a cluster-(2,1,1) kernel whose TMEM was allocated with cta_group::2, may each CTA also issue
//       cta_group::1 MMAs (own operands -> own TMEM) and commit them to its own barrier?  ptxas accepts it; if the
//       hardware does too, the GLOBAL pass can keep its second GEMM (contraction over the CTA's own cells) per CTA
//       while the rate GEMM runs on the pair.
//   (3) A operand from tensor memory (pair MMA): packing of bf16 pairs in a 32-bit column written with tcgen05.st.
// Operands are small integers (exact in bf16): A[m][k] = a_r(m), B[n][k] = b_r(n), so D[m][n] = 16 a(m) b(n).
// Every wait is bounded: a wrong protocol prints "timeout" instead of hanging the device.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_probe3 scripts/mma_probe3.cu && scripts/mma_probe3
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {   // 8 x 16 B core matrices, K = 16: LBO 128, SBO 256
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(256u >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool wait_bounded(uint64_t* bar, uint32_t parity) {
  for (long long i = 0; i < 20000000LL; ++i) if (mbar_test(bar, parity)) return true;
  return false;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float a_val(int rank, int m) { return (float)(rank ? 2 * (m % 5 + 1) : (m % 7 + 1)); }
__device__ __forceinline__ float b_val(int rank, int n) { return (float)(rank ? (n % 4 + 2) : (n % 3 + 1)); }

// out[rank][test][0] = status (1 ok, 2 mismatch, 3 timeout), [1] = number of mismatching elements, [2..3] = first bad (m, n)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_probe3(int* out, int do_mixed, int do_tmem_a) {
  __shared__ __align__(1024) uint8_t sA[128 * 32];   // A: 128 rows x 16 k, bf16
  __shared__ __align__(1024) uint8_t sB[128 * 32];   // B: up to 128 rows (n) x 16 k
  __shared__ uint64_t bar_pair, bar_own;
  __shared__ uint32_t tmem_base;
  __shared__ int bad_count, bad_m, bad_n;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * 16; i += 128) {
    const int row = i / 16, k = i % 16;
    const uint32_t off = (uint32_t)(row >> 3) * 256u + (uint32_t)(k >> 3) * 128u + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(a_val((int)rank, row));
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(b_val((int)rank, row));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_pair)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_own)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint64_t da = desc_kmajor(smem_u32(sA)), db = desc_kmajor(smem_u32(sB));
  auto check = [&](int test, uint32_t col0, int ncols, bool timed_out, auto expect) {
    if (tid == 0) { bad_count = 0; bad_m = -1; bad_n = -1; }
    __syncthreads();
    if (!timed_out) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = warp * 32 + lane;
      for (int c = 0; c < ncols; c += 16) {
        float v[16];
        ld16(tmem + ((uint32_t)(warp * 32) << 16) + col0 + (uint32_t)c, v);
        for (int i = 0; i < 16; ++i)
          if (v[i] != expect(m, c + i)) { if (atomicAdd(&bad_count, 1) == 0) { bad_m = m; bad_n = c + i; } }
      }
    }
    __syncthreads();
    if (tid == 0) {
      int* o = out + ((int)rank * 3 + test) * 4;
      o[0] = timed_out ? 3 : (bad_count ? 2 : 1); o[1] = bad_count; o[2] = bad_m; o[3] = bad_n;
    }
    __syncthreads();
  };

  // ---- (1) pair MMA: M = 256, N = 128 (each CTA supplies 64 rows of B), D at columns [0, 128) ----
  if (rank == 0 && tid == 0) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem), "l"(da), "l"(db), "r"(idesc(256, 128)), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar_pair)), "h"((uint16_t)3) : "memory");
  }
  bool to = !wait_bounded(&bar_pair, 0);
  check(0, 0u, 128, to, [&](int m, int n) { return 16.f * a_val((int)rank, m) * b_val(n / 64, n % 64); });

  // ---- (2) mixed: each CTA issues a cta_group::1 MMA (M = 128, N = 64) on its own operands, D at columns [128, 192) ----
  if (do_mixed) {
    if (tid == 0) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tmem + 128u), "l"(da), "l"(db), "r"(idesc(128, 64)), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_own)) : "memory");
    }
    to = !wait_bounded(&bar_own, 0);
    check(1, 128u, 64, to, [&](int m, int n) { return 16.f * a_val((int)rank, m) * b_val((int)rank, n); });
  }

  // ---- (3) pair MMA with A from tensor memory: each thread stores its row's 16 k as 8 packed bf16x2 columns at [256, 264) ----
  if (do_tmem_a) {
    {
      const float av = a_val((int)rank, warp * 32 + lane);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(av, av);
      const uint32_t w = *reinterpret_cast<uint32_t*>(&p2);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + 256u), "r"(w), "r"(w), "r"(w), "r"(w), "r"(w), "r"(w), "r"(w), "r"(w) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (rank == 0 && tid == 0) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                   ::"r"(tmem + 320u), "r"(tmem + 256u), "l"(db), "r"(idesc(256, 128)), "r"(0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                   ::"r"(smem_u32(&bar_pair)), "h"((uint16_t)3) : "memory");
    }
    to = !wait_bounded(&bar_pair, 1);
    check(2, 320u, 128, to, [&](int m, int n) { return 16.f * a_val((int)rank, m) * b_val(n / 64, n % 64); });
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static int run(int do_mixed, int do_tmem_a, const char* what) {
  int* o;
  cudaMalloc(&o, 2 * 3 * 4 * sizeof(int));
  cudaMemset(o, 0, 2 * 3 * 4 * sizeof(int));
  k_probe3<<<2, 128>>>(o, do_mixed, do_tmem_a);
  cudaError_t e = cudaDeviceSynchronize();
  printf("== %s\n", what);
  if (e != cudaSuccess) { printf("   CUDA error: %s (the context is gone: later tests run in a fresh process)\n", cudaGetErrorString(e)); return 1; }
  int h[24];
  cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[3] = {"pair MMA (N split, row ownership)", "cta_group::1 MMA inside the pair kernel", "pair MMA, A from tensor memory"};
  const char* st[4] = {"not run", "OK", "MISMATCH", "TIMEOUT"};
  for (int r = 0; r < 2; ++r)
    for (int t = 0; t < 3; ++t) {
      const int* p = h + (r * 3 + t) * 4;
      if (p[0]) printf("   rank %d  %-42s %s  (bad elements %d, first at m=%d n=%d)\n", r, names[t], st[p[0]], p[1], p[2], p[3]);
    }
  cudaFree(o);
  return 0;
}

int main(int argc, char** argv) {
  // one question per process run is safest (an illegal instruction kills the context): `mma_probe3 1`, `2`, `3`; no argument = all
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0 || which == 1) run(0, 0, "(1) pair MMA only");
  if (which == 0 || which == 2) run(1, 0, "(2) pair MMA + cta_group::1 MMA in the same kernel");
  if (which == 0 || which == 3) run(0, 1, "(3) pair MMA + pair MMA with A from tensor memory");
  return 0;
}
