// tcgen05.mma issue-rate probe for the CTA PAIR (cta_group::2, sm_100a): cycles per MMA instruction of shape
// M=256 (128 rows per CTA) x N x K=16, bf16 -> f32, issued by one thread of the leader CTA, operands in each CTA's own
// shared memory (zero-filled: values do not matter for timing).  Companion of scripts/mma_probe.cu (single CTA).
// Every wait is bounded, so a wrong barrier protocol ends in a "timeout" line instead of a hung GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_probe2 scripts/mma_probe2.cu && scripts/mma_probe2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
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
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Cfg { int N, swz; };
constexpr int ITER = 2048;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_probe2(const Cfg* cfgs, int n_cfg, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t sbase = smem_u32(sm);
  const uint32_t A0 = sbase, B0 = sbase + 64 * 1024;
  uint32_t parity = 0;
  for (int c = 0; c < n_cfg; ++c) {
    const Cfg cf = cfgs[c];
    if (rank == 0 && threadIdx.x == 0) {
      const uint32_t id = idesc(256, cf.N);
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < ITER; i += 8) {
          const uint32_t ks = (uint32_t)((i >> 3) & 3);
          uint64_t a, b;
          if (cf.swz) { a = desc(A0 + ks * 32u, 16u, 1024u, 2u); b = desc(B0 + ks * 32u, 16u, 1024u, 2u); }
          else { a = desc(A0 + ks * 256u, 128u, 1024u, 0u); b = desc(B0 + ks * 256u, 128u, 1024u, 0u); }
#pragma unroll
          for (int q = 0; q < 8; ++q) mma2(tmem, a, b, id, (i + q) > 0);
        }
        commit2(&bar);
        long long spins = 0;
        while (!mbar_test(&bar, parity) && spins < 20000000LL) ++spins;
        const long long t1 = clock64();
        if (rep == 1) out[c] = spins >= 20000000LL ? ~0ull : (unsigned long long)(t1 - t0);
        parity ^= 1u;
      }
    } else if (rank == 1 && threadIdx.x == 0) {
      for (int rep = 0; rep < 2; ++rep) {   // the commit is multicast: the peer's barrier flips too
        long long spins = 0;
        while (!mbar_test(&bar, parity) && spins < 20000000LL) ++spins;
        parity ^= 1u;
      }
    }
    __syncthreads();
    cluster_sync();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const Cfg h[] = {{256, 1}, {256, 0}, {128, 1}, {128, 0}, {64, 1}, {64, 0}, {32, 1}};
  const int n = sizeof(h) / sizeof(h[0]);
  Cfg* d; unsigned long long* o;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, n * 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMemset(o, 0, n * 8);
  cudaFuncSetAttribute(k_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k_probe2<<<2, 128, 160 * 1024>>>(d, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long r[16];
  cudaMemcpy(r, o, n * 8, cudaMemcpyDeviceToHost);
  printf("cta_group::2, M=256 (2 x 128) K=16 bf16->f32, %d back-to-back MMAs, one issuing thread in the leader CTA\n", ITER);
  for (int i = 0; i < n; ++i) {
    if (r[i] == ~0ull) printf("N=%3d layout=%-6s : timeout (barrier never completed)\n", h[i].N, h[i].swz ? "swz128" : "none");
    else
      printf("N=%3d layout=%-6s : %7.1f cycles/MMA for 256x%dx16 (tensor floor for the pair: 128*N/512 = %.0f)\n", h[i].N,
             h[i].swz ? "swz128" : "none", (double)r[i] / ITER, h[i].N, 128.0 * h[i].N / 512.0);
  }
  return 0;
}
