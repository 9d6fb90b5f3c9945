// tcgen05.mma issue-rate probe (sm_100a): cycles per MMA instruction (M=128, K=16, bf16 -> f32) as a function
// of N, the shared-memory layout (no swizzle = the layout the likelihood kernel uses today, 128B swizzle) and
// the A-operand source (shared memory descriptor vs tensor memory).  The operand VALUES are irrelevant for the
// timing, so the shared memory is simply zero-filled.  One CTA, one issuing thread, ITER back-to-back MMAs that
// accumulate into one TMEM tile, then one commit + mbarrier wait; reports (t_end - t_start) / ITER.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_probe scripts/mma_probe.cu && scripts/mma_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(id), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
               "r"(a_tmem), "l"(b), "r"(id), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}

struct Cfg { int N, swz, a_tmem, b_mn, a_mn, n_acc, unroll, M; };  // n_acc: independent accumulator tiles used round-robin
constexpr int ITER = 2048;

__global__ void __launch_bounds__(128, 1) k_probe(const Cfg* cfgs, int n_cfg, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  const uint32_t sbase = smem_u32(sm);
  const uint32_t A0 = sbase, B0 = sbase + 64 * 1024;   // A: up to 128 rows x 256 B; B: up to 256 rows x 256 B
  uint32_t parity = 0;
  if (threadIdx.x == 0) {
    for (int c = 0; c < n_cfg; ++c) {
      const Cfg cf = cfgs[c];
      const uint32_t id = idesc(cf.M ? cf.M : 128, cf.N, cf.a_mn, cf.b_mn);
      // K extent of the staged tiles: 64 elements = 4 MMA K-steps (cycled), like one 64-gene / 64-factor slab
      for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
        const long long t0 = clock64();
        for (int i = 0; i < ITER; ++i) {
          const uint32_t ks = (uint32_t)(i & 3);
          uint64_t a, b;
          if (cf.swz) {   // 128B swizzle: rows of 128 B, 8-row atoms of 1024 B; K-step = +32 B inside the row
            a = desc(A0 + ks * 32u, 16u, 1024u, 2u);
            b = desc(B0 + ks * 32u, 16u, 1024u, 2u);
          } else {        // no swizzle: core matrices of 8 rows x 16 B; K-adjacent cores 128 B apart, row groups 1024 B apart
            a = cf.a_mn ? desc(A0 + ks * 2048u, 1024u, 128u, 0u) : desc(A0 + ks * 256u, 128u, 1024u, 0u);
            b = cf.b_mn ? desc(B0 + ks * 2048u, 1024u, 128u, 0u) : desc(B0 + ks * 256u, 128u, 1024u, 0u);
          }
          const uint32_t dt = tmem + (uint32_t)(i % cf.n_acc) * (uint32_t)cf.N;
          if (cf.unroll) {   // 8 MMAs per loop trip with the same descriptors: isolates the issue-loop overhead
#pragma unroll
            // nothing but the MMAs inside the unrolled group (an earlier version computed `q % n_acc` per MMA and
            // measured that integer division instead of the tensor core: ~110 cycles "whatever N")
            if (cf.a_tmem) {
              const uint32_t at = tmem + 384u + ks * 8u;
              for (int q = 0; q < 8; ++q) mma_ts(tmem, at, b, id, 1u);
            } else if (cf.n_acc == 2) {
              const uint32_t d1 = tmem + (uint32_t)cf.N;
              for (int q = 0; q < 4; ++q) { mma_ss(tmem, a, b, id, 1u); mma_ss(d1, a, b, id, 1u); }
            } else {
              for (int q = 0; q < 8; ++q) mma_ss(tmem, a, b, id, 1u);
            }
            i += 7;
            continue;
          }
          if (cf.a_tmem) mma_ts(dt, tmem + 384u + ks * 8u, b, id, i >= cf.n_acc);
          else mma_ss(dt, a, b, id, i >= cf.n_acc);
        }
        commit(&bar);
        mbar_wait(&bar, parity);
        parity ^= 1u;
        const long long t1 = clock64();
        if (rep == 1) out[c] = (unsigned long long)(t1 - t0);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const Cfg h[] = {
      {16, 0, 0, 0, 0, 1, 1, 128},  {32, 0, 0, 0, 0, 1, 1, 128},  {64, 0, 0, 0, 0, 1, 1, 128},  {64, 1, 0, 0, 0, 1, 1, 128},
      {112, 0, 0, 0, 0, 1, 1, 128}, {128, 0, 0, 0, 0, 1, 1, 128}, {256, 0, 0, 0, 0, 1, 1, 128}, {256, 1, 0, 0, 0, 1, 1, 128},
      {64, 0, 0, 1, 0, 1, 1, 128},  {64, 0, 0, 1, 1, 1, 1, 128},  {64, 0, 0, 0, 0, 2, 1, 128},
      {64, 0, 1, 0, 0, 1, 1, 128},  {112, 0, 1, 0, 0, 1, 1, 128}, {256, 0, 1, 0, 0, 1, 1, 128},
      {64, 0, 0, 0, 0, 1, 1, 64},   {256, 0, 0, 0, 0, 1, 1, 64},
      {64, 0, 0, 0, 0, 1, 0, 128},
  };



  const int n = sizeof(h) / sizeof(h[0]);
  Cfg* d; unsigned long long* o;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, n * 8);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  k_probe<<<1, 128, 160 * 1024>>>(d, n, o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  unsigned long long r[64];
  cudaMemcpy(r, o, n * 8, cudaMemcpyDeviceToHost);
  printf("M=128 K=16 bf16->f32, %d back-to-back MMAs, one issuing thread\n", ITER);
  for (int i = 0; i < n; ++i)
    printf("N=%3d  layout=%-8s A=%-4s A-major=%s B-major=%s accs=%d unrolled=%d M=%d : %7.1f cycles/MMA   (floor 128*N/256 = %.0f)\n", h[i].N,
           h[i].swz ? "swz128" : "none", h[i].a_tmem ? "tmem" : "smem", h[i].a_mn ? "MN" : "K ", h[i].b_mn ? "MN" : "K ",
           h[i].n_acc, h[i].unroll, h[i].M, (double)r[i] / ITER, 128.0 * h[i].N / 256.0);
  return 0;
}
