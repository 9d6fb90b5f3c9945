// Shared pieces of the warp-specialised likelihood kernels (included by lik_tc.cu inside its anonymous namespace;
// uses its PTX wrappers and layouts): thread roles, the layout of a W_0 tile image, mbarrier / TMA wrappers.
// (The single-CTA pipelined kernel of round 1 that lived here was replaced by the CTA-pair kernel of lik_flat.cuh.)
//
// One CTA = 18 warps: 16 EPILOGUE warps, one MMA-issuing warp, one COPY (TMA) warp.
#pragma once

constexpr int PIPE_NT = 576;           // 18 warps
constexpr int E_THREADS = 512, E_WARPS = 16, MMA_WARP = 16, COPY_WARP = 17;

// layout of one W_0 tile image in global memory (k_tc_wgen): 64 genes x KP k0 rows as [hi plane | lo plane | gene-scale tile]
struct PipeSmem {
  uint32_t z_rs, z_plane, w_plane, q_plane;
  uint32_t w_unit;                       // bytes of one W image unit: hi | lo | gene-scale tile (MAX_NB x TG floats)
};
__host__ __device__ inline PipeSmem plan_pipe_smem(int KP) {
  PipeSmem s;
  s.z_rs = (uint32_t)(KP / 8) * 128u;
  s.z_plane = 16u * s.z_rs;                 // 128 cells
  s.w_plane = (uint32_t)(KP / 8) * W_RS;    // KP rows x 64 genes
  s.q_plane = 16u * Q_RS;                   // 128 cells x 64 genes
  s.w_unit = 2 * s.w_plane + (uint32_t)(MAX_NB * TG * 4);
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

