// CTA-pair machinery shared by the fused likelihood kernel (lik_flat.cuh; included by lik_tc.cu inside its anonymous
// namespace after lik_tc_pipe.cuh): cluster-scope PTX wrappers, bounded barrier waits, the cta_group::2 MMA / commit
// wrappers, the barrier block of a CTA, the chunking of the flattened work.
//
// Why pairs (profiles/r01_mma_probe.txt): a cta_group::1 tcgen05.mma occupies the tensor pipe for ~86 cycles whatever
// N <= 128 is; the pair's 256 x 128 x 16 MMA costs 64.
//
// Barrier protocol (modelled in tests/test_pair_protocol.py).  Only the leader (rank 0) issues pair MMAs.
// tcgen05.commit is multicast to both CTAs, so "tensor core is done with ..." barriers are waited locally in each CTA.
// "... is ready for the tensor core" barriers live in the LEADER's shared memory and are arrived on remotely
// (mapa + mbarrier.arrive on the shared::cluster address) by one elected lane per epilogue warp of both CTAs, or by the
// peer's relay thread (the peer's MMA warp forwards "my TMA landed": pz_full, pw_full).
#pragma once

constexpr long long PAIR_DEADLINE = 2000000000ll;  // cycles (~1 s) before a stuck wait traps

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
#ifdef SCDEF_PAIR_SPIN
  // A/B switch: poll with the non-blocking test_wait instead of the suspending try_wait
  if (mbar_test_remote(bar, parity)) return;
  const long long ts = clock64();
  while (!mbar_test_remote(bar, parity)) {
    if (clock64() - ts > PAIR_DEADLINE) pair_stuck(code, (int)blockIdx.x, (int)(threadIdx.x >> 5), (int)parity);
  }
  return;
#endif
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

// tensor-memory columns of the LOCAL orientation: R[2] at 0 / 128; dz0[2] at 256 / 384
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
  uint32_t tmem_base, pad;
  double red[32];
};
static_assert(sizeof(FlatMisc) <= 1024, "FlatMisc must fit the misc block of plan_flat_smem");

// chunk that contains flat index f when [0, T) is cut at floor(c * T / P)
__host__ __device__ inline int flat_chunk_of(long long f, long long T, int P) { return (int)(((f + 1) * P - 1) / T); }
