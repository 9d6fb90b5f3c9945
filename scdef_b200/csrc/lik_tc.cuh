// tcgen05 / TMEM fused likelihood kernel — interface (implementation in lik_tc.cu).
#pragma once
#include "common.cuh"

namespace scdef {

struct LikTcArgs {
  const float* mu_W0;   // (K0,G)
  const float* rho_W0;  // (K0,G)
  const float* eps_W0;  // explicit noise (S,K0,G) or nullptr -> philox
  PhiloxKey key;
  NoiseTag tag;
  const float* smp_z0;  // (S,B,K0)
  const float* smp_c;   // (S,B)
  const float* smp_s;   // (S,nb,G)
  const float* smp_tau; // (S,K0)
  const float* smp_om;  // (S,K0)
  const float* X;
  const int64_t* idx_x;
  const int64_t* idx;
  const int32_t* batch_id;
  float* G_z0;  // (S,B,K0) LOCAL  (accumulated)
  float* G_c;   // (S,B)    LOCAL  (accumulated)
  float* G_s;   // (S,nb,G) GLOBAL (accumulated)
  float* G_tau; // (S,K0)   GLOBAL (accumulated)
  float* G_om;  // (S,K0)   GLOBAL (accumulated)
  float* gmu_W0;   // (K0,G) GLOBAL: final gradient (written)
  float* grho_W0;
  const float* P;  // W_0 prior matrices or nullptr
  const float* R;
  float P_scalar, R_scalar;
  int use_brd, B, G, K0, nb, pass, S, w0_stopped, lik_on, reuse_w0, want_loss;
  int n_wimg;       // W_0 tile images per MC sample in the workspace (set by lik_tc_launch)
  float scale, global_weight, anneal;
  void* ws;
  size_t ws_bytes;
  cudaEvent_t ev0, ev1;  // optional: recorded around the main fused kernel (measurement)
};

bool lik_tc_supported(int K0, int G, int nb);
size_t lik_tc_workspace_bytes(int B, int S, int K0, int G, int nb);
void lik_tc_set_debug(float* p);
void lik_tc_set_sm_count(int n);   // SMs of the device (decomposition of the persistent kernels); default 148
void lik_tc_read_pair_timing(unsigned long long* out);  // [rank][role][slot] of the CTA-pair kernel (-DSCDEF_TC_TIMING)
void lik_tc_pair_stuck(int* out);
void lik_tc_debug_geom(int B, int S, int K0, int G, int* out);   // work decomposition (host logic, CPU-testable)
int lik_tc_debug_flat_chunk_of(long long f, long long T, int P);  // bring-up: which wait of the CTA-pair kernel timed out (code, block, warp, parity)
int lik_tc_launch(const LikTcArgs& a, double* loss_out, int sm_count, cudaStream_t st);

}  // namespace scdef
