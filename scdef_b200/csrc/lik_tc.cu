// tcgen05 / TMEM fused likelihood kernel (placeholder until the kernel lands).
#include "lik_tc.cuh"

namespace scdef {
bool lik_tc_supported(int, int, int) { return false; }
size_t lik_tc_workspace_bytes(int, int, int, int, int) { return 0; }
int lik_tc_launch(const LikTcArgs&, double*, int, cudaStream_t) { return SCDEF_EUNSUPPORTED; }
}  // namespace scdef
