// psnode_fused_fwd.cu -- register-resident-weight forward kernel (placeholder until the fused path lands).
#include "psnode_internal.cuh"
bool psn_fused_supports(const psnode_problem*) { return false; }
int64_t psn_fused_forward_workspace(const psnode_problem*) { return 0; }
int psn_fused_forward(const psnode_problem*, void*, int64_t, cudaStream_t) { return PSNODE_EUNSUPPORTED; }
