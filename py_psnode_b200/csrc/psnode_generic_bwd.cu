// psnode_generic_bwd.cu -- discrete-adjoint reverse sweep (placeholder until the backward kernel lands).
#include "psnode_internal.cuh"
int64_t psn_generic_backward_workspace(const psnode_problem*, const psnode_adjoint*) { return 0; }
int psn_generic_backward(const psnode_problem*, const psnode_adjoint*, void*, int64_t, cudaStream_t) { return PSNODE_EUNSUPPORTED; }
