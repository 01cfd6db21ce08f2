// psnode_internal.cuh -- declarations shared by the translation units of libpsnode_b200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/psnode_b200.h"
#include "psnode_math.cuh"

// ---- launch bookkeeping (psnode_api.cu) -------------------------------------------------------
void psn_count_launch(const char* kernel_name);
int psn_cuda_fail(cudaError_t e, const char* where);
#define PSN_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return psn_cuda_fail(e__, #call);   \
    } while (0)

// ---- packed weights ---------------------------------------------------------------------------
// Every Linear layer is re-laid out once per call into the workspace as rows of `kpad` floats
// (kpad = in_dim rounded up to a multiple of 4 with kpad/4 odd, tail zero-filled): rows are 16-byte aligned for
// 128-bit loads and 8 consecutive rows fall into 8 disjoint bank quads, so a quarter-warp LDS.128 over consecutive
// neurons is conflict free.
struct PsnPackedNet {
    int n_layers;
    int in_dim[PSNODE_MAX_LAYERS];
    int out_dim[PSNODE_MAX_LAYERS];
    int kpad[PSNODE_MAX_LAYERS];
    int w_off[PSNODE_MAX_LAYERS];     // float offset of the layer's packed weights inside the packed buffer
    int b_off[PSNODE_MAX_LAYERS];     // float offset of its bias
    int smem_off[PSNODE_MAX_LAYERS];  // float offset inside the shared-memory weight region, or -1: stream from global/L2
    int total;                        // floats used by this net in the packed buffer
};

__host__ __device__ inline int psn_pad4(int k) { return (k + 3) & ~3; }
__host__ __device__ inline int psn_kpad(int k) {
    int p = psn_pad4(k < 1 ? 1 : k);
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}

static inline int psn_S(const psnode_problem* p) { return p->X + p->Z + p->V + p->I; }

// ---- kernel families (each returns a PSNODE_* status) --------------------------------------------
int psn_generic_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_generic_forward_workspace(const psnode_problem* p);
int psn_generic_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_generic_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);
// same kernels with 2 trajectories per CTA: taken when the 8-trajectory layout does not fit shared memory (EUNSUPPORTED)
int psn_generic_forward_tb2(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_generic_forward_workspace_tb2(const psnode_problem* p);
int psn_generic_backward_tb2(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_generic_backward_workspace_tb2(const psnode_problem* p, const psnode_adjoint* a);

bool psn_prefer_tb2(const psnode_problem* p);      // psnode_api.cu

bool psn_fused_supports(const psnode_problem* p);
int psn_fused_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_fused_forward_workspace(const psnode_problem* p);

bool psn_tc_supports(const psnode_problem* p);
int psn_tc_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_tc_forward_workspace(const psnode_problem* p);

// tape-based tensor-core reverse sweep (psnode_tc_bwd.cu)
bool psn_tc_bwd_supports(const psnode_problem* p, const psnode_adjoint* a);
int psn_tc_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_tc_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);
int psn_tc8_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);   // psnode_tc8_fwd.cu
// tape-based tensor-core reverse sweep for the H = 64 DAE nets (psnode_tc_bwd_dae.cu)
bool psn_tc_dae_bwd_supports(const psnode_problem* p, const psnode_adjoint* a);
int psn_tc_dae_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
int64_t psn_tc_dae_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);

// ---- fused masked-MSE upstream gradient (psnode_adjoint.fuse_x / fuse_i) ------------------------------------------------------
#if defined(__CUDACC__)
struct PsnFuse {
    psnode_loss_term term;
    psnode_series sol;          // the forward result the term refers to (p->x_sol / p->i_sol)
};
__device__ __forceinline__ float psn_fuse_scale(const PsnFuse& f) { return f.term.scale ? __ldg(f.term.scale) : 1.0f; }
// d/d sol[j,b,c] of  scale * sum w_c * mask * (sol - target)^2
__device__ __forceinline__ float psn_fuse_grad(const PsnFuse& f, float scale, int j, int b, int c) {
    const float m = __ldg(f.term.mask.p + (int64_t)j * f.term.mask.st + (int64_t)b * f.term.mask.sb);
    const float d = __ldg(f.sol.p + (int64_t)j * f.sol.st + (int64_t)b * f.sol.sb + c) -
                    __ldg(f.term.target.p + (int64_t)j * f.term.target.st + (int64_t)b * f.term.target.sb + c);
    const float w = f.term.feat_weight ? __ldg(f.term.feat_weight + c) : 1.0f;
    return (scale * (2.0f * w) * m) * d;
}
static inline PsnFuse psn_make_fuse(const psnode_loss_term& t, const psnode_series_out& sol) {
    PsnFuse f;
    f.term = t;
    f.sol.p = sol.p; f.sol.st = sol.st; f.sol.sb = sol.sb;
    return f;
}
#endif
