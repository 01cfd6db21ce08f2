// psnode_api.cu -- the extern "C" surface declared in include/psnode_b200.h: validation, kernel dispatch,
// the device-side event table, launch accounting and the host-buffer convenience entry.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "psnode_internal.cuh"
#include "psnode_tc_tape.cuh"
#include "psnode_wide.cuh"

namespace {
std::atomic<int64_t> g_launches{0};
char g_last_kernel[128] = "none";
char g_last_cuda_error[256] = "";
}  // namespace

void psn_count_launch(const char* kernel_name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    std::snprintf(g_last_kernel, sizeof(g_last_kernel), "%s", kernel_name);
}

int psn_cuda_fail(cudaError_t e, const char* where) {
    std::snprintf(g_last_cuda_error, sizeof(g_last_cuda_error), "%s: %s", where, cudaGetErrorString(e));
    return PSNODE_ECUDA;
}

namespace {

bool mlp_ok(const psnode_mlp& m, int in0, int out_last) {
    if (m.n_layers < 1 || m.n_layers > PSNODE_MAX_LAYERS) return false;
    if (m.in_dim[0] != in0 || m.out_dim[m.n_layers - 1] != out_last) return false;
    for (int l = 0; l < m.n_layers; l++) {
        if (m.in_dim[l] < 1 || m.out_dim[l] < 1 || !m.W[l] || !m.b[l]) return false;
        if (l > 0 && m.in_dim[l] != m.out_dim[l - 1]) return false;
    }
    return true;
}

int validate(const psnode_problem* p) {
    if (!p) return PSNODE_EINVAL;
    if (p->kind != PSNODE_ODE && p->kind != PSNODE_DAE) return PSNODE_EINVAL;
    if (p->method < PSNODE_EULER || p->method > PSNODE_RK4) return PSNODE_EINVAL;
    if (p->B < 1 || p->T < 1 || p->X < 1 || p->Z < 0 || p->V < 0 || p->I < 0) return PSNODE_EINVAL;
    const bool dae = p->kind == PSNODE_DAE;
    if (!dae && (p->V != 0 || p->I != 0)) return PSNODE_EINVAL;
    if (dae && p->I < 1) return PSNODE_EINVAL;
    const int S = psn_S(p);
    if (!mlp_ok(p->de, 3 * S, p->X)) return PSNODE_EINVAL;
    if (dae && !mlp_ok(p->ae, S + p->X + p->Z + p->V, p->I)) return PSNODE_EINVAL;
    if (!p->t.p || !p->a0 || !p->x_sol.p) return PSNODE_EINVAL;
    if (p->Z > 0 && !p->z.p) return PSNODE_EINVAL;
    if (dae) {
        if (!p->x_init || !p->i_sol.p) return PSNODE_EINVAL;
        if (p->V > 0 && !p->v.p) return PSNODE_EINVAL;
        if (p->teacher_i && !p->i.p) return PSNODE_EINVAL;
    }
    if ((!dae || p->teacher_x) && !p->x.p) return PSNODE_EINVAL;
    if (p->event_idx) {
        if (p->E < 1) return PSNODE_EINVAL;
        if (p->Z > 0 && !p->z_jump) return PSNODE_EINVAL;
        if (dae && p->V > 0 && !p->v_jump) return PSNODE_EINVAL;
    }
    return PSNODE_OK;
}

}  // namespace
// generic kernels: prefer the 2-trajectories-per-CTA build when 8 per CTA would leave most of the chip idle (PSNODE_GENERIC_TB2=0/1 forces)
bool psn_prefer_tb2(const psnode_problem* p) {
    static const int forced = std::getenv("PSNODE_GENERIC_TB2") ? std::atoi(std::getenv("PSNODE_GENERIC_TB2")) : -1;
    if (forced >= 0) return forced != 0;
    return p->B <= 8 * 37;          // <= 37 CTAs of 8 trajectories: a quarter of the 148 SMs
}
namespace {
__global__ void psn_event_table_kernel(const float* __restrict__ t0, int64_t t_st, int T, const float* __restrict__ ev0,
                                       int64_t ev_se, int E, int* __restrict__ event_idx, int* __restrict__ err) {
    if (blockIdx.x == 0 && threadIdx.x == 0) err[0] = 0;
    __syncthreads();   // only orders block 0; other blocks can only ever write 1
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < T - 1; j += gridDim.x * blockDim.x) {
        const float tj = t0[(int64_t)j * t_st];
        int hit = -1, nhit = 0;
        for (int k = 0; k < E; k++)
            if (ev0[(int64_t)k * ev_se] == tj) { if (hit < 0) hit = k; nhit++; }
        event_idx[j] = hit;
        if (nhit > 1) atomicExch(err, 1);
    }
}

}  // namespace

extern "C" {

int psnode_abi_version(void) { return PSNODE_ABI_VERSION; }

const char* psnode_status_string(int status) {
    switch (status) {
        case PSNODE_OK: return "ok";
        case PSNODE_EINVAL: return "invalid problem description (dimensions / null pointers)";
        case PSNODE_EUNSUPPORTED: return "problem shape not supported by the requested kernel";
        case PSNODE_EWORKSPACE: return "workspace missing or too small";
        case PSNODE_ECUDA: return "CUDA runtime error (see psnode_last_cuda_error)";
        case PSNODE_ENODEVICE: return "no sm_100 CUDA device";
        default: return "unknown status";
    }
}

const char* psnode_last_cuda_error(void) { return g_last_cuda_error; }
int64_t psnode_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
const char* psnode_last_kernel(void) { return g_last_kernel; }

int64_t psnode_mlp_param_count(const psnode_mlp* m) {
    if (!m) return 0;
    int64_t n = 0;
    for (int l = 0; l < m->n_layers; l++) n += (int64_t)m->out_dim[l] * m->in_dim[l] + m->out_dim[l];
    return n;
}

int psnode_event_table(const float* t0, int64_t t_st, int32_t T, const float* ev0, int64_t ev_se, int32_t E,
                       int32_t* event_idx, int32_t* err, void* stream) {
    if (!t0 || !ev0 || !event_idx || !err || T < 1 || E < 1) return PSNODE_EINVAL;
    if (T == 1) return PSNODE_OK;
    // a single block keeps the err[0] = 0 initialisation ordered before every write of 1
    psn_event_table_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(t0, t_st, T, ev0, ev_se, E, event_idx, err);
    psn_count_launch("psn_event_table_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

int psnode_tape_covers_input_grads(const psnode_problem* p) {
    if (validate(p) != PSNODE_OK) return 0;
    return ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide_supports(p)) ? 1 : 0;
}

int64_t psnode_tape_floats(const psnode_problem* p) {
    if (validate(p) != PSNODE_OK) return 0;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide_supports(p) && !psn_tc_supports(p))
        return psw_tape_floats(p->B, p->T, p->method);
    // the 4-layer ODE_01 net on the wide4 kernels: a tape only when its tensor-core reverse sweep is enabled and the forward dispatch lands there
    if (psn_wide4_bwd_enabled() && psn_wide4_supports(p) && !psn_tc_supports(p) &&
        (p->impl == PSNODE_IMPL_WIDE || (p->impl == PSNODE_IMPL_AUTO && psn_wide4_auto(p))))
        return psw4_tape_floats(p->B, p->T, p->method);
    if (p->impl != PSNODE_IMPL_AUTO && p->impl != PSNODE_IMPL_TC && p->impl != PSNODE_IMPL_TC8) return 0;
    if (!psn_tc_supports(p)) return 0;
    if (p->kind == PSNODE_DAE)      // only the 8-warp forward kernel records the DAE tape
        return p->impl == PSNODE_IMPL_TC ? 0 : psn_tc_dae_tape_floats(p->B, p->T, p->method, p->event_idx ? p->E : 0);
    return psn_tc_tape_floats(p->B, p->T, p->method);
}

int64_t psnode_forward_workspace(const psnode_problem* p) {
    if (validate(p) != PSNODE_OK) return 0;
    int64_t g = psn_generic_forward_workspace(p);
    { const int64_t g2 = psn_generic_forward_workspace_tb2(p); if (g2 > g) g = g2; }
    int64_t f = psn_fused_supports(p) ? psn_fused_forward_workspace(p) : 0;
    int64_t t = psn_tc_supports(p) ? psn_tc_forward_workspace(p) : 0;
    if (f > g) g = f;
    if (psn_wide_supports(p)) { const int64_t w = psn_wide_forward_workspace(p); if (w > g) g = w; }
    if (psn_lg_supports(p)) { const int64_t w = psn_lg_forward_workspace(p); if (w > g) g = w; }
    if (psn_wide4_supports(p)) { const int64_t w = psn_wide4_forward_workspace(p); if (w > g) g = w; }
    return g > t ? g : t;
}

int psnode_forward(const psnode_problem* p, void* workspace, int64_t workspace_bytes, void* stream) {
    const int st = validate(p);
    if (st != PSNODE_OK) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (p->impl == PSNODE_IMPL_TC) {
        if (!psn_tc_supports(p) || p->X != 16) return PSNODE_EUNSUPPORTED;      // the 4-warp kernel keeps X = 16
        return psn_tc_forward(p, workspace, workspace_bytes, s);
    }
    if (p->impl == PSNODE_IMPL_TC8) {
        if (!psn_tc_supports(p)) return PSNODE_EUNSUPPORTED;
        return psn_tc8_forward(p, workspace, workspace_bytes, s);
    }
    if (p->impl == PSNODE_IMPL_FUSED) {
        if (!psn_fused_supports(p)) return PSNODE_EUNSUPPORTED;
        return psn_fused_forward(p, workspace, workspace_bytes, s);
    }
    if (p->impl == PSNODE_IMPL_WIDE) {
        if (psn_wide_supports(p)) return psn_wide_forward(p, workspace, workspace_bytes, s);
        if (psn_wide4_supports(p)) return psn_wide4_forward(p, workspace, workspace_bytes, s);
        return PSNODE_EUNSUPPORTED;
    }
    if (p->impl == PSNODE_IMPL_LAYER) {
        if (!psn_lg_supports(p)) return PSNODE_EUNSUPPORTED;
        return psn_lg_forward(p, workspace, workspace_bytes, s);
    }
    if (p->impl == PSNODE_IMPL_AUTO && psn_tc_supports(p)) return psn_tc8_forward(p, workspace, workspace_bytes, s);
    if (p->impl == PSNODE_IMPL_AUTO && psn_wide_supports(p)) return psn_wide_forward(p, workspace, workspace_bytes, s);
    if (p->impl == PSNODE_IMPL_AUTO && psn_lg_supports(p)) return psn_lg_forward(p, workspace, workspace_bytes, s);
    if (p->impl == PSNODE_IMPL_AUTO && psn_wide4_supports(p) && psn_wide4_auto(p)) return psn_wide4_forward(p, workspace, workspace_bytes, s);
    if (p->impl == PSNODE_IMPL_AUTO && psn_fused_supports(p)) return psn_fused_forward(p, workspace, workspace_bytes, s);
    // small batches: 2 trajectories per CTA instead of 8 puts 4 x as many SMs to work (the kernels are latency-bound per CTA)
    if (psn_prefer_tb2(p)) {
        const int t2 = psn_generic_forward_tb2(p, workspace, workspace_bytes, s);
        if (t2 != PSNODE_EUNSUPPORTED) return t2;
    }
    const int gst = psn_generic_forward(p, workspace, workspace_bytes, s);
    if (gst != PSNODE_EUNSUPPORTED) return gst;
    return psn_generic_forward_tb2(p, workspace, workspace_bytes, s);      // per-trajectory vectors too wide for 8 trajectories per CTA
}

int psnode_sweep_fuses_loss(const psnode_problem* p, const psnode_adjoint* a) {
    if (validate(p) != PSNODE_OK || !a) return 0;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide_bwd_supports(p, a)) return 1;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_TC || p->impl == PSNODE_IMPL_TC8) && psn_tc_bwd_supports(p, a)) return 1;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_TC8) && psn_tc_dae_bwd_supports(p, a)) return 1;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_LAYER) && psn_lg_bwd_supports(p, a)) return 1;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide4_bwd_supports(p, a)) return 1;
    return 0;
}

int64_t psnode_backward_workspace(const psnode_problem* p, const psnode_adjoint* a) {
    if (validate(p) != PSNODE_OK || !a) return 0;
    int64_t g = psn_generic_backward_workspace(p, a);
    { const int64_t g2 = psn_generic_backward_workspace_tb2(p, a); if (g2 > g) g = g2; }
    const int64_t t = !psn_tc_supports(p) ? 0 : (p->kind == PSNODE_ODE ? psn_tc_backward_workspace(p, a) : psn_tc_dae_backward_workspace(p, a));
    if (psn_wide_bwd_supports(p, a)) { const int64_t w = psn_wide_backward_workspace(p, a); if (w > g) g = w; }
    if (psn_lg_bwd_supports(p, a)) { const int64_t w = psn_lg_backward_workspace(p, a); if (w > g) g = w; }
    if (psn_wide4_bwd_supports(p, a)) { const int64_t w = psn_wide4_backward_workspace(p, a); if (w > g) g = w; }
    return g > t ? g : t;
}

int psnode_backward(const psnode_problem* p, const psnode_adjoint* a, void* workspace, int64_t workspace_bytes,
                    void* stream) {
    const int st = validate(p);
    if (st != PSNODE_OK) return st;
    if (!a || !a->d_theta) return PSNODE_EINVAL;
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide_bwd_supports(p, a))
        return psn_wide_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_TC || p->impl == PSNODE_IMPL_TC8) && psn_tc_bwd_supports(p, a))
        return psn_tc_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_TC8) && psn_tc_dae_bwd_supports(p, a))
        return psn_tc_dae_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_LAYER) && psn_lg_bwd_supports(p, a))
        return psn_lg_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if ((p->impl == PSNODE_IMPL_AUTO || p->impl == PSNODE_IMPL_WIDE) && psn_wide4_bwd_supports(p, a))
        return psn_wide4_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (a->fuse_x.target.p || a->fuse_i.target.p) return PSNODE_EUNSUPPORTED;      // the generic recomputing sweeps take gx / gi only
    if (psn_prefer_tb2(p)) {
        const int t2 = psn_generic_backward_tb2(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
        if (t2 != PSNODE_EUNSUPPORTED) return t2;
    }
    const int gst = psn_generic_backward(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (gst != PSNODE_EUNSUPPORTED) return gst;
    return psn_generic_backward_tb2(p, a, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int64_t psnode_forward_encoded_workspace(const psnode_problem* p, const psnode_codec* c) {
    if (psn_wide_encoded_supports(p, c)) return psn_wide_encoded_workspace(p, c);
    if (!psn_lg_encoded_supports(p, c)) return 0;
    return psn_lg_encoded_workspace(p, c);
}

int psnode_forward_encoded(const psnode_problem* p, const psnode_codec* c, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!p || !c) return PSNODE_EINVAL;
    if (psn_wide_encoded_supports(p, c)) return psn_wide_forward_encoded(p, c, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (!psn_lg_encoded_supports(p, c)) return PSNODE_EUNSUPPORTED;
    return psn_lg_forward_encoded(p, c, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
