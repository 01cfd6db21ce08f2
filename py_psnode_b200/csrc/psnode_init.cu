// psnode_init.cu -- Init_Func + all_initial construction folded into one launch.
//
// The DAE models start every forward pass with (neural_01_DAE_01_no_encode.py:50-58, :98-99; neural_01_DAE_02_direct_encode.py:126-127)
//     x0          = init_func(z[0], v[0], i[0])  = MLP(cat(z0, v0, i0))          Linear/ELU chain, Z+V+I -> H -> H -> X
//     all_initial = cat(x0, z[0], v[0], i[0])
// i.e. ~10 small torch launches forward and ~20 backward per training step.  psnode_init_state does both in one launch straight from
// the (strided) first rows of the series, psnode_init_state_backward is its exact reverse mode: given dL/dx0 (initial-state
// gradient of the reverse sweep) and dL/d all_initial it returns the Init_Func parameter gradients and the gradients of the
// three input rows.  fp32 FMA on CUDA cores (B x ~10 K MACs: microseconds); deterministic (per-block slabs, fixed-order reduce).
#include <cstdio>
#include "psnode_internal.cuh"

namespace {

constexpr int TB = 8;                // trajectories per block
constexpr int THREADS = 256;

struct InitParams {
    psnode_mlp net;
    const float* z0; int64_t z_sb;
    const float* v0; int64_t v_sb;
    const float* i0; int64_t i_sb;
    int B, Z, V, I, X, S, maxw;
    float* x0; int64_t x0_sb;
    float* a0; int64_t a0_sb;
    // backward
    const float* d_x0; int64_t d_x0_sb;      // may be NULL
    const float* d_a0; int64_t d_a0_sb;      // may be NULL
    float* slab; int n_theta;                // [gridDim.x][n_theta]
    float* d_z0; int64_t d_z0_sb;
    float* d_v0; int64_t d_v0_sb;
    float* d_i0; int64_t d_i0_sb;
};

// activations of all layers of the block's TB trajectories: act[l] is the INPUT of layer l (act[0] = cat(z0, v0, i0)), act[L] the output
__device__ __forceinline__ float* act_ptr(float* sm, int l, int maxw) { return sm + (size_t)l * TB * maxw; }

__device__ void init_forward_block(const InitParams& q, float* sm, int b0) {
    const int L = q.net.n_layers, in0 = q.Z + q.V + q.I;
    float* a = act_ptr(sm, 0, q.maxw);
    for (int e = threadIdx.x; e < TB * in0; e += blockDim.x) {
        const int n = e / in0, k = e - n * in0, b = min(b0 + n, q.B - 1);
        float v;
        if (k < q.Z) v = __ldg(q.z0 + (int64_t)b * q.z_sb + k);
        else if (k < q.Z + q.V) v = __ldg(q.v0 + (int64_t)b * q.v_sb + (k - q.Z));
        else v = __ldg(q.i0 + (int64_t)b * q.i_sb + (k - q.Z - q.V));
        a[n * q.maxw + k] = v;
    }
    __syncthreads();
    for (int l = 0; l < L; l++) {
        const int in = q.net.in_dim[l], out = q.net.out_dim[l];
        const float* W = q.net.W[l];
        const float* bias = q.net.b[l];
        const float* src = act_ptr(sm, l, q.maxw);
        float* dst = act_ptr(sm, l + 1, q.maxw);
        for (int e = threadIdx.x; e < TB * out; e += blockDim.x) {
            const int n = e / out, m = e - n * out;
            float acc = __ldg(bias + m);
            const float* wr = W + (int64_t)m * in;
            const float* ar = src + n * q.maxw;
            for (int k = 0; k < in; k++) acc = fmaf(__ldg(wr + k), ar[k], acc);
            dst[n * q.maxw + m] = l + 1 < L ? psn_elu(acc) : acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(THREADS) psn_init_state_kernel(const __grid_constant__ InitParams q) {
    extern __shared__ float sm[];
    const int b0 = blockIdx.x * TB;
    init_forward_block(q, sm, b0);
    const float* xo = act_ptr(sm, q.net.n_layers, q.maxw);
    const float* in = act_ptr(sm, 0, q.maxw);
    for (int e = threadIdx.x; e < TB * q.S; e += blockDim.x) {
        const int n = e / q.S, c = e - n * q.S;
        if (b0 + n >= q.B) continue;
        const float v = c < q.X ? xo[n * q.maxw + c] : in[n * q.maxw + (c - q.X)];
        q.a0[(int64_t)(b0 + n) * q.a0_sb + c] = v;
        if (c < q.X) q.x0[(int64_t)(b0 + n) * q.x0_sb + c] = v;
    }
}

__global__ void __launch_bounds__(THREADS) psn_init_state_bwd_kernel(const __grid_constant__ InitParams q) {
    extern __shared__ float sm[];
    const int L = q.net.n_layers, b0 = blockIdx.x * TB;
    init_forward_block(q, sm, b0);
    // delta buffers behind the activations: two ping-pong areas
    float* dA = act_ptr(sm, L + 1, q.maxw);
    float* dB = act_ptr(sm, L + 2, q.maxw);
    // delta of the output layer = dL/dx0 (+ the x0 columns of dL/d all_initial); rows beyond B are zero
    for (int e = threadIdx.x; e < TB * q.X; e += blockDim.x) {
        const int n = e / q.X, c = e - n * q.X;
        float v = 0.0f;
        if (b0 + n < q.B) {
            if (q.d_x0) v += __ldg(q.d_x0 + (int64_t)(b0 + n) * q.d_x0_sb + c);
            if (q.d_a0) v += __ldg(q.d_a0 + (int64_t)(b0 + n) * q.d_a0_sb + c);
        }
        dA[n * q.maxw + c] = v;
    }
    __syncthreads();
    float* slab = q.slab + (int64_t)blockIdx.x * q.n_theta;
    int off_end = q.n_theta;
    float* dcur = dA;
    float* dnext = dB;
    for (int l = L - 1; l >= 0; l--) {
        const int in = q.net.in_dim[l], out = q.net.out_dim[l];
        const int off_b = off_end - out, off_W = off_b - out * in;
        off_end = off_W;
        const float* W = q.net.W[l];
        const float* src = act_ptr(sm, l, q.maxw);
        // parameter gradients of this block's trajectories
        for (int e = threadIdx.x; e < out * in; e += blockDim.x) {
            const int m = e / in, k = e - m * in;
            float s = 0.0f;
#pragma unroll
            for (int n = 0; n < TB; n++) s = fmaf(dcur[n * q.maxw + m], src[n * q.maxw + k], s);
            slab[off_W + e] = s;
        }
        for (int m = threadIdx.x; m < out; m += blockDim.x) {
            float s = 0.0f;
#pragma unroll
            for (int n = 0; n < TB; n++) s += dcur[n * q.maxw + m];
            slab[off_b + m] = s;
        }
        // delta of the layer input (through the ELU of the layer before, whose OUTPUT is `src`)
        for (int e = threadIdx.x; e < TB * in; e += blockDim.x) {
            const int n = e / in, k = e - n * in;
            float s = 0.0f;
            for (int m = 0; m < out; m++) s = fmaf(__ldg(W + (int64_t)m * in + k), dcur[n * q.maxw + m], s);
            if (l > 0) s *= psn_elu_grad_from_out(src[n * q.maxw + k]);
            dnext[n * q.maxw + k] = s;
        }
        __syncthreads();
        float* t = dcur; dcur = dnext; dnext = t;
    }
    // gradients of the three input rows: Init_Func path + their own columns of all_initial
    const int in0 = q.Z + q.V + q.I;
    for (int e = threadIdx.x; e < TB * in0; e += blockDim.x) {
        const int n = e / in0, k = e - n * in0;
        if (b0 + n >= q.B) continue;
        float v = dcur[n * q.maxw + k];
        if (q.d_a0) v += __ldg(q.d_a0 + (int64_t)(b0 + n) * q.d_a0_sb + q.X + k);
        if (k < q.Z) { if (q.d_z0) q.d_z0[(int64_t)(b0 + n) * q.d_z0_sb + k] = v; }
        else if (k < q.Z + q.V) { if (q.d_v0) q.d_v0[(int64_t)(b0 + n) * q.d_v0_sb + (k - q.Z)] = v; }
        else if (q.d_i0) q.d_i0[(int64_t)(b0 + n) * q.d_i0_sb + (k - q.Z - q.V)] = v;
    }
}

__global__ void psn_init_reduce_kernel(const float* __restrict__ slab, int nblocks, int n_theta, float* __restrict__ d_theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_theta) return;
    float s = 0.0f;
    for (int b = 0; b < nblocks; b++) s += slab[(int64_t)b * n_theta + idx];
    d_theta[idx] = s;
}

bool fill(InitParams& q, const psnode_mlp* init, const float* z0, int64_t z_sb, const float* v0, int64_t v_sb, const float* i0, int64_t i_sb, int B,
          int Z, int V, int I) {
    if (!init || init->n_layers < 1 || init->n_layers > PSNODE_MAX_LAYERS || B < 1 || Z < 0 || V < 0 || I < 0 || Z + V + I < 1) return false;
    if ((Z > 0 && !z0) || (V > 0 && !v0) || (I > 0 && !i0)) return false;
    if (init->in_dim[0] != Z + V + I) return false;
    int maxw = Z + V + I;
    for (int l = 0; l < init->n_layers; l++) {
        if (!init->W[l] || !init->b[l] || init->out_dim[l] < 1) return false;
        if (l > 0 && init->in_dim[l] != init->out_dim[l - 1]) return false;
        if (init->out_dim[l] > maxw) maxw = init->out_dim[l];
    }
    q.net = *init;
    q.z0 = z0; q.z_sb = z_sb; q.v0 = v0; q.v_sb = v_sb; q.i0 = i0; q.i_sb = i_sb;
    q.B = B; q.Z = Z; q.V = V; q.I = I;
    q.X = init->out_dim[init->n_layers - 1];
    q.S = q.X + Z + V + I;
    q.maxw = maxw;
    return true;
}
size_t smem_bytes(const InitParams& q, int extra) { return (size_t)(q.net.n_layers + 1 + extra) * TB * q.maxw * sizeof(float); }

}  // namespace

extern "C" {

int psnode_init_state(const psnode_mlp* init, const float* z0, int64_t z_sb, const float* v0, int64_t v_sb, const float* i0, int64_t i_sb,
                      int32_t B, int32_t Z, int32_t V, int32_t I, float* x0, int64_t x0_sb, float* a0, int64_t a0_sb, void* stream) {
    InitParams q = {};
    if (!fill(q, init, z0, z_sb, v0, v_sb, i0, i_sb, B, Z, V, I) || !x0 || !a0) return PSNODE_EINVAL;
    q.x0 = x0; q.x0_sb = x0_sb; q.a0 = a0; q.a0_sb = a0_sb;
    const size_t smem = smem_bytes(q, 0);
    if (smem > 200 * 1024) return PSNODE_EUNSUPPORTED;
    PSN_CUDA(cudaFuncSetAttribute(psn_init_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psn_init_state_kernel<<<(B + TB - 1) / TB, THREADS, smem, static_cast<cudaStream_t>(stream)>>>(q);
    psn_count_launch("psn_init_state_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

int64_t psnode_init_state_backward_workspace(const psnode_mlp* init, int32_t B) {
    if (!init || B < 1) return 0;
    return (int64_t)((B + TB - 1) / TB) * psnode_mlp_param_count(init) * 4;
}

int psnode_init_state_backward(const psnode_mlp* init, const float* z0, int64_t z_sb, const float* v0, int64_t v_sb, const float* i0, int64_t i_sb,
                               int32_t B, int32_t Z, int32_t V, int32_t I, const float* d_x0, int64_t d_x0_sb, const float* d_a0, int64_t d_a0_sb,
                               float* d_theta, float* d_z0, int64_t d_z0_sb, float* d_v0, int64_t d_v0_sb, float* d_i0, int64_t d_i0_sb,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    InitParams q = {};
    if (!fill(q, init, z0, z_sb, v0, v_sb, i0, i_sb, B, Z, V, I) || !d_theta || (!d_x0 && !d_a0)) return PSNODE_EINVAL;
    if (!workspace || workspace_bytes < psnode_init_state_backward_workspace(init, B)) return PSNODE_EWORKSPACE;
    q.d_x0 = d_x0; q.d_x0_sb = d_x0_sb; q.d_a0 = d_a0; q.d_a0_sb = d_a0_sb;
    q.slab = static_cast<float*>(workspace);
    q.n_theta = (int)psnode_mlp_param_count(init);
    q.d_z0 = d_z0; q.d_z0_sb = d_z0_sb; q.d_v0 = d_v0; q.d_v0_sb = d_v0_sb; q.d_i0 = d_i0; q.d_i0_sb = d_i0_sb;
    const size_t smem = smem_bytes(q, 2);
    if (smem > 200 * 1024) return PSNODE_EUNSUPPORTED;
    const int nblocks = (B + TB - 1) / TB;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PSN_CUDA(cudaFuncSetAttribute(psn_init_state_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    psn_init_state_bwd_kernel<<<nblocks, THREADS, smem, s>>>(q);
    psn_count_launch("psn_init_state_bwd_kernel");
    psn_init_reduce_kernel<<<(q.n_theta + 255) / 256, 256, 0, s>>>(q.slab, nblocks, q.n_theta, d_theta);
    psn_count_launch("psn_init_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

}  // extern "C"
