// psnode_loss.cu -- the training scripts' masked squared-error loss and its gradient, one pass each.
//
// Reference (neural_00_ODE_01_no_encode.py:353-355):
//     x_loss = sum_b sum_t ( mse_loss(x_pred, x, reduction='none') * mask ) / sum(mask);   loss = sum(x_loss)
// and the DAE script's weighted variant (neural_01_DAE_01_no_encode.py:414-418: feature 1 counted 10 x).  Both are
//     numerator / sum(mask),   numerator = sum_{o,i,c} w_c * mask[o,i] * (pred[o,i,c] - target[o,i,c])^2 .
// In torch eager the numerator and its backward are ~10 elementwise / reduction launches over the (T,B,X) trajectory
// (3.4 GB of HBM traffic at cfg2); here the forward reads pred, target, mask once and the backward reads them once more and
// writes d numerator / d pred (1.3 GB).  HBM bound; grid = 4 CTAs per SM, grid-stride over rows,
// one thread per X-float row with 128-bit accesses (a warp covers 32 consecutive rows; coalesced for either (T,B,X) or
// (B,T,X) storage: the caller passes the dimension with the smaller stride as the inner one).  The reduction order is fixed (per-thread fp32 partials -> per-CTA double ->
// one CTA sums the CTA partials), so the loss is bit-reproducible run to run.
#include "psnode_internal.cuh"

namespace {

constexpr int L_THREADS = 256;

struct LossParams {
    const float* pred; int64_t p_so, p_si;
    const float* target; int64_t t_so, t_si;
    const float* mask; int64_t m_so, m_si;
    const float* w;            // X per-feature weights or nullptr
    int n_outer, n_inner, X;
};

// One thread per (outer, inner) row of X contiguous floats: 128-bit loads when VEC (X % 4 == 0, 16-byte aligned rows).
template <bool VEC>
__global__ void __launch_bounds__(L_THREADS) psn_masked_sse_kernel(const LossParams q, double* __restrict__ partial) {
    float acc = 0.0f;
    for (int o = blockIdx.y; o < q.n_outer; o += gridDim.y) {
        const float* pp = q.pred + (int64_t)o * q.p_so;
        const float* tp = q.target + (int64_t)o * q.t_so;
        const float* mp = q.mask + (int64_t)o * q.m_so;
        for (int i = blockIdx.x * L_THREADS + threadIdx.x; i < q.n_inner; i += gridDim.x * L_THREADS) {
            const float m = __ldg(mp + (int64_t)i * q.m_si);
            const float* pr = pp + (int64_t)i * q.p_si;
            const float* tr = tp + (int64_t)i * q.t_si;
            float row = 0.0f;
            if (VEC) {
                for (int c = 0; c < q.X; c += 4) {
                    const float4 a = __ldcs(reinterpret_cast<const float4*>(pr + c)), b = __ldcs(reinterpret_cast<const float4*>(tr + c));
                    float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (q.w) w = __ldg(reinterpret_cast<const float4*>(q.w + c));
                    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
                    row = fmaf(w.x * d0, d0, row); row = fmaf(w.y * d1, d1, row);
                    row = fmaf(w.z * d2, d2, row); row = fmaf(w.w * d3, d3, row);
                }
            } else {
                for (int c = 0; c < q.X; c++) {
                    const float d = __ldg(pr + c) - __ldg(tr + c);
                    row = fmaf((q.w ? __ldg(q.w + c) : 1.0f) * d, d, row);
                }
            }
            acc = fmaf(m, row, acc);
        }
    }
    __shared__ double red[L_THREADS / 32];
    double v = (double)acc;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < L_THREADS / 32; k++) t += red[k];
        partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(L_THREADS) psn_masked_sse_final_kernel(const double* __restrict__ partial, int n, float* __restrict__ loss) {
    __shared__ double red[L_THREADS];
    double t = 0.0;
    for (int k = threadIdx.x; k < n; k += L_THREADS) t += partial[k];
    red[threadIdx.x] = t;
    __syncthreads();
    for (int s = L_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)red[0];
}

// grad[o,i,c] = upstream * 2 * w_c * mask[o,i] * (pred - target).  VEC: consecutive threads own consecutive 16-byte
// pieces of an outer slice (one instruction of a warp writes 512 contiguous bytes = whole sectors; with a row per thread
// every 32-byte sector was written in two halves and L2 fetched it from DRAM first: 938 MB read instead of 540 MB).
template <bool VEC>
__global__ void __launch_bounds__(L_THREADS) psn_masked_sse_grad_kernel(const LossParams q, const float* __restrict__ upstream,
                                                                        float* __restrict__ grad, int64_t g_so, int64_t g_si) {
    const float up2 = 2.0f * __ldg(upstream);
    const int x4 = q.X >> 2;
    for (int o = blockIdx.y; o < q.n_outer; o += gridDim.y) {
        const float* pp = q.pred + (int64_t)o * q.p_so;
        const float* tp = q.target + (int64_t)o * q.t_so;
        const float* mp = q.mask + (int64_t)o * q.m_so;
        float* gp = grad + (int64_t)o * g_so;
        if (VEC) {
            const int n4 = q.n_inner * x4;
            for (int f = blockIdx.x * L_THREADS + threadIdx.x; f < n4; f += gridDim.x * L_THREADS) {
                const int i = f / x4, c = (f - i * x4) * 4;
                const float m = up2 * __ldg(mp + (int64_t)i * q.m_si);
                const float4 a = __ldcs(reinterpret_cast<const float4*>(pp + (int64_t)i * q.p_si + c));
                const float4 b = __ldcs(reinterpret_cast<const float4*>(tp + (int64_t)i * q.t_si + c));
                float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                if (q.w) w = __ldg(reinterpret_cast<const float4*>(q.w + c));
                *reinterpret_cast<float4*>(gp + (int64_t)i * g_si + c) =
                    make_float4(m * w.x * (a.x - b.x), m * w.y * (a.y - b.y), m * w.z * (a.z - b.z), m * w.w * (a.w - b.w));
            }
        } else {
            for (int i = blockIdx.x * L_THREADS + threadIdx.x; i < q.n_inner; i += gridDim.x * L_THREADS) {
                const float m = up2 * __ldg(mp + (int64_t)i * q.m_si);
                const float* pr = pp + (int64_t)i * q.p_si;
                const float* tr = tp + (int64_t)i * q.t_si;
                float* gr = gp + (int64_t)i * g_si;
                for (int c = 0; c < q.X; c++) gr[c] = m * (q.w ? __ldg(q.w + c) : 1.0f) * (__ldg(pr + c) - __ldg(tr + c));
            }
        }
    }
}

// grid: x over the rows of one outer slice, y over outer slices; about 4 CTAs per SM in total
void loss_grid(int n_outer, int n_inner, dim3& grid) {
    int gx = (n_inner + L_THREADS - 1) / L_THREADS;
    if (gx > 592) gx = 592;
    int gy = 592 / gx;
    if (gy < 1) gy = 1;
    if (gy > n_outer) gy = n_outer;
    grid = dim3((unsigned)gx, (unsigned)gy, 1);
}

bool aligned16(const void* p, int64_t a, int64_t b) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (a & 3) == 0 && (b & 3) == 0; }

bool fill(LossParams& q, const psnode_series* pred, const psnode_series* target, const psnode_series* mask, const float* w,
          int n_outer, int n_inner, int X) {
    if (!pred || !target || !mask || !pred->p || !target->p || !mask->p || n_outer < 1 || n_inner < 1 || X < 1) return false;
    q.pred = pred->p; q.p_so = pred->st; q.p_si = pred->sb;
    q.target = target->p; q.t_so = target->st; q.t_si = target->sb;
    q.mask = mask->p; q.m_so = mask->st; q.m_si = mask->sb;
    q.w = w; q.n_outer = n_outer; q.n_inner = n_inner; q.X = X;
    return true;
}

}  // namespace

int64_t psnode_masked_sse_workspace(void) { return 1024 * (int64_t)sizeof(double); }

int psnode_masked_sse(const psnode_series* pred, const psnode_series* target, const psnode_series* mask, const float* feat_weight,
                      int32_t n_outer, int32_t n_inner, int32_t X, float* loss, void* workspace, int64_t workspace_bytes, void* stream) {
    LossParams q;
    if (!fill(q, pred, target, mask, feat_weight, n_outer, n_inner, X) || !loss) return PSNODE_EINVAL;
    if (!workspace || workspace_bytes < psnode_masked_sse_workspace()) return PSNODE_EWORKSPACE;
    dim3 grid;
    loss_grid(n_outer, n_inner, grid);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(workspace);
    const bool vec = (X & 3) == 0 && aligned16(q.pred, q.p_so, q.p_si) && aligned16(q.target, q.t_so, q.t_si) && (!q.w || aligned16(q.w, 0, 0));
    if (vec) psn_masked_sse_kernel<true><<<grid, L_THREADS, 0, s>>>(q, partial);
    else psn_masked_sse_kernel<false><<<grid, L_THREADS, 0, s>>>(q, partial);
    psn_count_launch("psn_masked_sse_kernel");
    PSN_CUDA(cudaGetLastError());
    psn_masked_sse_final_kernel<<<1, L_THREADS, 0, s>>>(partial, (int)(grid.x * grid.y), loss);
    psn_count_launch("psn_masked_sse_final_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

int psnode_masked_sse_grad(const psnode_series* pred, const psnode_series* target, const psnode_series* mask, const float* feat_weight,
                           int32_t n_outer, int32_t n_inner, int32_t X, const float* upstream, const psnode_series_out* grad, void* stream) {
    LossParams q;
    if (!fill(q, pred, target, mask, feat_weight, n_outer, n_inner, X) || !upstream || !grad || !grad->p) return PSNODE_EINVAL;
    const bool vec = (X & 3) == 0 && aligned16(q.pred, q.p_so, q.p_si) && aligned16(q.target, q.t_so, q.t_si) &&
                     aligned16(grad->p, grad->st, grad->sb) && (!q.w || aligned16(q.w, 0, 0)) && (int64_t)n_inner * (X >> 2) < (1ll << 31);
    dim3 grid;
    loss_grid(n_outer, vec ? n_inner * (X >> 2) : n_inner, grid);
    if (vec) psn_masked_sse_grad_kernel<true><<<grid, L_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(q, upstream, grad->p, grad->st, grad->sb);
    else psn_masked_sse_grad_kernel<false><<<grid, L_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(q, upstream, grad->p, grad->st, grad->sb);
    psn_count_launch("psn_masked_sse_grad_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
