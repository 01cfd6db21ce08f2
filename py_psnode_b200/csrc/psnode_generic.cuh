// psnode_generic.cuh -- pieces shared by the generic forward (psnode_generic_fwd.cu) and the generic reverse sweep
// (psnode_generic_bwd.cu): weight packing, the shared-memory plan, and the Linear(+ELU) layer routine.
// Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include "psnode_internal.cuh"

namespace {


// Trajectories per CTA.  The default build uses 8; psnode_generic_tb2.cu compiles the same kernels with 2 for the latent
// widths of the `*_02_direct_encode` models at H = 256 (BASELINE configs[4]: S = 1024, layer-1 input 3072 floats per
// trajectory -- 8 trajectories of per-trajectory vectors alone would exceed the 227 KB of shared memory).
#ifndef PSN_G_TB
#define PSN_G_TB 8
#endif
#ifndef PSN_G_NAME
#define PSN_G_NAME(x) x
#endif
constexpr int G_TB = PSN_G_TB;                 // trajectories per CTA
constexpr int G_TM = G_TB < 4 ? G_TB : 4;      // trajectories per work item (register tile height)
constexpr int G_NT = 128;   // threads per CTA (nets resident in shared memory)
constexpr int G_NT_MAX = 512;  // threads per CTA when layers stream from L2 (wide latent nets)
constexpr int G_MAXNETLAYERS = 2 * PSNODE_MAX_LAYERS;

struct PackDesc {
    const float* W[G_MAXNETLAYERS];
    const float* b[G_MAXNETLAYERS];
    int in[G_MAXNETLAYERS], out[G_MAXNETLAYERS], kpad[G_MAXNETLAYERS], w_off[G_MAXNETLAYERS], b_off[G_MAXNETLAYERS];
    int n;
};

__global__ void psn_pack_kernel(const __grid_constant__ PackDesc d, float* __restrict__ packed) {
    for (int l = blockIdx.y; l < d.n; l += gridDim.y) {
        const int kpad = d.kpad[l], in = d.in[l], out = d.out[l];
        const int total = out * kpad;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
            const int n = e / kpad, k = e - n * kpad;
            packed[d.w_off[l] + e] = k < in ? d.W[l][(size_t)n * in + k] : 0.0f;
        }
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < out; e += gridDim.x * blockDim.x)
            packed[d.b_off[l] + e] = d.b[l][e];
    }
}

struct GenericParams {
    psnode_problem p;
    PsnPackedNet de, ae;
    const float* packed;
    int S, S4, K0, KA0, HM, X4, I4;
    // shared-memory offsets, in floats
    int o_w, o_a0, o_u3, o_uae, o_actA, o_actB, o_xprev, o_start, o_k1, o_k2, o_k3, o_k4, o_iprev, o_dt;
    int buf_begin, total_floats;
};

__device__ __forceinline__ float ld_series(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}

// one Linear (+ ELU) layer over the CTA's G_TB trajectories; work item = (neuron, group of G_TM trajectories)
template <bool SMEMW>
__device__ __forceinline__ void layer(const float* __restrict__ W, const float* __restrict__ bias, int kpad, int nout,
                                      const float* __restrict__ src, int ss, float* __restrict__ dst, int ds, int dcols,
                                      bool elu) {
    constexpr int NG = G_TB / G_TM;
    for (int it = threadIdx.x; it < nout * NG; it += blockDim.x) {
        const int g = it / nout, n = it - g * nout;
        const float* wrow = W + (size_t)n * kpad;
        const float* arow = src + g * G_TM * ss;
        float acc[G_TM];
        const float bn = __ldg(bias + n);
#pragma unroll
        for (int m = 0; m < G_TM; m++) acc[m] = bn;
        for (int k = 0; k < kpad; k += 4) {
            float4 w;
            if (SMEMW) w = *reinterpret_cast<const float4*>(wrow + k);
            else w = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
            for (int m = 0; m < G_TM; m++) {
                const float4 a = *reinterpret_cast<const float4*>(arow + m * ss + k);
                acc[m] = fmaf(a.x, w.x, acc[m]);
                acc[m] = fmaf(a.y, w.y, acc[m]);
                acc[m] = fmaf(a.z, w.z, acc[m]);
                acc[m] = fmaf(a.w, w.w, acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < G_TM; m++) dst[(g * G_TM + m) * ds + n] = elu ? psn_elu(acc[m]) : acc[m];
    }
    const int padc = dcols - nout;   // keep the next layer's zero-padded input tail clean
    if (padc > 0)
        for (int e = threadIdx.x; e < G_TB * padc; e += blockDim.x) dst[(e / padc) * ds + nout + (e % padc)] = 0.0f;
}

// whole net; every layer ends with a CTA barrier, so `out` is visible to all threads on return
__device__ void run_mlp(const PsnPackedNet& net, const float* __restrict__ packed, const float* __restrict__ wsm,
                        const float* in, int in_stride, float* out, int out_stride, int out_cols, float* actA,
                        float* actB, int HM) {
    const float* src = in;
    int ss = in_stride;
    for (int l = 0; l < net.n_layers; l++) {
        const bool last = (l == net.n_layers - 1);
        float* dst = last ? out : ((l & 1) ? actB : actA);
        const int ds = last ? out_stride : HM;
        const int dcols = last ? out_cols : net.kpad[l + 1];
        const float* bias = packed + net.b_off[l];
        if (net.smem_off[l] >= 0)
            layer<true>(wsm + net.smem_off[l], bias, net.kpad[l], net.out_dim[l], src, ss, dst, ds, dcols, !last);
        else
            layer<false>(packed + net.w_off[l], bias, net.kpad[l], net.out_dim[l], src, ss, dst, ds, dcols, !last);
        __syncthreads();
        src = dst;
        ss = ds;
    }
}

// ---- host side -------------------------------------------------------------------------------------

void pack_layout(const psnode_mlp& m, PsnPackedNet& pn, int& cursor) {
    pn.n_layers = m.n_layers;
    pn.total = 0;
    for (int l = 0; l < m.n_layers; l++) {
        pn.in_dim[l] = m.in_dim[l];
        pn.out_dim[l] = m.out_dim[l];
        pn.kpad[l] = psn_kpad(m.in_dim[l]);
        pn.w_off[l] = cursor;
        cursor += m.out_dim[l] * pn.kpad[l];
        pn.b_off[l] = cursor;
        cursor += psn_pad4(m.out_dim[l]);
        pn.smem_off[l] = -1;
    }
}

int build_params(const psnode_problem* p, GenericParams& q, int& packed_floats, int max_smem_bytes) {
    q.p = *p;
    const bool dae = p->kind == PSNODE_DAE;
    int cursor = 0;
    pack_layout(p->de, q.de, cursor);
    if (dae) pack_layout(p->ae, q.ae, cursor);
    else q.ae.n_layers = 0;
    packed_floats = cursor;
    q.S = psn_S(p);
    q.S4 = psn_pad4(q.S);
    q.X4 = psn_pad4(p->X);
    q.I4 = psn_pad4(p->I > 0 ? p->I : 1);
    q.K0 = q.de.kpad[0];
    q.KA0 = dae ? q.ae.kpad[0] : 4;
    int hm = 4;
    for (int l = 1; l < q.de.n_layers; l++) hm = hm > q.de.kpad[l] ? hm : q.de.kpad[l];
    for (int l = 1; l < q.ae.n_layers; l++) hm = hm > q.ae.kpad[l] ? hm : q.ae.kpad[l];
    q.HM = hm;
    // buffers first (so the weight region gets whatever is left)
    int o = 0;
    auto take = [&](int n) { int r = o; o += psn_pad4(n); return r; };
    q.buf_begin = 0;
    q.o_a0 = take(G_TB * q.S4);
    q.o_u3 = take(G_TB * q.K0);
    q.o_uae = take(dae ? G_TB * q.KA0 : 4);
    q.o_actA = take(G_TB * q.HM);
    q.o_actB = take(G_TB * q.HM);
    q.o_xprev = take(G_TB * q.X4);
    q.o_start = take(G_TB * q.X4);
    q.o_k1 = take(G_TB * q.X4);
    q.o_k2 = take(G_TB * q.X4);
    q.o_k3 = take(G_TB * q.X4);
    q.o_k4 = take(G_TB * q.X4);
    q.o_iprev = take(G_TB * q.I4);
    q.o_dt = take(G_TB);
    q.total_floats = o;     // end of the zero-initialised buffer region
    q.o_w = o;
    const int budget = max_smem_bytes / 4 - o;
    if (budget < 0) return PSNODE_EUNSUPPORTED;
    // greedy placement of layers into the remaining shared memory, smallest first (more layers resident)
    struct Item { int net, l, sz; } items[G_MAXNETLAYERS];
    int n = 0;
    for (int l = 0; l < q.de.n_layers; l++) items[n++] = {0, l, q.de.out_dim[l] * q.de.kpad[l]};
    for (int l = 0; l < q.ae.n_layers; l++) items[n++] = {1, l, q.ae.out_dim[l] * q.ae.kpad[l]};
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (items[b].sz < items[a].sz) { Item t = items[a]; items[a] = items[b]; items[b] = t; }
    int used = 0;
    for (int a = 0; a < n; a++) {
        if (used + items[a].sz > budget) continue;
        (items[a].net ? q.ae : q.de).smem_off[items[a].l] = used;
        used += items[a].sz;
    }
    q.total_floats = o;            // zero-fill stops here; weights follow
    return (o + used) * 4;          // dynamic shared memory bytes
}


}  // namespace
