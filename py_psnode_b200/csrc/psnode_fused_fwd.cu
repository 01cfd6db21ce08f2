// psnode_fused_fwd.cu -- the fast forward path for the reference's H = 64 ODE nets (BASELINE configs[1]):
// integrate_ODE (neural_dae/my_solvers.py:52-80) with the 4-layer DE_Func of neural_00_ODE_01_no_encode.py:58-68
// (3S -> 64 -> 64 -> 64 -> 16, ELU), Euler / Midpoint / RK4-3/8 (neural_dae/my_fixed_grid.py:15-59), event jumps.
//
// Design (measurements: bench_micro/micro.cu, profiles/r01_microbench.log)
//   * Every mapping that streams weights from shared memory is shared-memory-bandwidth bound on B200 (broadcast
//     LDS.128 = 2.1 cyc, distinct = 4 cyc against 4 warp-FFMA/cyc): 2300-3000 cyc per 64x64 layer for 28
//     trajectories/SM, <= 40 % of the FMA peak.  So the weights live in REGISTERS for the whole kernel: a CTA is a
//     team of 64 threads that owns TB (7) trajectories for all time steps; thread (ng, ks) holds the 4x16 block
//     W[4ng..4ng+3][16ks..16ks+15] of each 64x64 layer as 32 packed f32x2 registers (168 weight registers per
//     thread for the four layers), 4 teams per SM (4 x 64 x <=255 registers = the whole register file).
//   * Per trajectory and layer a thread loads its 16 activations (4 LDS.128), issues 32 fma.rn.f32x2 (packed over
//     k pairs: half the issue slots of scalar FFMA), and the 4-way split-K is resolved by a reduce-scatter over the
//     4 adjacent lanes (3 SHFL) that leaves thread `tid` holding neuron `tid`: bias + ELU + one STS.
//   * Layer 1 is folded: W1 [a0; s-a0; s] + b1 = (Wb+Wc) s + ((Wa-Wb) a0 + b1).  The bracket is a per-trajectory
//     constant (prologue), the held-input columns are applied once per step, only the 16 state columns per stage.
//   * State x, stage slopes k1..k3 and dt live in registers of the thread that owns element (trajectory, n); the
//     trajectory row is staged in shared memory and written with one 128-bit store per 4 elements (448 contiguous
//     bytes per step and team).  Next-step inputs (t, z / jump values) are prefetched one step ahead into registers.
#include "psnode_internal.cuh"

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 pack2(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}

#ifndef PSN_STAGE_UNROLL
#define PSN_STAGE_UNROLL 1
#endif
constexpr int kStageUnroll = PSN_STAGE_UNROLL;   // 1 = stage loop rolled: the fully unrolled body (4) overflowed the instruction cache (ncu: no_instruction 1.2 warps/issue); rolled is 15 % faster (A/B on B200: 181 -> 209 M traj-steps/s)
constexpr int FH = 64;     // hidden width
constexpr int FX = 16;     // state width
constexpr int FU = 8;      // held-input width (padded)
constexpr int FS = 24;     // S = X + U padded
constexpr int F_NT = 64;   // threads per team / CTA

// folded-weight buffer layout (floats)
constexpr int OFF_W1X = 0;                      // [64][16]  (Wb+Wc) state columns
constexpr int OFF_W1U = OFF_W1X + FH * FX;      // [64][8]   (Wb+Wc) held-input columns, zero padded
constexpr int OFF_W1C = OFF_W1U + FH * FU;      // [64][24]  (Wa-Wb), zero padded
constexpr int OFF_B1 = OFF_W1C + FH * FS;       // [64]
constexpr int OFF_W2 = OFF_B1 + FH;             // [64][64]
constexpr int OFF_B2 = OFF_W2 + FH * FH;
constexpr int OFF_W3 = OFF_B2 + FH;
constexpr int OFF_B3 = OFF_W3 + FH * FH;
constexpr int OFF_W4 = OFF_B3 + FH;             // [16][64]
constexpr int OFF_B4 = OFF_W4 + FX * FH;        // [16]
constexpr int FW_TOTAL = OFF_B4 + FX;

struct FoldArgs {
    const float* W1; const float* b1; const float* W2; const float* b2;
    const float* W3; const float* b3; const float* W4; const float* b4;
    int S, U;
};

__global__ void psn_fused_fold_kernel(const FoldArgs a, float* __restrict__ fw) {
    const int S = a.S, K = 3 * S;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < FW_TOTAL; e += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (e < OFF_W1U) { const int n = e / FX, k = e % FX; v = a.W1[n * K + S + k] + a.W1[n * K + 2 * S + k]; }
        else if (e < OFF_W1C) { const int r = e - OFF_W1U, n = r / FU, c = r % FU; if (c < a.U) v = a.W1[n * K + S + FX + c] + a.W1[n * K + 2 * S + FX + c]; }
        else if (e < OFF_B1) { const int r = e - OFF_W1C, n = r / FS, k = r % FS; if (k < S) v = a.W1[n * K + k] - a.W1[n * K + S + k]; }
        else if (e < OFF_W2) v = a.b1[e - OFF_B1];
        else if (e < OFF_B2) v = a.W2[e - OFF_W2];
        else if (e < OFF_W3) v = a.b2[e - OFF_B2];
        else if (e < OFF_B3) v = a.W3[e - OFF_W3];
        else if (e < OFF_W4) v = a.b3[e - OFF_B3];
        else if (e < OFF_B4) v = a.W4[e - OFF_W4];
        else v = a.b4[e - OFF_B4];
        fw[e] = v;
    }
}

struct FusedParams {
    int B, T, Z, S;
    psnode_series t, x, z;
    const float* a0; int64_t a0_sb;
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    psnode_series_out x_sol;
    const float* fw;
    int vec_out;      // x_sol rows 16-byte aligned -> 128-bit stores
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}

// one 64 -> 64 layer for the team's TB trajectories; thread ends up owning neuron `tid` of every trajectory
template <int TB>
__device__ __forceinline__ void layer64(const float (*__restrict__ in)[FH], float (*__restrict__ out)[FH], const u64 (&w)[32],
                                        const float bias, const int ks, const int tid) {
    const bool hi2 = (ks & 2) != 0, hi1 = (ks & 1) != 0;
#pragma unroll
    for (int i = 0; i < TB; i++) {
        const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(&in[i][16 * ks]);
        u64 a[8];
#pragma unroll
        for (int c = 0; c < 4; c++) { const ulonglong2 v = ap[c]; a[2 * c] = v.x; a[2 * c + 1] = v.y; }
        float s[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            u64 acc = 0ull;
#pragma unroll
            for (int p = 0; p < 8; p++) acc = fma2(a[p], w[q * 8 + p], acc);
            const float2 f = unpack2(acc);
            s[q] = f.x + f.y;
        }
        // reduce-scatter over the 4 k-slices (adjacent lanes): 2 values cross lane^2, 1 value crosses lane^1
        const float send0 = hi2 ? s[0] : s[2], send1 = hi2 ? s[1] : s[3];
        const float keep0 = hi2 ? s[2] : s[0], keep1 = hi2 ? s[3] : s[1];
        const float r0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 2);
        const float r1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 2);
        const float send = hi1 ? r0 : r1, keep = hi1 ? r1 : r0;
        const float o = (keep + __shfl_xor_sync(0xffffffffu, send, 1)) + bias;
        out[i][tid] = psn_elu(o);
    }
}

template <int TB, int METHOD>
__global__ void __launch_bounds__(F_NT, 4) psn_fused_ode_kernel(const __grid_constant__ FusedParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    constexpr int NSLOT = (TB + 3) / 4;
    __shared__ __align__(16) float xs[TB][FX];
    __shared__ __align__(16) float actA[TB][FH];
    __shared__ __align__(16) float actB[TB][FH];
    __shared__ __align__(16) float ubuf[2][TB][FU];
    __shared__ __align__(16) float ostage[TB][FX];
    __shared__ float dts[2][TB];
    __shared__ float c1s[TB][FH];
    __shared__ float a0s[TB][FS];

    const int tid = threadIdx.x;
    const int ks = tid & 3, ng = tid >> 2;
    const int b0 = blockIdx.x * TB;
    const int B = q.B, T = q.T, Z = q.Z, S = q.S;
    const float* __restrict__ fw = q.fw;

    // ---- weights -> registers -------------------------------------------------------------------------
    u64 w1x[8], w2[32], w3[32], w4[8];
    float w1u[FU];
    {
        const ulonglong2* g = reinterpret_cast<const ulonglong2*>(fw + OFF_W1X + tid * FX);
#pragma unroll
        for (int c = 0; c < 4; c++) { const ulonglong2 v = __ldg(g + c); w1x[2 * c] = v.x; w1x[2 * c + 1] = v.y; }
        const float4* gu = reinterpret_cast<const float4*>(fw + OFF_W1U + tid * FU);
        const float4 u0 = __ldg(gu), u1 = __ldg(gu + 1);
        w1u[0] = u0.x; w1u[1] = u0.y; w1u[2] = u0.z; w1u[3] = u0.w; w1u[4] = u1.x; w1u[5] = u1.y; w1u[6] = u1.z; w1u[7] = u1.w;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const ulonglong2* g2 = reinterpret_cast<const ulonglong2*>(fw + OFF_W2 + (4 * ng + r) * FH + 16 * ks);
            const ulonglong2* g3 = reinterpret_cast<const ulonglong2*>(fw + OFF_W3 + (4 * ng + r) * FH + 16 * ks);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const ulonglong2 v2 = __ldg(g2 + c), v3 = __ldg(g3 + c);
                w2[r * 8 + 2 * c] = v2.x; w2[r * 8 + 2 * c + 1] = v2.y;
                w3[r * 8 + 2 * c] = v3.x; w3[r * 8 + 2 * c + 1] = v3.y;
            }
        }
        const ulonglong2* g4 = reinterpret_cast<const ulonglong2*>(fw + OFF_W4 + ng * FH + 16 * ks);
#pragma unroll
        for (int c = 0; c < 4; c++) { const ulonglong2 v = __ldg(g4 + c); w4[2 * c] = v.x; w4[2 * c + 1] = v.y; }
    }
    const float bias2 = __ldg(fw + OFF_B2 + tid), bias3 = __ldg(fw + OFF_B3 + tid), bias4 = __ldg(fw + OFF_B4 + ng);

    // ---- per-trajectory constants ----------------------------------------------------------------------
    for (int e = tid; e < TB * FS; e += F_NT) {
        const int i = e / FS, c = e % FS;
        const int bb = min(b0 + i, B - 1);
        a0s[i][c] = c < S ? __ldg(q.a0 + (int64_t)bb * q.a0_sb + c) : 0.0f;
    }
    // step-1 inputs straight into buffer 1; element e -> (trajectory, column): column 0 = dt, 1.. = held inputs
    auto load_step_input = [&](int j, int i, int c) -> float {   // inputs of the step that ENDS at grid point j
        const int bb = min(b0 + i, B - 1);
        if (c == 0) return __fsub_rn(ldser(q.t, j, bb, 0), ldser(q.t, j - 1, bb, 0));
        const int cz = c - 1;
        const int k = q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1;
        return k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + cz) : ldser(q.z, j - 1, bb, cz);
    };
    const int pf_i = tid / (1 + Z), pf_c = tid % (1 + Z);     // this thread's prefetch element (TB * (1+Z) <= 64)
    const bool pf_on = pf_i < TB;
    for (int e = tid; e < 2 * TB * FU; e += F_NT) (&ubuf[0][0][0])[e] = 0.0f;
    __syncthreads();
    if (T > 1 && pf_on) {
        const float v = load_step_input(1, pf_i, pf_c);
        if (pf_c == 0) dts[1][pf_i] = v; else ubuf[1][pf_i][pf_c - 1] = v;
    }
    {   // c1[i][n] = b1[n] + sum_k (Wa-Wb)[n][k] a0[i][k]
        const float b1 = __ldg(fw + OFF_B1 + tid);
        float c[TB];
#pragma unroll
        for (int i = 0; i < TB; i++) c[i] = b1;
        for (int k = 0; k < S; k++) {
            const float wv = __ldg(fw + OFF_W1C + tid * FS + k);
#pragma unroll
            for (int i = 0; i < TB; i++) c[i] = fmaf(wv, a0s[i][k], c[i]);
        }
#pragma unroll
        for (int i = 0; i < TB; i++) c1s[i][tid] = c[i];
    }
    // initial state: element (i, n = ng) is owned by the thread with ks == (i & 3)
    float x0s[NSLOT], k1[NSLOT], k2[NSLOT], k3[NSLOT];
#pragma unroll
    for (int i = 0; i < TB; i++) {
        if ((i & 3) == ks) {
            const int b = b0 + i, bb = min(b, B - 1);
            const float xv = ldser(q.x, 0, bb, ng);
            x0s[i >> 2] = xv;
            xs[i][ng] = xv;
            if (b < B) q.x_sol.p[(int64_t)b * q.x_sol.sb + ng] = xv;
        }
    }
#pragma unroll
    for (int s = 0; s < NSLOT; s++) { k1[s] = 0.f; k2[s] = 0.f; k3[s] = 0.f; }
    __syncthreads();

    const float c13 = (float)(1.0 / 3.0);
    for (int j = 1; j < T; j++) {
        const int cur = j & 1;
        // prefetch the inputs of step j+1 (consumed after the first barrier of this step)
        float pf = 0.0f;
        const bool do_pf = pf_on && (j + 1 < T);
        if (do_pf) pf = load_step_input(j + 1, pf_i, pf_c);
        // held-input part of layer 1, once per step
        float hz[TB];
#pragma unroll
        for (int i = 0; i < TB; i++) {
            const float4 u0 = *reinterpret_cast<const float4*>(&ubuf[cur][i][0]);
            const float4 u1 = *reinterpret_cast<const float4*>(&ubuf[cur][i][4]);
            float h = c1s[i][tid];
            h = fmaf(w1u[0], u0.x, h); h = fmaf(w1u[1], u0.y, h); h = fmaf(w1u[2], u0.z, h); h = fmaf(w1u[3], u0.w, h);
            h = fmaf(w1u[4], u1.x, h); h = fmaf(w1u[5], u1.y, h); h = fmaf(w1u[6], u1.z, h); h = fmaf(w1u[7], u1.w, h);
            hz[i] = h;
        }
#pragma unroll kStageUnroll
        for (int e = 0; e < NST; e++) {
            // ---- layer 1 (folded): thread = neuron tid, 16 state columns ----
#pragma unroll
            for (int i = 0; i < TB; i++) {
                const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(&xs[i][0]);
                u64 acc = pack2(hz[i], 0.0f);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const ulonglong2 v = xp[c];
                    acc = fma2(v.x, w1x[2 * c], acc);
                    acc = fma2(v.y, w1x[2 * c + 1], acc);
                }
                const float2 f = unpack2(acc);
                actA[i][tid] = psn_elu(f.x + f.y);
            }
            __syncthreads();
            if (e == 0 && do_pf) { if (pf_c == 0) dts[cur ^ 1][pf_i] = pf; else ubuf[cur ^ 1][pf_i][pf_c - 1] = pf; }
            layer64<TB>(actA, actB, w2, bias2, ks, tid);
            __syncthreads();
            layer64<TB>(actB, actA, w3, bias3, ks, tid);
            __syncthreads();
            // ---- layer 4 (64 -> 16) + stage algebra: thread (n = ng, ks) ----
#pragma unroll
            for (int i = 0; i < TB; i++) {
                const ulonglong2* ap = reinterpret_cast<const ulonglong2*>(&actA[i][16 * ks]);
                u64 acc = 0ull;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const ulonglong2 v = ap[c];
                    acc = fma2(v.x, w4[2 * c], acc);
                    acc = fma2(v.y, w4[2 * c + 1], acc);
                }
                const float2 f = unpack2(acc);
                float s = f.x + f.y;
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                const float kv = s + bias4;
                if ((i & 3) == ks) {
                    constexpr int dummy = 0; (void)dummy;
                    const int sl = i >> 2;
                    const float dt = dts[cur][i];
                    const float x0 = x0s[sl];
                    float xn;
                    if (METHOD == PSNODE_EULER) {
                        xn = __fadd_rn(x0, __fmul_rn(dt, kv));
                    } else if (METHOD == PSNODE_MIDPOINT) {
                        if (e == 0) xn = __fadd_rn(x0, __fmul_rn(kv, __fmul_rn(0.5f, dt)));
                        else xn = __fadd_rn(x0, __fmul_rn(dt, kv));
                    } else {
                        if (e == 0) { k1[sl] = kv; xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(dt, kv), c13)); }
                        else if (e == 1) { k2[sl] = kv; xn = __fadd_rn(x0, __fmul_rn(dt, __fsub_rn(kv, __fmul_rn(k1[sl], c13)))); }
                        else if (e == 2) { k3[sl] = kv; xn = __fadd_rn(x0, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[sl], k2[sl]), kv))); }
                        else {
                            const float ksum = __fadd_rn(__fadd_rn(k1[sl], __fmul_rn(3.0f, __fadd_rn(k2[sl], k3[sl]))), kv);
                            xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(ksum, dt), 0.125f));
                        }
                    }
                    xs[i][ng] = xn;
                    if (e == NST - 1) { x0s[sl] = xn; ostage[i][ng] = xn; }
                }
            }
            __syncthreads();
        }
        // ---- trajectory row j: TB x 16 floats, one 128-bit store per 4 elements ----
        if (q.vec_out) {
            if (tid < TB * 4) {
                const int i = tid >> 2, c = tid & 3, b = b0 + i;
                if (b < B)
                    *reinterpret_cast<float4*>(q.x_sol.p + (int64_t)j * q.x_sol.st + (int64_t)b * q.x_sol.sb + 4 * c) =
                        *reinterpret_cast<const float4*>(&ostage[i][4 * c]);
            }
        } else {
            for (int e = tid; e < TB * FX; e += F_NT) {
                const int i = e / FX, c = e % FX, b = b0 + i;
                if (b < B) q.x_sol.p[(int64_t)j * q.x_sol.st + (int64_t)b * q.x_sol.sb + c] = ostage[i][c];
            }
        }
    }
}

template <int TB>
int launch_tb(const psnode_problem* p, const FusedParams& q, cudaStream_t stream) {
    const int grid = (p->B + TB - 1) / TB;
    switch (p->method) {
        case PSNODE_EULER: psn_fused_ode_kernel<TB, PSNODE_EULER><<<grid, F_NT, 0, stream>>>(q); psn_count_launch("psn_fused_ode_kernel<euler>"); break;
        case PSNODE_MIDPOINT: psn_fused_ode_kernel<TB, PSNODE_MIDPOINT><<<grid, F_NT, 0, stream>>>(q); psn_count_launch("psn_fused_ode_kernel<midpoint>"); break;
        default: psn_fused_ode_kernel<TB, PSNODE_RK4><<<grid, F_NT, 0, stream>>>(q); psn_count_launch("psn_fused_ode_kernel<rk4>"); break;
    }
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

}  // namespace

bool psn_fused_supports(const psnode_problem* p) {
    if (p->kind != PSNODE_ODE || p->teacher_x) return false;
    if (p->X != FX || p->Z < 0 || p->Z > FU - 1) return false;      // TB*(1+Z) <= 64 prefetch slots
    const psnode_mlp& m = p->de;
    if (m.n_layers != 4) return false;
    const int S = p->X + p->Z;
    return m.in_dim[0] == 3 * S && m.out_dim[0] == FH && m.out_dim[1] == FH && m.out_dim[2] == FH && m.out_dim[3] == FX;
}

int64_t psn_fused_forward_workspace(const psnode_problem*) { return (int64_t)FW_TOTAL * 4; }

int psn_fused_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < (int64_t)FW_TOTAL * 4) return PSNODE_EWORKSPACE;
    float* fw = static_cast<float*>(ws);
    FoldArgs fa;
    fa.W1 = p->de.W[0]; fa.b1 = p->de.b[0]; fa.W2 = p->de.W[1]; fa.b2 = p->de.b[1];
    fa.W3 = p->de.W[2]; fa.b3 = p->de.b[2]; fa.W4 = p->de.W[3]; fa.b4 = p->de.b[3];
    fa.S = p->X + p->Z; fa.U = p->Z;
    psn_fused_fold_kernel<<<16, 256, 0, stream>>>(fa, fw);
    psn_count_launch("psn_fused_fold_kernel");
    PSN_CUDA(cudaGetLastError());

    FusedParams q;
    q.B = p->B; q.T = p->T; q.Z = p->Z; q.S = p->X + p->Z;
    q.t = p->t; q.x = p->x; q.z = p->z;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.x_sol = p->x_sol;
    q.fw = fw;
    q.vec_out = ((reinterpret_cast<uintptr_t>(p->x_sol.p) & 15) == 0 && (p->x_sol.st & 3) == 0 && (p->x_sol.sb & 3) == 0) ? 1 : 0;
    // team size: 7 trajectories fill 148 SMs x 4 teams with B = 4096 (586 teams for 592 slots); 8 when B is a multiple of 8
    // and that wastes fewer slots
    const int slots = 148 * 4;
    auto cost = [&](int tb) { const int teams = (p->B + tb - 1) / tb; return (int64_t)((teams + slots - 1) / slots) * tb; };
    if (cost(8) < cost(7)) return launch_tb<8>(p, q, stream);
    return launch_tb<7>(p, q, stream);
}
