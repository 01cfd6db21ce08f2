// psnode_generic_bwd.cu -- generic reverse sweep (discrete adjoint) of the fixed-grid integrator: any widths, any layer
// count, ODE and DAE, Euler / Midpoint / RK4-3/8, events, teacher forcing, gradients of every differentiable input.
//
// Replaces the autograd graph the reference builds while unrolling its Python loop and replays at loss.backward()
// (neural_00_ODE_01_no_encode.py:359, neural_00_ODE_02_direct_encode.py:274, neural_01_DAE_01_no_encode.py:423,
// neural_01_DAE_02_direct_encode.py:369): exact reverse mode of FixedGridODESolver.integrate_ODE / integrate_DAE
// (neural_dae/my_solvers.py:52-80, 82-131) and of the step functions (neural_dae/my_fixed_grid.py:15-59).
//
// One persistent CTA owns a tile of G_TB trajectories and walks the grid BACKWARDS.  The forward trajectory
// (x_sol / i_sol, already in HBM) is the checkpoint: for step j the CTA reloads x_sol[j-1], recomputes the 1/2/4 stages
// with every layer's activations kept in shared memory, then back-propagates through the stages in reverse order.
// Weight gradients are accumulated over all steps in shared memory (layers that do not fit accumulate in the CTA's
// private slab in the workspace); at the end every CTA flushes to its slab and a second kernel sums the slabs in a fixed
// order, so the result is deterministic (no floating-point atomics anywhere).
//
// Reverse-mode bookkeeping per grid point j (lam = dL/dx_j, mu = dL/di_j, pend_* = gradient of the INPUT series at j):
//   point j : lam += gx[j]; mu += gi[j]; DAE: back through i_j = ae(x_j | x[j], z[j], v[j])  (my_solvers.py:121 / :95)
//             -> emit d_z[j], d_v[j], d_xteach[j], d_iteach[j]
//   step  j : recompute from x_sol[j-1] (or the teacher x[j-1]); back through x_j = start + dx(start, held inputs)
//             -> lam / pend_xt for j-1; held-input gradients go to d_z[j-1] / d_v[j-1] (or the jump tensors on an event
//             step, where additionally the recomputed i_0 = ae(x_{j-1}, z_jump, v_jump) is back-propagated, :109-110).
#include "psnode_generic.cuh"

namespace {

constexpr int R_NT_MAX = 1024;   // threads per CTA: 256 for nets that live in shared memory, 1024 when weights / gradients
                                // stream through L2 (wide latent nets: the sweep is then bound by L2 latency, not by issue slots)
constexpr int R_MAXST = 4;  // stages

struct BwdParams {
    psnode_problem p;
    psnode_adjoint a;
    PsnPackedNet de, ae;
    const float* packed;
    float* slab;             // [gridDim.x][packed_floats]
    int packed_floats, n_tiles, nst;
    int S, S4, K0, KA0, HM, X4, I4, Z4, V4, U, U4, DM;
    int gw_de[PSNODE_MAX_LAYERS], gw_ae[PSNODE_MAX_LAYERS];   // smem float offset of the layer's dW accumulator, or -1 (slab)
    int gb_de[PSNODE_MAX_LAYERS], gb_ae[PSNODE_MAX_LAYERS];   // smem float offset of the layer's db accumulator (always smem)
    int o_a0, o_da0, o_u3, o_uae, o_ys, o_acts, o_aeacts, o_kst, o_dk, o_dA, o_dB, o_lam, o_mu, o_dxs, o_dy, o_pz, o_pv, o_pxt,
        o_pit, o_dheld, o_ihold, o_start, o_itmp, o_dt, o_w, o_g;
    int zero_begin, zero_end, g_begin, g_end;
};

// forward net evaluation that keeps every hidden activation: layer l output -> acts + l*G_TB*HM (post-ELU)
__device__ void run_mlp_keep(const PsnPackedNet& net, const float* __restrict__ packed, const float* __restrict__ wsm,
                             const float* in, int in_stride, float* out, int out_stride, int out_cols, float* acts, int HM) {
    const float* src = in;
    int ss = in_stride;
    for (int l = 0; l < net.n_layers; l++) {
        const bool last = (l == net.n_layers - 1);
        float* dst = last ? out : acts + l * G_TB * HM;
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

// Back-propagate delta (dcur[i*DM + n], n < out_dim[last]) through the net.  Accumulates dW / db, returns the buffer that
// holds dL/d(input) (width in_dim[0]).  Ends with a barrier.
__device__ float* mlp_backward(const PsnPackedNet& net, const float* __restrict__ packed, const float* __restrict__ wsm,
                               float* sm, const int* gw_off, const int* gb_off, float* slab, const float* u, int us,
                               const float* acts, int HM, float* dcur, float* dnext, int DM) {
    const int tid = threadIdx.x;
    for (int l = net.n_layers - 1; l >= 0; l--) {
        const float* in = l == 0 ? u : acts + (l - 1) * G_TB * HM;
        const int ins = l == 0 ? us : HM;
        const int K = net.in_dim[l], N = net.out_dim[l], kp = net.kpad[l], kq = kp >> 2;
        // ---- dW[n][k] += sum_i delta[i][n] * in[i][k]  (4 columns per work item; pad columns of `in` are zero) ----
        float* gW = gw_off[l] >= 0 ? sm + gw_off[l] : slab + net.w_off[l];
        for (int e = tid; e < N * kq; e += (int)blockDim.x) {
            const int n = e / kq, k4 = (e - n * kq) << 2;
            float4 g = *reinterpret_cast<const float4*>(gW + (size_t)n * kp + k4);
#pragma unroll
            for (int i = 0; i < G_TB; i++) {
                const float d = dcur[i * DM + n];
                const float4 a = *reinterpret_cast<const float4*>(in + i * ins + k4);
                g.x = fmaf(d, a.x, g.x); g.y = fmaf(d, a.y, g.y); g.z = fmaf(d, a.z, g.z); g.w = fmaf(d, a.w, g.w);
            }
            *reinterpret_cast<float4*>(gW + (size_t)n * kp + k4) = g;
        }
        float* gB = sm + gb_off[l];
        for (int n = tid; n < N; n += (int)blockDim.x) {
            float g = gB[n];
#pragma unroll
            for (int i = 0; i < G_TB; i++) g += dcur[i * DM + n];
            gB[n] = g;
        }
        // ---- delta_in[i][k] = (sum_n W[n][k] delta[i][n]) * elu'(in[i][k]) ----
        const bool wsmem = net.smem_off[l] >= 0;
        const float* W = wsmem ? wsm + net.smem_off[l] : packed + net.w_off[l];
        for (int e = tid; e < G_TB * K; e += (int)blockDim.x) {
            const int i = e / K, k = e - i * K;
            const float* dr = dcur + i * DM;
            float acc = 0.0f;
            if (wsmem) {
                for (int n = 0; n < N; n++) acc = fmaf(W[(size_t)n * kp + k], dr[n], acc);
            } else {
                for (int n = 0; n < N; n++) acc = fmaf(__ldg(W + (size_t)n * kp + k), dr[n], acc);
            }
            if (l > 0) acc *= psn_elu_grad_from_out(in[i * ins + k]);
            dnext[i * DM + k] = acc;
        }
        __syncthreads();
        float* t = dcur; dcur = dnext; dnext = t;
    }
    return dcur;
}

template <bool DAE>
__global__ void __launch_bounds__(R_NT_MAX) psn_generic_bwd_kernel(const __grid_constant__ BwdParams q) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const psnode_problem& p = q.p;
    const psnode_adjoint& a = q.a;
    const int tid = threadIdx.x;
    const int B = p.B, T = p.T, X = p.X, Z = p.Z, V = p.V, I = p.I, S = q.S, U = q.U;
    const int S4 = q.S4, K0 = q.K0, KA0 = q.KA0, X4 = q.X4, I4 = q.I4, Z4 = q.Z4, V4 = q.V4, U4 = q.U4, HM = q.HM, DM = q.DM;
    const int nst = q.nst;
    float* wsm = sm + q.o_w;
    float* a0s = sm + q.o_a0;
    float* da0 = sm + q.o_da0;
    float* u3 = sm + q.o_u3;
    float* uae = sm + q.o_uae;
    float* ys = sm + q.o_ys;          // [nst][TB][X4]
    float* acts = sm + q.o_acts;      // [nst][n_de-1][TB][HM]
    float* aeacts = sm + q.o_aeacts;  // [n_ae-1][TB][HM]
    float* kst = sm + q.o_kst;        // [nst][TB][X4]
    float* dk = sm + q.o_dk;          // [nst][TB][X4]
    float* dA = sm + q.o_dA;
    float* dB = sm + q.o_dB;
    float* lam = sm + q.o_lam;
    float* mu = sm + q.o_mu;
    float* dxs = sm + q.o_dxs;
    float* dy = sm + q.o_dy;
    float* pz = sm + q.o_pz;
    float* pv = sm + q.o_pv;
    float* pxt = sm + q.o_pxt;
    float* pit = sm + q.o_pit;
    float* dheld = sm + q.o_dheld;
    float* ihold = sm + q.o_ihold;
    float* start = sm + q.o_start;
    float* itmp = sm + q.o_itmp;
    float* dts = sm + q.o_dt;
    float* slab = q.slab + (size_t)blockIdx.x * q.packed_floats;
    const int acts_stage = (q.de.n_layers - 1) * G_TB * HM;

    // Butcher tableau of the scheme (my_fixed_grid.py): y_s = start + dt * sum_r A[s][r] k_r ; x1 = start + dt * sum_s bw[s] k_s
    float A[R_MAXST][R_MAXST] = {};
    float bw[R_MAXST] = {};
    const float c13 = (float)(1.0 / 3.0);
    if (p.method == PSNODE_EULER) bw[0] = 1.0f;
    else if (p.method == PSNODE_MIDPOINT) { A[1][0] = 0.5f; bw[1] = 1.0f; }
    else { A[1][0] = c13; A[2][0] = -c13; A[2][1] = 1.0f; A[3][0] = 1.0f; A[3][1] = -1.0f; A[3][2] = 1.0f;
           bw[0] = 0.125f; bw[1] = 0.375f; bw[2] = 0.375f; bw[3] = 0.125f; }

    // ---- one-time: weights -> smem, zero the gradient accumulators ----------------------------------------
    for (int net = 0; net < (DAE ? 2 : 1); net++) {
        const PsnPackedNet& pn = net ? q.ae : q.de;
        for (int l = 0; l < pn.n_layers; l++) {
            if (pn.smem_off[l] < 0) continue;
            const int n4 = pn.out_dim[l] * pn.kpad[l] / 4;
            const float4* g = reinterpret_cast<const float4*>(q.packed + pn.w_off[l]);
            float4* s = reinterpret_cast<float4*>(wsm + pn.smem_off[l]);
            for (int e = tid; e < n4; e += (int)blockDim.x) s[e] = __ldg(g + e);
        }
    }
    for (int e = q.g_begin + tid; e < q.g_end; e += (int)blockDim.x) sm[e] = 0.0f;

    auto set_state = [&](int i, int c, float xv) {
        u3[i * K0 + S + c] = __fsub_rn(xv, a0s[i * S4 + c]);
        u3[i * K0 + 2 * S + c] = xv;
    };
    // AE input vector: x part from smem rows (stride X4) or the teacher series at grid point jx; z/v of grid point jz or event k
    auto ae_forward = [&](int b0, const float* xsrc, int jx_teacher, int jz, int k, float* out) {
        const int W3 = X + Z + V;
        for (int e = tid; e < G_TB * W3; e += (int)blockDim.x) {
            const int i = e / W3, c = e - i * W3;
            const int bb = min(b0 + i, B - 1);
            float val;
            if (c < X) val = (jx_teacher >= 0) ? ld_series(p.x, jx_teacher, bb, c) : xsrc[i * X4 + c];
            else if (c < X + Z) {
                const int cz = c - X;
                val = (k >= 0) ? __ldg(p.z_jump + (int64_t)bb * p.zj_sb + (int64_t)k * p.zj_se + cz) : ld_series(p.z, jz, bb, cz);
            } else {
                const int cv = c - X - Z;
                val = (k >= 0) ? __ldg(p.v_jump + (int64_t)bb * p.vj_sb + (int64_t)k * p.vj_se + cv) : ld_series(p.v, jz, bb, cv);
            }
            uae[i * KA0 + S + c] = val;
        }
        __syncthreads();
        run_mlp_keep(q.ae, q.packed, wsm, uae, KA0, out, I4, I, aeacts, HM);
    };

    for (int tile = blockIdx.x; tile < q.n_tiles; tile += gridDim.x) {
        const int b0 = tile * G_TB;
        __syncthreads();
        for (int e = q.zero_begin + tid; e < q.zero_end; e += (int)blockDim.x) sm[e] = 0.0f;
        __syncthreads();
        for (int e = tid; e < G_TB * S; e += (int)blockDim.x) {
            const int i = e / S, c = e - i * S;
            const int bb = min(b0 + i, B - 1);
            const float av = __ldg(p.a0 + (int64_t)bb * p.a0_sb + c);
            a0s[i * S4 + c] = av;
            u3[i * K0 + c] = av;
            if (DAE) uae[i * KA0 + c] = av;
        }
        // jump gradients are accumulated (+=) over the steps that fire the same event: clear this tile's rows first
        if (p.event_idx) {
            if (a.d_zjump && Z > 0)
                for (int e = tid; e < G_TB * p.E * Z; e += (int)blockDim.x) {
                    const int i = e / (p.E * Z), r = e - i * (p.E * Z), b = b0 + i;
                    if (b < B) a.d_zjump[(int64_t)b * a.d_zj_sb + (int64_t)(r / Z) * a.d_zj_se + (r % Z)] = 0.0f;
                }
            if (DAE && a.d_vjump && V > 0)
                for (int e = tid; e < G_TB * p.E * V; e += (int)blockDim.x) {
                    const int i = e / (p.E * V), r = e - i * (p.E * V), b = b0 + i;
                    if (b < B) a.d_vjump[(int64_t)b * a.d_vj_sb + (int64_t)(r / V) * a.d_vj_se + (r % V)] = 0.0f;
                }
        }
        __syncthreads();

        for (int j = T - 1; j >= 0; j--) {
            // ================= point j =================
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, b = b0 + i;
                if (b < B && a.gx.p) lam[i * X4 + c] += ld_series(a.gx, j, b, c);
            }
            if (DAE) {
                for (int e = tid; e < G_TB * I; e += (int)blockDim.x) {
                    const int i = e / I, c = e - i * I, b = b0 + i;
                    if (b < B && a.gi.p) mu[i * I4 + c] += ld_series(a.gi, j, b, c);
                }
                // x_j rows for the AE input (x_sol[j]; for j = 0 this is x_init, written by the forward kernel)
                if (!p.teacher_x)
                    for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                        const int i = e / X, c = e - i * X, bb = min(b0 + i, B - 1);
                        start[i * X4 + c] = __ldg(p.x_sol.p + (int64_t)j * p.x_sol.st + (int64_t)bb * p.x_sol.sb + c);
                    }
                __syncthreads();
                ae_forward(b0, start, p.teacher_x ? j : -1, j, -1, itmp);
                for (int e = tid; e < G_TB * I; e += (int)blockDim.x) { const int i = e / I, c = e - i * I; dA[i * DM + c] = mu[i * I4 + c]; }
                __syncthreads();
                const float* du = mlp_backward(q.ae, q.packed, wsm, sm, q.gw_ae, q.gb_ae, slab, uae, KA0, aeacts, HM, dA, dB, DM);
                for (int e = tid; e < G_TB * KA0; e += (int)blockDim.x) {
                    const int i = e / KA0, c = e - i * KA0;
                    if (c >= S + X + Z + V) continue;
                    const float g = du[i * DM + c];
                    if (c < S) da0[i * S4 + c] += g;
                    else if (c < S + X) { if (p.teacher_x) pxt[i * X4 + (c - S)] += g; else lam[i * X4 + (c - S)] += g; }
                    else if (c < S + X + Z) pz[i * Z4 + (c - S - X)] += g;
                    else pv[i * V4 + (c - S - X - Z)] += g;
                }
            }
            __syncthreads();
            // gradients of the input series at grid point j are complete
            if (a.d_z.p)
                for (int e = tid; e < G_TB * Z; e += (int)blockDim.x) {
                    const int i = e / Z, c = e - i * Z, b = b0 + i;
                    if (b < B) a.d_z.p[(int64_t)j * a.d_z.st + (int64_t)b * a.d_z.sb + c] = pz[i * Z4 + c];
                }
            if (DAE && a.d_v.p)
                for (int e = tid; e < G_TB * V; e += (int)blockDim.x) {
                    const int i = e / V, c = e - i * V, b = b0 + i;
                    if (b < B) a.d_v.p[(int64_t)j * a.d_v.st + (int64_t)b * a.d_v.sb + c] = pv[i * V4 + c];
                }
            if (a.d_xteach.p)
                for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                    const int i = e / X, c = e - i * X, b = b0 + i;
                    if (b < B) a.d_xteach.p[(int64_t)j * a.d_xteach.st + (int64_t)b * a.d_xteach.sb + c] = pxt[i * X4 + c];
                }
            if (DAE && a.d_iteach.p)
                for (int e = tid; e < G_TB * I; e += (int)blockDim.x) {
                    const int i = e / I, c = e - i * I, b = b0 + i;
                    if (b < B) a.d_iteach.p[(int64_t)j * a.d_iteach.st + (int64_t)b * a.d_iteach.sb + c] = pit[i * I4 + c];
                }
            if (j == 0) break;
            __syncthreads();

            // ================= step j: recompute =================
            const int k = p.event_idx ? __ldg(p.event_idx + (j - 1)) : -1;
            for (int i = tid; i < G_TB; i += (int)blockDim.x) {
                const int bb = min(b0 + i, B - 1);
                dts[i] = __fsub_rn(ld_series(p.t, j, bb, 0), ld_series(p.t, j - 1, bb, 0));
            }
            // predicted previous state (needed for the event AE even under teacher forcing)
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, bb = min(b0 + i, B - 1);
                dy[i * X4 + c] = __ldg(p.x_sol.p + (int64_t)(j - 1) * p.x_sol.st + (int64_t)bb * p.x_sol.sb + c);
            }
            __syncthreads();
            const bool event_ae = DAE && k >= 0 && !p.teacher_i;
            if (event_ae) ae_forward(b0, dy, -1, j - 1, k, ihold);   // i_0 from the jumped inputs (my_solvers.py:109-110)
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, bb = min(b0 + i, B - 1);
                const float xv = p.teacher_x ? ld_series(p.x, j - 1, bb, c) : dy[i * X4 + c];
                start[i * X4 + c] = xv;
                ys[i * X4 + c] = xv;
                set_state(i, c, xv);
            }
            for (int e = tid; e < G_TB * U; e += (int)blockDim.x) {
                const int i = e / U, c = e - i * U, bb = min(b0 + i, B - 1);
                float hv;
                if (c < Z) hv = (k >= 0) ? __ldg(p.z_jump + (int64_t)bb * p.zj_sb + (int64_t)k * p.zj_se + c) : ld_series(p.z, j - 1, bb, c);
                else if (c < Z + V) {
                    const int cv = c - Z;
                    hv = (k >= 0) ? __ldg(p.v_jump + (int64_t)bb * p.vj_sb + (int64_t)k * p.vj_se + cv) : ld_series(p.v, j - 1, bb, cv);
                } else {
                    const int ci = c - Z - V;
                    if (p.teacher_i) hv = ld_series(p.i, j - 1, bb, ci);
                    else if (event_ae) hv = ihold[i * I4 + ci];
                    else hv = __ldg(p.i_sol.p + (int64_t)(j - 1) * p.i_sol.st + (int64_t)bb * p.i_sol.sb + ci);
                }
                u3[i * K0 + S + X + c] = __fsub_rn(hv, a0s[i * S4 + X + c]);
                u3[i * K0 + 2 * S + X + c] = hv;
                dheld[i * U4 + c] = 0.0f;
            }
            __syncthreads();
            for (int s = 0; s < nst; s++) {
                if (s > 0) {   // y_s, same operation order as the forward kernels
                    for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                        const int i = e / X, c = e - i * X, o = i * X4 + c;
                        const float dt = dts[i];
                        const float* k1 = kst; const float* k2 = kst + G_TB * X4; const float* k3 = kst + 2 * G_TB * X4;
                        float yv;
                        if (p.method == PSNODE_MIDPOINT) yv = __fadd_rn(start[o], __fmul_rn(k1[o], __fmul_rn(0.5f, dt)));
                        else if (s == 1) yv = __fadd_rn(start[o], __fmul_rn(__fmul_rn(dt, k1[o]), c13));
                        else if (s == 2) yv = __fadd_rn(start[o], __fmul_rn(dt, __fsub_rn(k2[o], __fmul_rn(k1[o], c13))));
                        else yv = __fadd_rn(start[o], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[o], k2[o]), k3[o])));
                        ys[s * G_TB * X4 + o] = yv;
                        set_state(i, c, yv);
                    }
                    __syncthreads();
                }
                run_mlp_keep(q.de, q.packed, wsm, u3, K0, kst + s * G_TB * X4, X4, X, acts + s * acts_stage, HM);
            }
            // ================= step j: reverse =================
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                const float l = lam[o], dt = dts[i];
                dxs[o] = l;
                for (int s = 0; s < nst; s++) dk[s * G_TB * X4 + o] = l * dt * bw[s];
            }
            __syncthreads();
            for (int s = nst - 1; s >= 0; s--) {
                for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                    const int i = e / X, c = e - i * X, o = i * X4 + c;
                    set_state(i, c, ys[s * G_TB * X4 + o]);
                    dA[i * DM + c] = dk[s * G_TB * X4 + o];
                }
                __syncthreads();
                const float* du = mlp_backward(q.de, q.packed, wsm, sm, q.gw_de, q.gb_de, slab, u3, K0, acts + s * acts_stage, HM, dA, dB, DM);
                for (int e = tid; e < G_TB * S; e += (int)blockDim.x) {
                    const int i = e / S, c = e - i * S;
                    const float ga = du[i * DM + c], gb = du[i * DM + S + c], gc = du[i * DM + 2 * S + c];
                    da0[i * S4 + c] += ga - gb;
                    const float gs = gb + gc;
                    if (c < X) {
                        const int o = i * X4 + c;
                        const float dt = dts[i];
                        dxs[o] += gs;
                        for (int r = 0; r < s; r++) dk[r * G_TB * X4 + o] += dt * A[s][r] * gs;
                    } else dheld[i * U4 + (c - X)] += gs;
                }
                __syncthreads();
            }
            // ================= route the step's input gradients =================
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                if (p.teacher_x) { pxt[o] = dxs[o]; lam[o] = 0.0f; } else { lam[o] = dxs[o]; pxt[o] = 0.0f; }
            }
            for (int e = tid; e < G_TB * U; e += (int)blockDim.x) {
                const int i = e / U, c = e - i * U, b = b0 + i;
                const float g = dheld[i * U4 + c];
                if (c < Z) {
                    if (k >= 0) { pz[i * Z4 + c] = 0.0f; if (a.d_zjump && b < B) a.d_zjump[(int64_t)b * a.d_zj_sb + (int64_t)k * a.d_zj_se + c] += g; }
                    else pz[i * Z4 + c] = g;
                } else if (c < Z + V) {
                    const int cv = c - Z;
                    if (k >= 0) { pv[i * V4 + cv] = 0.0f; if (a.d_vjump && b < B) a.d_vjump[(int64_t)b * a.d_vj_sb + (int64_t)k * a.d_vj_se + cv] += g; }
                    else pv[i * V4 + cv] = g;
                } else {
                    const int ci = c - Z - V;
                    if (p.teacher_i) { pit[i * I4 + ci] = g; mu[i * I4 + ci] = 0.0f; }
                    else if (event_ae) { dA[i * DM + ci] = g; mu[i * I4 + ci] = 0.0f; pit[i * I4 + ci] = 0.0f; }
                    else { mu[i * I4 + ci] = g; pit[i * I4 + ci] = 0.0f; }
                }
            }
            __syncthreads();
            if (event_ae) {   // back through i_0 = ae(x_{j-1}, z_jump[k], v_jump[k]); uae / aeacts still hold that evaluation
                const float* du = mlp_backward(q.ae, q.packed, wsm, sm, q.gw_ae, q.gb_ae, slab, uae, KA0, aeacts, HM, dA, dB, DM);
                for (int e = tid; e < G_TB * KA0; e += (int)blockDim.x) {
                    const int i = e / KA0, c = e - i * KA0, b = b0 + i;
                    if (c >= S + X + Z + V) continue;
                    const float g = du[i * DM + c];
                    if (c < S) da0[i * S4 + c] += g;
                    else if (c < S + X) lam[i * X4 + (c - S)] += g;
                    else if (c < S + X + Z) { if (a.d_zjump && b < B) a.d_zjump[(int64_t)b * a.d_zj_sb + (int64_t)k * a.d_zj_se + (c - S - X)] += g; }
                    else { if (a.d_vjump && b < B) a.d_vjump[(int64_t)b * a.d_vj_sb + (int64_t)k * a.d_vj_se + (c - S - X - Z)] += g; }
                }
                __syncthreads();
            }
        }
        // ---- tile epilogue: gradients of the initial state and of all_initial ----
        __syncthreads();
        if (a.d_x0)
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, b = b0 + i;
                if (b < B) a.d_x0[(int64_t)b * a.d_x0_sb + c] = lam[i * X4 + c];
            }
        if (a.d_a0)
            for (int e = tid; e < G_TB * S; e += (int)blockDim.x) {
                const int i = e / S, c = e - i * S, b = b0 + i;
                if (b < B) a.d_a0[(int64_t)b * a.d_a0_sb + c] = da0[i * S4 + c];
            }
    }
    // ---- flush the shared-memory accumulators to this CTA's slab (packed layout) ----
    __syncthreads();
    for (int net = 0; net < (DAE ? 2 : 1); net++) {
        const PsnPackedNet& pn = net ? q.ae : q.de;
        const int* gw = net ? q.gw_ae : q.gw_de;
        const int* gb = net ? q.gb_ae : q.gb_de;
        for (int l = 0; l < pn.n_layers; l++) {
            if (gw[l] >= 0)
                for (int e = tid; e < pn.out_dim[l] * pn.kpad[l]; e += (int)blockDim.x) slab[pn.w_off[l] + e] = sm[gw[l] + e];
            for (int e = tid; e < pn.out_dim[l]; e += (int)blockDim.x) slab[pn.b_off[l] + e] = sm[gb[l] + e];
        }
    }
}

// d_theta (reference parameter order: for net in (de, ae): for layer: W row-major, then b) = sum over the CTA slabs
struct ReduceDesc {
    int n;
    int theta_off[2 * G_MAXNETLAYERS], packed_off[2 * G_MAXNETLAYERS], rows[2 * G_MAXNETLAYERS], cols[2 * G_MAXNETLAYERS],
        kpad[2 * G_MAXNETLAYERS];
};

__global__ void psn_grad_reduce_kernel(const __grid_constant__ ReduceDesc d, const float* __restrict__ slab, int n_slabs,
                                       int packed_floats, float* __restrict__ d_theta) {
    for (int part = blockIdx.y; part < d.n; part += gridDim.y) {
        const int rows = d.rows[part], cols = d.cols[part], kp = d.kpad[part];
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < rows * cols; e += gridDim.x * blockDim.x) {
            const int r = e / cols, c = e - r * cols;
            const float* src = slab + d.packed_off[part] + (size_t)r * kp + c;
            float acc = 0.0f;
            for (int s = 0; s < n_slabs; s++) acc += src[(size_t)s * packed_floats];
            d_theta[d.theta_off[part] + e] = acc;
        }
    }
}

int bwd_grid(const psnode_problem* p) {
    const int tiles = (p->B + G_TB - 1) / G_TB;
    return tiles < 592 ? tiles : 592;      // 148 SMs x 4: bounds the slab workspace, tiles are looped over
}

int build_bwd_params(const psnode_problem* p, const psnode_adjoint* a, BwdParams& q, int max_smem_bytes) {
    q.p = *p;
    q.a = *a;
    const bool dae = p->kind == PSNODE_DAE;
    int cursor = 0;
    pack_layout(p->de, q.de, cursor);
    if (dae) pack_layout(p->ae, q.ae, cursor);
    else q.ae.n_layers = 0;
    q.packed_floats = cursor;
    q.n_tiles = (p->B + G_TB - 1) / G_TB;
    q.nst = p->method == PSNODE_EULER ? 1 : (p->method == PSNODE_MIDPOINT ? 2 : 4);
    q.S = psn_S(p);
    q.S4 = psn_pad4(q.S);
    q.X4 = psn_pad4(p->X);
    q.I4 = psn_pad4(p->I > 0 ? p->I : 1);
    q.Z4 = psn_pad4(p->Z > 0 ? p->Z : 1);
    q.V4 = psn_pad4(p->V > 0 ? p->V : 1);
    q.U = q.S - p->X;
    q.U4 = psn_pad4(q.U > 0 ? q.U : 1);
    q.K0 = q.de.kpad[0];
    q.KA0 = dae ? q.ae.kpad[0] : 4;
    int hm = 4;
    for (int l = 1; l < q.de.n_layers; l++) hm = hm > q.de.kpad[l] ? hm : q.de.kpad[l];
    for (int l = 1; l < q.ae.n_layers; l++) hm = hm > q.ae.kpad[l] ? hm : q.ae.kpad[l];
    q.HM = hm;
    int dm = hm > q.K0 ? hm : q.K0;
    dm = dm > q.KA0 ? dm : q.KA0;
    dm = dm > q.X4 ? dm : q.X4;
    dm = dm > q.I4 ? dm : q.I4;
    q.DM = dm;
    int o = 0;
    auto take = [&](int n) { int r = o; o += psn_pad4(n > 0 ? n : 1); return r; };
    // per-tile state that must start at zero
    q.zero_begin = o;
    q.o_da0 = take(G_TB * q.S4);
    q.o_u3 = take(G_TB * q.K0);
    q.o_uae = take(dae ? G_TB * q.KA0 : 4);
    q.o_lam = take(G_TB * q.X4);
    q.o_mu = take(G_TB * q.I4);
    q.o_pz = take(G_TB * q.Z4);
    q.o_pv = take(G_TB * q.V4);
    q.o_pxt = take(G_TB * q.X4);
    q.o_pit = take(G_TB * q.I4);
    q.o_acts = take(q.nst * (q.de.n_layers - 1) * G_TB * q.HM);
    q.o_aeacts = take(dae ? (q.ae.n_layers - 1) * G_TB * q.HM : 4);
    q.o_dA = take(G_TB * q.DM);
    q.o_dB = take(G_TB * q.DM);
    q.zero_end = o;
    q.o_a0 = take(G_TB * q.S4);
    q.o_ys = take(q.nst * G_TB * q.X4);
    q.o_kst = take(q.nst * G_TB * q.X4);
    q.o_dk = take(q.nst * G_TB * q.X4);
    q.o_dxs = take(G_TB * q.X4);
    q.o_dy = take(G_TB * q.X4);
    q.o_dheld = take(G_TB * q.U4);
    q.o_ihold = take(G_TB * q.I4);
    q.o_start = take(G_TB * q.X4);
    q.o_itmp = take(G_TB * q.I4);
    q.o_dt = take(G_TB);
    // bias-gradient accumulators (always shared memory)
    q.g_begin = o;
    for (int l = 0; l < q.de.n_layers; l++) q.gb_de[l] = take(q.de.out_dim[l]);
    for (int l = 0; l < q.ae.n_layers; l++) q.gb_ae[l] = take(q.ae.out_dim[l]);
    int budget = max_smem_bytes / 4 - o;
    if (budget < 0) return PSNODE_EUNSUPPORTED;
    // greedy: smallest items first; at equal size the gradient accumulator (read-modify-write) before the weights (read)
    struct Item { int net, l, sz, grad; } items[2 * G_MAXNETLAYERS];
    int n = 0;
    for (int l = 0; l < q.de.n_layers; l++) { q.gw_de[l] = -1; int sz = q.de.out_dim[l] * q.de.kpad[l]; items[n++] = {0, l, sz, 1}; items[n++] = {0, l, sz, 0}; }
    for (int l = 0; l < q.ae.n_layers; l++) { q.gw_ae[l] = -1; int sz = q.ae.out_dim[l] * q.ae.kpad[l]; items[n++] = {1, l, sz, 1}; items[n++] = {1, l, sz, 0}; }
    for (int x = 0; x < n; x++)
        for (int y = x + 1; y < n; y++)
            if (items[y].sz < items[x].sz || (items[y].sz == items[x].sz && items[y].grad > items[x].grad)) { Item t = items[x]; items[x] = items[y]; items[y] = t; }
    // gradient accumulators are laid out right after the bias accumulators so one loop zeroes both
    int used_g = 0, used_w = 0;
    bool place[2 * G_MAXNETLAYERS];
    for (int x = 0; x < n; x++) {
        place[x] = used_g + used_w + items[x].sz <= budget;
        if (place[x]) { if (items[x].grad) used_g += items[x].sz; else used_w += items[x].sz; }
    }
    int go = o, wo = o + used_g;
    q.o_g = go;
    q.o_w = wo;
    int wcur = 0;
    for (int x = 0; x < n; x++) {
        if (!place[x]) continue;
        if (items[x].grad) { (items[x].net ? q.gw_ae : q.gw_de)[items[x].l] = go; go += items[x].sz; }
        else { (items[x].net ? q.ae : q.de).smem_off[items[x].l] = wcur; wcur += items[x].sz; }
    }
    q.g_end = o + used_g;
    return (o + used_g + used_w) * 4;
}

}  // namespace

int64_t PSN_G_NAME(psn_generic_backward_workspace)(const psnode_problem* p, const psnode_adjoint*) {
    int cursor = 0;
    PsnPackedNet de, ae;
    pack_layout(p->de, de, cursor);
    if (p->kind == PSNODE_DAE) pack_layout(p->ae, ae, cursor);
    return (int64_t)cursor * 4 * (1 + bwd_grid(p));
}

int PSN_G_NAME(psn_generic_backward)(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    static BwdParams q;   // large POD; single-caller library (no re-entrancy)
    int dev = 0, max_smem = 0;
    PSN_CUDA(cudaGetDevice(&dev));
    PSN_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int smem_bytes = build_bwd_params(p, a, q, max_smem);
    if (smem_bytes < 0) return smem_bytes;
    const int grid = bwd_grid(p);
    const int64_t need = (int64_t)q.packed_floats * 4 * (1 + grid);
    if (ws == nullptr || ws_bytes < need) return PSNODE_EWORKSPACE;
    const bool dae = p->kind == PSNODE_DAE;
    if (a->n_theta != psnode_mlp_param_count(&p->de) + (dae ? psnode_mlp_param_count(&p->ae) : 0)) return PSNODE_EINVAL;
    float* packed = static_cast<float*>(ws);
    q.packed = packed;
    q.slab = packed + q.packed_floats;

    PackDesc d;
    ReduceDesc rd;
    d.n = 0;
    rd.n = 0;
    int theta = 0;
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        const psnode_mlp& m = net ? p->ae : p->de;
        const PsnPackedNet& pn = net ? q.ae : q.de;
        for (int l = 0; l < m.n_layers; l++) {
            d.W[d.n] = m.W[l]; d.b[d.n] = m.b[l];
            d.in[d.n] = m.in_dim[l]; d.out[d.n] = m.out_dim[l]; d.kpad[d.n] = pn.kpad[l];
            d.w_off[d.n] = pn.w_off[l]; d.b_off[d.n] = pn.b_off[l];
            d.n++;
            rd.theta_off[rd.n] = theta; rd.packed_off[rd.n] = pn.w_off[l]; rd.rows[rd.n] = m.out_dim[l]; rd.cols[rd.n] = m.in_dim[l];
            rd.kpad[rd.n] = pn.kpad[l]; rd.n++;
            theta += m.out_dim[l] * m.in_dim[l];
            rd.theta_off[rd.n] = theta; rd.packed_off[rd.n] = pn.b_off[l]; rd.rows[rd.n] = 1; rd.cols[rd.n] = m.out_dim[l];
            rd.kpad[rd.n] = m.out_dim[l]; rd.n++;
            theta += m.out_dim[l];
        }
    }
    psn_pack_kernel<<<dim3(32, d.n), 256, 0, stream>>>(d, packed);
    psn_count_launch("psn_pack_kernel");
    PSN_CUDA(cudaGetLastError());
    PSN_CUDA(cudaMemsetAsync(q.slab, 0, (size_t)q.packed_floats * 4 * grid, stream));
    // nets whose weights or gradient accumulators do not fit shared memory stream them through L2: hide that latency with 4x the warps
    bool streamed = false;
    for (int l = 0; l < q.de.n_layers; l++) streamed |= q.de.smem_off[l] < 0 || q.gw_de[l] < 0;
    for (int l = 0; l < q.ae.n_layers; l++) streamed |= q.ae.smem_off[l] < 0 || q.gw_ae[l] < 0;
    const int nthreads = streamed ? R_NT_MAX : 256;
    if (dae) {
        PSN_CUDA(cudaFuncSetAttribute(psn_generic_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        psn_generic_bwd_kernel<true><<<grid, nthreads, smem_bytes, stream>>>(q);
        psn_count_launch("psn_generic_bwd_kernel<dae>");
    } else {
        PSN_CUDA(cudaFuncSetAttribute(psn_generic_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        psn_generic_bwd_kernel<false><<<grid, nthreads, smem_bytes, stream>>>(q);
        psn_count_launch("psn_generic_bwd_kernel<ode>");
    }
    PSN_CUDA(cudaGetLastError());
    psn_grad_reduce_kernel<<<dim3(16, rd.n), 256, 0, stream>>>(rd, q.slab, grid, q.packed_floats, a->d_theta);
    psn_count_launch("psn_grad_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
