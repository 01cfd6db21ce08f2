// psnode_generic_fwd.cu -- generic forward integrator: any widths, any layer count, ODE and DAE, all three
// schemes, events and teacher forcing.  One persistent CTA per tile of G_TB trajectories walks the whole time
// grid; weights live in shared memory (layers that do not fit are streamed from L2), the layer-1 input vector
// cat(a0, s - a0, s) is materialised exactly as the reference builds it (neural_00_ODE_01_no_encode.py:66-68),
// so this kernel is the "reference formulation" on the GPU.  The fused kernel (psnode_fused_fwd.cu) is the fast
// path for the H = 64 nets; this one is the fallback for every other shape and its on-device cross-check.
//
// Replaces: FixedGridODESolver.integrate_ODE / integrate_DAE (neural_dae/my_solvers.py:52-80, 82-131),
//           step_integrate (:48-50), Euler/Midpoint/RK4._step_func (neural_dae/my_fixed_grid.py:15-59),
//           DE_Func.forward / AE_Func.forward (script-local, see include/psnode_b200.h).
#include "psnode_internal.cuh"

#include "psnode_generic.cuh"

namespace {

template <bool DAE>
__global__ void __launch_bounds__(G_NT_MAX) psn_generic_fwd_kernel(const __grid_constant__ GenericParams q) {
    extern __shared__ float4 smem4[];
    float* sm = reinterpret_cast<float*>(smem4);
    const psnode_problem& p = q.p;
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * G_TB;
    const int B = p.B, X = p.X, Z = p.Z, V = p.V, I = p.I, S = q.S;
    const int S4 = q.S4, K0 = q.K0, KA0 = q.KA0, X4 = q.X4, I4 = q.I4, HM = q.HM;
    float* wsm = sm + q.o_w;
    float* a0s = sm + q.o_a0;
    float* u3 = sm + q.o_u3;
    float* uae = sm + q.o_uae;
    float* actA = sm + q.o_actA;
    float* actB = sm + q.o_actB;
    float* xprev = sm + q.o_xprev;
    float* start = sm + q.o_start;
    float* k1 = sm + q.o_k1;
    float* k2 = sm + q.o_k2;
    float* k3 = sm + q.o_k3;
    float* k4 = sm + q.o_k4;
    float* iprev = sm + q.o_iprev;
    float* dts = sm + q.o_dt;

    // ---- prologue: weights -> smem, zero the buffers, per-trajectory constants -------------------------
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
    for (int e = q.buf_begin + tid; e < q.total_floats; e += (int)blockDim.x) sm[e] = 0.0f;
    __syncthreads();
    for (int e = tid; e < G_TB * S; e += (int)blockDim.x) {
        const int i = e / S, c = e - i * S;
        const int bb = min(b0 + i, B - 1);
        const float a = __ldg(p.a0 + (int64_t)bb * p.a0_sb + c);
        a0s[i * S4 + c] = a;
        u3[i * K0 + c] = a;
        if (DAE) uae[i * KA0 + c] = a;
    }
    for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
        const int i = e / X, c = e - i * X;
        const int b = b0 + i, bb = min(b, B - 1);
        const float xv = DAE ? __ldg(p.x_init + (int64_t)bb * p.x_init_sb + c) : ld_series(p.x, 0, bb, c);
        xprev[i * X4 + c] = xv;
        if (b < B) p.x_sol.p[(int64_t)b * p.x_sol.sb + c] = xv;
    }
    __syncthreads();

    // AE evaluation helper: x source (smem, stride X4) or teacher series at grid point jx; z/v from grid point jz or event k
    auto ae_eval = [&](const float* xsrc, int jx_teacher, int jz, int k) {
        for (int e = tid; e < G_TB * (X + Z + V); e += (int)blockDim.x) {
            const int i = e / (X + Z + V), c = e - i * (X + Z + V);
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
        run_mlp(q.ae, q.packed, wsm, uae, KA0, iprev, I4, I, actA, actB, HM);
    };
    auto store_i = [&](int j) {
        for (int e = tid; e < G_TB * I; e += (int)blockDim.x) {
            const int i = e / I, c = e - i * I;
            const int b = b0 + i;
            if (b < B) p.i_sol.p[(int64_t)j * p.i_sol.st + (int64_t)b * p.i_sol.sb + c] = iprev[i * I4 + c];
        }
    };
    // write the x part of the DE input vector: (s - a0) block and s block
    auto set_state = [&](int i, int c, float xv) {
        u3[i * K0 + S + c] = __fsub_rn(xv, a0s[i * S4 + c]);
        u3[i * K0 + 2 * S + c] = xv;
    };
    auto rhs = [&](float* kout) {
        __syncthreads();
        run_mlp(q.de, q.packed, wsm, u3, K0, kout, X4, X, actA, actB, HM);
    };

    if (DAE) {   // i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95)
        ae_eval(xprev, p.teacher_x ? 0 : -1, 0, -1);
        store_i(0);
    }

    const float c13 = (float)(1.0 / 3.0);
    const int T = p.T;
    for (int j = 1; j < T; j++) {
        __syncthreads();
        const int k = p.event_idx ? __ldg(p.event_idx + (j - 1)) : -1;
        if (DAE && k >= 0) ae_eval(xprev, -1, j - 1, k);   // i_0 recomputed from the jumped inputs (my_solvers.py:109-110)
        // per-step inputs: dt, start state, held inputs
        for (int i = tid; i < G_TB; i += (int)blockDim.x) {
            const int bb = min(b0 + i, B - 1);
            dts[i] = __fsub_rn(ld_series(p.t, j, bb, 0), ld_series(p.t, j - 1, bb, 0));
        }
        for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
            const int i = e / X, c = e - i * X;
            const int bb = min(b0 + i, B - 1);
            const float xv = p.teacher_x ? ld_series(p.x, j - 1, bb, c) : xprev[i * X4 + c];
            start[i * X4 + c] = xv;
            set_state(i, c, xv);
        }
        const int U = S - X;
        for (int e = tid; e < G_TB * U; e += (int)blockDim.x) {
            const int i = e / U, c = e - i * U;
            const int bb = min(b0 + i, B - 1);
            float hv;
            if (c < Z) hv = (k >= 0) ? __ldg(p.z_jump + (int64_t)bb * p.zj_sb + (int64_t)k * p.zj_se + c) : ld_series(p.z, j - 1, bb, c);
            else if (c < Z + V) {
                const int cv = c - Z;
                hv = (k >= 0) ? __ldg(p.v_jump + (int64_t)bb * p.vj_sb + (int64_t)k * p.vj_se + cv) : ld_series(p.v, j - 1, bb, cv);
            } else {
                const int ci = c - Z - V;
                hv = p.teacher_i ? ld_series(p.i, j - 1, bb, ci) : iprev[i * I4 + ci];
            }
            u3[i * K0 + S + X + c] = __fsub_rn(hv, a0s[i * S4 + X + c]);
            u3[i * K0 + 2 * S + X + c] = hv;
        }
        rhs(k1);
        if (p.method == PSNODE_EULER) {
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                xprev[o] = __fadd_rn(start[o], __fmul_rn(dts[i], k1[o]));
            }
        } else if (p.method == PSNODE_MIDPOINT) {
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                const float half_dt = __fmul_rn(0.5f, dts[i]);
                set_state(i, c, __fadd_rn(start[o], __fmul_rn(k1[o], half_dt)));
            }
            rhs(k2);
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                xprev[o] = __fadd_rn(start[o], __fmul_rn(dts[i], k2[o]));
            }
        } else {   // RK4, 3/8 rule; operation order of rk4_alt_step_func (my_fixed_grid.py:38-51)
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                set_state(i, c, __fadd_rn(start[o], __fmul_rn(__fmul_rn(dts[i], k1[o]), c13)));
            }
            rhs(k2);
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                set_state(i, c, __fadd_rn(start[o], __fmul_rn(dts[i], __fsub_rn(k2[o], __fmul_rn(k1[o], c13)))));
            }
            rhs(k3);
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                set_state(i, c, __fadd_rn(start[o], __fmul_rn(dts[i], __fadd_rn(__fsub_rn(k1[o], k2[o]), k3[o]))));
            }
            rhs(k4);
            for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
                const int i = e / X, c = e - i * X, o = i * X4 + c;
                const float ks = __fadd_rn(__fadd_rn(k1[o], __fmul_rn(3.0f, __fadd_rn(k2[o], k3[o]))), k4[o]);
                xprev[o] = __fadd_rn(start[o], __fmul_rn(__fmul_rn(ks, dts[i]), 0.125f));
            }
        }
        // every thread re-reads only the xprev entries it wrote itself (same e -> (i,c) mapping)
        for (int e = tid; e < G_TB * X; e += (int)blockDim.x) {
            const int i = e / X, c = e - i * X;
            const int b = b0 + i;
            if (b < B) p.x_sol.p[(int64_t)j * p.x_sol.st + (int64_t)b * p.x_sol.sb + c] = xprev[i * X4 + c];
        }
        if (DAE) {   // i_j = ae(x_j, z[j], v[j])  (my_solvers.py:121)
            __syncthreads();
            ae_eval(xprev, p.teacher_x ? j : -1, j, -1);
            store_i(j);
        }
    }
}

}  // namespace

int64_t PSN_G_NAME(psn_generic_forward_workspace)(const psnode_problem* p) {
    GenericParams q;
    int packed_floats = 0;
    build_params(p, q, packed_floats, 227 * 1024);
    return (int64_t)packed_floats * 4;
}

int PSN_G_NAME(psn_generic_forward)(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    static GenericParams q;   // large POD; the library is documented as single-caller (no re-entrancy)
    int packed_floats = 0;
    int dev = 0, max_smem = 0;
    PSN_CUDA(cudaGetDevice(&dev));
    PSN_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int smem_bytes = build_params(p, q, packed_floats, max_smem);
    if (smem_bytes < 0) return smem_bytes;
    if (ws == nullptr || ws_bytes < (int64_t)packed_floats * 4) return PSNODE_EWORKSPACE;
    float* packed = static_cast<float*>(ws);
    q.packed = packed;

    PackDesc d;
    d.n = 0;
    for (int net = 0; net < (p->kind == PSNODE_DAE ? 2 : 1); net++) {
        const psnode_mlp& m = net ? p->ae : p->de;
        const PsnPackedNet& pn = net ? q.ae : q.de;
        for (int l = 0; l < m.n_layers; l++) {
            d.W[d.n] = m.W[l]; d.b[d.n] = m.b[l];
            d.in[d.n] = m.in_dim[l]; d.out[d.n] = m.out_dim[l]; d.kpad[d.n] = pn.kpad[l];
            d.w_off[d.n] = pn.w_off[l]; d.b_off[d.n] = pn.b_off[l];
            d.n++;
        }
    }
    psn_pack_kernel<<<dim3(32, d.n), 256, 0, stream>>>(d, packed);
    psn_count_launch("psn_pack_kernel");
    PSN_CUDA(cudaGetLastError());

    const int grid = (p->B + G_TB - 1) / G_TB;
    // layers that do not fit shared memory stream from L2: hide that latency with 4x the warps
    bool streamed = false;
    for (int l = 0; l < q.de.n_layers; l++) streamed |= q.de.smem_off[l] < 0;
    for (int l = 0; l < q.ae.n_layers; l++) streamed |= q.ae.smem_off[l] < 0;
    const int nthreads = streamed ? G_NT_MAX : G_NT;
    if (p->kind == PSNODE_DAE) {
        PSN_CUDA(cudaFuncSetAttribute(psn_generic_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        psn_generic_fwd_kernel<true><<<grid, nthreads, smem_bytes, stream>>>(q);
        psn_count_launch("psn_generic_fwd_kernel<dae>");
    } else {
        PSN_CUDA(cudaFuncSetAttribute(psn_generic_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        psn_generic_fwd_kernel<false><<<grid, nthreads, smem_bytes, stream>>>(q);
        psn_count_launch("psn_generic_fwd_kernel<ode>");
    }
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
