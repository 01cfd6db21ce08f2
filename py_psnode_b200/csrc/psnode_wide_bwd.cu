// psnode_wide_bwd.cu -- tensor-core reverse sweep (discrete adjoint) for the latent nets of the `*_02_direct_encode` scripts.
//
// The reference trains by autograd through the unrolled loop (loss.backward(), neural_00_ODE_02_direct_encode.py:274); this is
// the exact reverse mode of psnode_wide_fwd.cu's step, walking the grid backwards over the forward kernel's tape (the
// post-ELU hidden activation a1 of every stage -- what autograd keeps alive).  Per stage two dependent 128 x 128 GEMMs
// (3xTF32, M = 128, N = 16 trajectories), the transposes of the forward ones:
//     g1      = W2^T . dk_e            delta1 = g1 * ELU'(a1)              W2^T hi + lo resident in TMEM
//     dy_e    = F_x^T . delta1         -> Runge-Kutta adjoint algebra      F_x^T hi in TMEM, lo a shared-memory A operand
// with the same thread <-> element mapping as the forward kernel (one state / neuron element of 8 trajectories per thread).
// The sweep does NOT form weight gradients: it records delta2 = dk_e and delta1 per stage, and sum_e delta1 plus the held
// input z per step, as 128 x 16 operand blocks (psnode_wide.cuh) that psnode_wide_grad.cu turns into dW2, dF_x, dF_z with
// full-rate M = N = 128 MMAs; sum_e delta1 is also written row-major (`dpre`: gradient of the hoisted layer-1 half), from
// which psnode_wide_proj.cu produces the input-series / jump gradients d_z = F_z^T dpre the encoders need.
#include <cstddef>
#include <cstdlib>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int H = PSW_H, TN = PSW_N;
constexpr int LBO = 144;
constexpr int SBO_ACT = (H / 4) * LBO;
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;
constexpr int LBO_W = 128, SBO_W = (H / 4) * LBO_W;
constexpr int TM_A_HI = 0, TM_A_LO = 128, TM_B_HI = 256, TM_ACC = 384;     // A = W2^T, B = F_x^T
constexpr int GROUP_THREADS = PSW_GROUP_THREADS;

struct WideBwdParams {
    int B, T, ngroups;
    psnode_series t, z, gx;
    PsnFuse fx;                                   // fused masked-MSE upstream gradient (replaces gx when its target is set)
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    const float* W1; const float* W2;
    const float* tape;
    float* btape;
    float* stape;
    float* dpre; int64_t dpre_sr;
    float* db2_slab;
    float* d_x0; int64_t d_x0_sb;
    int* err;
};

struct __align__(128) GroupSmem {
    unsigned char act_hi[ACT_TILE];
    unsigned char act_lo[ACT_TILE];
    float dts[2][TN];
    uint64_t bar;
};
struct __align__(128) CtaSmem {
    float fxt_lo[H * H];
    GroupSmem g[PSW_GROUPS_PER_CTA];
    uint32_t tmem_base;
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }

template <int METHOD, int NP>
__global__ void __launch_bounds__(PSW_GROUPS_PER_CTA * GROUP_THREADS, 1) psn_wide_bwd_kernel(const __grid_constant__ WideBwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = cw >> 3, wk = cw & 7, wq = wk & 3, h = wk >> 2;
    const bool issuer = h == 0 && wq < NP;      // NP K-partials, one issuing warp each
    constexpr int KPI = 16 / NP;                 // K-steps (of 8) per issuer
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T;
    const int gid = blockIdx.x * PSW_GROUPS_PER_CTA + g;
    const int b0 = gid * TN;
    const bool live = gid < q.ngroups;
    const int m = 32 * wq + lane;

    if (tid == 0) {
        for (int gg = 0; gg < PSW_GROUPS_PER_CTA; gg++) mbar_init(&sm.g[gg].bar, NP);
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    {   // transposed weights: lane m = input index of the forward layer, columns = its output neurons
        const int cc = cw >> 2;
        for (int ch = 0; ch < 4; ch++) {
            const int k0 = 32 * cc + 8 * ch;
            float ah[8], al[8], bh[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = k0 + i;
                split_tf32(__ldg(q.W2 + k * H + m), ah[i], al[i]);
                float lo;
                split_tf32(__ldg(q.W1 + (int64_t)k * (6 * H) + 2 * H + m) + __ldg(q.W1 + (int64_t)k * (6 * H) + 4 * H + m), bh[i], lo);
                sm.fxt_lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = lo;
            }
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_HI + k0, ah);
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_LO + k0, al);
            tmem_st_32x32b_x8(tmem + lane_base + TM_B_HI + k0, bh);
        }
        tmem_st_wait();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (live) {
        const int off0 = h * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4;
        const int toff = psw_block_off(m, 8 * h);
        const uint32_t idesc = make_idesc_tf32(H, TN);
        static_assert(offsetof(GroupSmem, act_lo) - offsetof(GroupSmem, act_hi) == ACT_TILE, "act_lo must follow act_hi");
        const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = d_act_hi + (uint64_t)(ACT_TILE >> 4);
        const uint64_t d_blo = make_desc(smem_u32(sm.fxt_lo), LBO_W, SBO_W);
        constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
        const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4 * TN);      // (column budget stays 4 x 16 per group)
        const uint32_t my_acc = acc_base + (uint32_t)(wq * TN);
        uint32_t phase = 0;

        auto issue_first = [&]() {          // g1 = W2^T dk
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint32_t wa = term == 0 ? TM_A_LO : TM_A_HI;
                        const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
#pragma unroll
                        for (int kk = 0; kk < KPI; kk++) {
                            const int ks = KPI * wq + kk;
                            mma_tf32_ts(my_acc, tmem + wa + 8 * ks, bd + KSTEP_B * ks, idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
        };
        auto issue_second = [&]() {         // dy = F_x^T delta1
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KPI; kk++) {
                        const int ks = KPI * wq + kk;
                        mma_tf32(my_acc, d_blo + KSTEP_W * ks, d_act_hi + KSTEP_B * ks, idesc, kk > 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int term = 1; term < 3; term++) {
                        const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
#pragma unroll
                        for (int kk = 0; kk < KPI; kk++) {
                            const int ks = KPI * wq + kk;
                            mma_tf32_ts(my_acc, tmem + TM_B_HI + 8 * ks, bd + KSTEP_B * ks, idesc, 1u);
                        }
                    }
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
        };
        auto collect = [&](float (&d)[8]) {
            if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 5); __trap(); }
            phase ^= 1;
            tc_fence_after();
            float t0[8], t1[8], t2[8], t3[8];
            const uint32_t a = acc_base + lane_base + 8 * h;
            tmem_ld_32x32b_x8(a, t0);
            tmem_ld_32x32b_x8(a + TN, t1);
            if constexpr (NP == 4) {
                tmem_ld_32x32b_x8(a + 2 * TN, t2);
                tmem_ld_32x32b_x8(a + 3 * TN, t3);
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if constexpr (NP == 4) d[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
                else d[i] = t0[i] + t1[i];
            }
        };
        auto publish = [&]() {
            fence_async_smem();
            tc_fence_before();
            group_sync(g);
        };
        auto store_tile = [&](const float (&a)[8]) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float hi, lo;
                split_tf32_fast(a[i], hi, lo);
                st_f32(gs.act_hi, off0 + 16 * i, hi);
                st_f32(gs.act_lo, off0 + 16 * i, lo);
            }
        };
        auto tape_block = [&](float* blk, const float (&a)[8]) {
            __stcs(reinterpret_cast<float4*>(blk + toff), make_float4(a[0], a[1], a[2], a[3]));
            __stcs(reinterpret_cast<float4*>(blk + toff + 32), make_float4(a[4], a[5], a[6], a[7]));
        };
        auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };
        auto stage_dt = [&](int j) {
            if (lane < TN) {
                const int bb = min(b0 + lane, B - 1);
                const float* tp = q.t.p + (int64_t)bb * q.t.sb;
                gs.dts[j & 1][lane] = __fsub_rn(__ldg(tp + (int64_t)j * q.t.st), __ldg(tp + (int64_t)(j - 1) * q.t.st));
            }
        };
        const bool fused = q.fx.term.target.p != nullptr;
        const float fscale = fused ? psn_fuse_scale(q.fx) : 0.0f;
        auto load_gx = [&](int j, float (&v)[8]) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int b = b0 + 8 * h + i;
                if (b >= B) v[i] = 0.0f;
                else if (fused) v[i] = psn_fuse_grad(q.fx, fscale, j, b, m);
                else v[i] = __ldg(q.gx.p + (int64_t)j * q.gx.st + (int64_t)b * q.gx.sb + m);
            }
        };

        float lam[8], dyA[8], dyB[8], dysum[8], sum1[8], dk[8];
        float db2 = 0.0f;
        load_gx(T - 1, lam);
        if (T > 1 && wk == 5) stage_dt(T - 1);
        publish();
        const int64_t grec = (int64_t)gid * (T - 1);

        for (int j = T - 1; j >= 1; j--) {
            const float* frec = q.tape + (grec + (j - 1)) * NST * PSW_FWD_REC;
            float* brec = q.btape + (grec + (j - 1)) * NST * PSW_BWD_REC;
            if (j > 1 && wk == 5) stage_dt(j - 1);
#pragma unroll
            for (int i = 0; i < 8; i++) { sum1[i] = 0.0f; dysum[i] = 0.0f; dyA[i] = 0.0f; dyB[i] = 0.0f; }
#pragma unroll 1
            for (int e = NST - 1; e >= 0; e--) {
                float dt[8];
                {
                    const float4 d0 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h]);
                    const float4 d1 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h + 4]);
                    dt[0] = d0.x; dt[1] = d0.y; dt[2] = d0.z; dt[3] = d0.w; dt[4] = d1.x; dt[5] = d1.y; dt[6] = d1.z; dt[7] = d1.w;
                }
                // ---- dL/dk_e from the Runge-Kutta adjoint algebra (my_fixed_grid.py:15-59 reversed) ----
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (METHOD == PSNODE_EULER) dk[i] = lam[i] * dt[i];
                    else if (METHOD == PSNODE_MIDPOINT) dk[i] = e == 1 ? lam[i] * dt[i] : 0.5f * dt[i] * dyA[i];     // dyA = dy_2
                    else {
                        const float l8 = lam[i] * (dt[i] * 0.125f);
                        if (e == 3) dk[i] = l8;
                        else if (e == 2) dk[i] = fmaf(dt[i], dyA[i], 3.0f * l8);                    // dyA = dy_4
                        else if (e == 1) dk[i] = fmaf(dt[i], dyB[i] - dyA[i], 3.0f * l8);           // dyB = dy_3
                        else dk[i] = l8 + dyB[i];                                                   // dyB = dt/3 (dy_2 - dy_3) + dt dy_4
                    }
                    db2 += dk[i];
                }
                store_tile(dk);
                tape_block(brec + (int64_t)e * PSW_BWD_REC, dk);
                publish();
                issue_first();
                float a1[8];
                {
                    const float4 u0 = __ldcs(reinterpret_cast<const float4*>(frec + (int64_t)e * PSW_FWD_REC + toff));
                    const float4 u1 = __ldcs(reinterpret_cast<const float4*>(frec + (int64_t)e * PSW_FWD_REC + toff + 32));
                    a1[0] = u0.x; a1[1] = u0.y; a1[2] = u0.z; a1[3] = u0.w; a1[4] = u1.x; a1[5] = u1.y; a1[6] = u1.z; a1[7] = u1.w;
                }
                float d[8];
                collect(d);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    d[i] *= psn_elu_grad_from_out(a1[i]);          // delta1
                    sum1[i] += d[i];
                }
                store_tile(d);
                tape_block(brec + (int64_t)e * PSW_BWD_REC + PSW_BLOCK, d);
                publish();
                issue_second();
                collect(d);                                        // dy_e
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    dysum[i] += d[i];
                    if (METHOD == PSNODE_MIDPOINT) dyA[i] = d[i];
                    if (METHOD == PSNODE_RK4) {
                        if (e == 3) dyA[i] = d[i];
                        else if (e == 2) dyB[i] = d[i];
                        else if (e == 1) dyB[i] = fmaf(dt[i] * (float)(1.0 / 3.0), d[i] - dyB[i], dt[i] * dyA[i]);
                    }
                }
            }
            // ---- end of step j: gradient of the hoisted layer-1 half, operand blocks of the step, lambda_{j-1} ----
            const int k = event_of_step(j);
            float* srec = q.stape + (grec + (j - 1)) * PSW_STEP_REC;
            tape_block(srec, sum1);
            {
                float zt[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int b = b0 + 8 * h + i, bb = min(b, B - 1);
                    zt[i] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + m)
                                   : __ldg(q.z.p + (int64_t)(j - 1) * q.z.st + (int64_t)bb * q.z.sb + m);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int b = b0 + 8 * h + i;
                    float* row = q.dpre + (int64_t)(j - 1) * q.dpre_sr + (int64_t)b * H + m;      // b < Bpad always
                    if (k >= 0) {
                        row[0] = 0.0f;                                                           // z[j-1] was replaced by the jump value
                        float* jr = q.dpre + ((int64_t)(T - 1) + k) * q.dpre_sr + (int64_t)b * H + m;
                        jr[0] += sum1[i];                                                        // same thread every step: race free
                    } else {
                        row[0] = sum1[i];
                    }
                }
                tape_block(srec + PSW_BLOCK, zt);
            }
            float gprev[8];
            load_gx(j - 1, gprev);
#pragma unroll
            for (int i = 0; i < 8; i++) lam[i] = (lam[i] + dysum[i]) + gprev[i];
        }
        // ---- lambda_0 = dL/dx[0]; per-group bias-gradient partials ----
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int b = b0 + 8 * h + i;
            if (b < B && q.d_x0) q.d_x0[(int64_t)b * q.d_x0_sb + m] = lam[i];
        }
        // db2[m] partial of this group: halves h = 0 / 1 hold disjoint trajectories of the same neuron
        float* slab = reinterpret_cast<float*>(gs.act_hi);
        group_sync(g);
        if (h == 1) slab[m] = db2;
        group_sync(g);
        if (h == 0) q.db2_slab[(int64_t)gid * H + m] = db2 + slab[m];
    }
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

int psn_wide_bwd_sweep(const psnode_problem* p, const psnode_adjoint* a, const float* tape, float* btape, float* stape, float* dpre,
                       float* db2_slab, int* err, cudaStream_t stream) {
    WideBwdParams q;
    q.B = p->B; q.T = p->T; q.ngroups = psw_ngroups(p->B);
    q.t = p->t; q.z = p->z; q.gx = a->gx;
    q.fx = psn_make_fuse(a->fuse_x, p->x_sol);
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.W1 = p->de.W[0]; q.W2 = p->de.W[1];
    q.tape = tape; q.btape = btape; q.stape = stape;
    q.dpre = dpre; q.dpre_sr = psw_bpad(p->B) * PSW_H;
    q.db2_slab = db2_slab;
    q.d_x0 = a->d_x0; q.d_x0_sb = a->d_x0_sb;
    q.err = err;
    const int grid = (q.ngroups + PSW_GROUPS_PER_CTA - 1) / PSW_GROUPS_PER_CTA;
    const int smem = (int)sizeof(CtaSmem) + 128;
    auto launch = [&](auto kern, const char* name) -> int {
        PSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, PSW_GROUPS_PER_CTA * GROUP_THREADS, smem, stream>>>(q);
        psn_count_launch(name);
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    };
    // NP = 4 K-partials; the NP = 2 instantiation lost its A/B and is not built
#define PSW_BWD(METH, NAME) launch(psn_wide_bwd_kernel<METH, 4>, NAME)
    switch (p->method) {
        case PSNODE_EULER: return PSW_BWD(PSNODE_EULER, "psn_wide_bwd_kernel<euler>");
        case PSNODE_MIDPOINT: return PSW_BWD(PSNODE_MIDPOINT, "psn_wide_bwd_kernel<midpoint>");
        default: return PSW_BWD(PSNODE_RK4, "psn_wide_bwd_kernel<rk4>");
    }
#undef PSW_BWD
}
