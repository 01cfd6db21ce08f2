// psnode_wide_fwd.cu -- tensor-core forward integrator for the latent nets of the `*_02_direct_encode` scripts (impl = wide):
// FixedGridODESolver.integrate_ODE (neural_dae/my_solvers.py:52-80) with Euler / Midpoint / RK4-3/8 steps
// (neural_dae/my_fixed_grid.py:15-59) of DE_Func(x_dim = z_dim = hidden = 128): L(768 -> 128) . ELU . L(128 -> 128)
// (neural_00_ODE_02_direct_encode.py:49-57, :70).  BASELINE configs[3] (per-GPU shard B = 4096 x 500 steps).
//
// Per stage two dependent 128 x 128 GEMMs on tcgen05 (3xTF32, M = 128 neurons = TMEM lanes, N = 16 trajectories):
//     a1 = ELU(F_x . y + pre[row])          F_x = (W_b + W_c)[:, 0:128] hi + lo resident in TMEM (TS MMAs)
//     k  = W2 . a1 + b2                     W2 hi resident in TMEM, W2 lo a shared-memory A operand
// `pre` = F_z . z + c is the hoisted, state-independent half of layer 1 (psnode_wide_proj.cu); the 8 KB tile of a group's 16
// trajectories arrives one step ahead by a TMA bulk copy (cp.async.bulk, mbarrier completion).
// TMEM: [0,128) F_x hi | [128,256) F_x lo | [256,384) W2 hi | [384,512) accumulators: 2 groups x 4 K-partials x 16 columns.
// A CTA runs two independent 16-trajectory groups of 8 warps; warps q and q + 4 share TMEM sub-partition q and split the 16
// columns 8 / 8, so a thread owns ONE neuron / state element of 8 trajectories: accumulator fragment, stage algebra in the
// reference's operation order, trajectory row (full 128-byte lines) and tape block (psnode_wide.cuh) all use that mapping.
// Bound: tensor pipe / dependent-layer latency (AI ~ 900 FLOP/B); HBM traffic 1 KB per trajectory-step (+4 KB with the tape).
#include <cstddef>
#include <cstdlib>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int H = PSW_H, TN = PSW_N;
constexpr int LBO = 144;                          // K-chunk stride of the activation tiles (padded: conflict-free stores)
constexpr int SBO_ACT = (H / 4) * LBO;            // 4608
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;      // 9216
constexpr int LBO_W = 128, SBO_W = (H / 4) * LBO_W;   // W2 lo tile in shared memory: 128 rows x K = 128
constexpr int TM_F_HI = 0, TM_F_LO = 128, TM_W2_HI = 256, TM_ACC = 384;
constexpr int GROUP_THREADS = PSW_GROUP_THREADS;
constexpr int PRE_TILE_BYTES = TN * H * 4;        // 8 KB

struct WideFwdParams {
    int B, T, ngroups;
    psnode_series t, x;
    const int32_t* event_idx;
    const float* pre; int64_t pre_sr;             // [rows][Bpad][128]
    const float* W1; const float* W2; const float* b2;
    psnode_series_out x_sol;
    float* tape;
    int* err;
};

struct __align__(128) GroupSmem {
    unsigned char act_hi[ACT_TILE];
    unsigned char act_lo[ACT_TILE];
    float pre[2][TN * H];
    float dts[2][TN];
    uint64_t bar;
    uint64_t pbar[2];
};
struct __align__(128) CtaSmem {
    float w2lo[H * H];
    GroupSmem g[PSW_GROUPS_PER_CTA];
    uint32_t tmem_base;
};

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }

template <int METHOD, bool TAPE, int NP>
__global__ void __launch_bounds__(PSW_GROUPS_PER_CTA * GROUP_THREADS, 1) psn_wide_fwd_kernel(const __grid_constant__ WideFwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform by construction (descriptors stay in uniform registers)
    const int g = cw >> 3, wk = cw & 7, wq = wk & 3, h = wk >> 2;
    const int gt = tid & (GROUP_THREADS - 1);
    const bool issuer = h == 0 && wq < NP;      // NP K-partials, one issuing warp each
    constexpr int KPI = 16 / NP;                 // K-steps (of 8) per issuer
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T;
    const int gid = blockIdx.x * PSW_GROUPS_PER_CTA + g;
    const int b0 = gid * TN;
    const bool live = gid < q.ngroups;
    const int m = 32 * wq + lane;                              // the neuron / state element this thread owns

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int gg = 0; gg < PSW_GROUPS_PER_CTA; gg++) {
            mbar_init(&sm.g[gg].bar, NP);
            mbar_init(&sm.g[gg].pbar[0], 1);
            mbar_init(&sm.g[gg].pbar[1], 1);
        }
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    {   // resident weights: warp (sub-partition wq, column quarter cc) writes lanes 32wq.., columns 32cc..32cc+31 of each matrix
        const int cc = cw >> 2;
        for (int ch = 0; ch < 4; ch++) {
            const int k0 = 32 * cc + 8 * ch;
            float fh[8], fl[8], wh[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = k0 + i;
                // folded layer 1, state part: (W_b + W_c)[m][k], W1 = [W_a | W_b | W_c] with blocks of S = 256 columns
                split_tf32(__ldg(q.W1 + (int64_t)m * (6 * H) + 2 * H + k) + __ldg(q.W1 + (int64_t)m * (6 * H) + 4 * H + k), fh[i], fl[i]);
                float lo;
                split_tf32(__ldg(q.W2 + m * H + k), wh[i], lo);
                sm.w2lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = lo;
            }
            tmem_st_32x32b_x8(tmem + lane_base + TM_F_HI + k0, fh);
            tmem_st_32x32b_x8(tmem + lane_base + TM_F_LO + k0, fl);
            tmem_st_32x32b_x8(tmem + lane_base + TM_W2_HI + k0, wh);
        }
        tmem_st_wait();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (live) {
        const float bias2 = __ldg(q.b2 + m);
        const int off0 = h * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4;       // element i of this thread: off0 + 16 i
        const int toff = psw_block_off(m, 8 * h);                          // tape block: float4 at toff, float4 at toff + 32
        const uint32_t idesc = make_idesc_tf32(H, TN);
        static_assert(offsetof(GroupSmem, act_lo) - offsetof(GroupSmem, act_hi) == ACT_TILE, "act_lo must follow act_hi");
        const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = d_act_hi + (uint64_t)(ACT_TILE >> 4);
        const uint64_t d_w2lo = make_desc(smem_u32(sm.w2lo), LBO_W, SBO_W);
        constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
        const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4 * TN);      // (column budget stays 4 x 16 per group)
        const uint32_t my_acc = acc_base + (uint32_t)(wq * TN);           // the K-partial this (issuing) warp accumulates
        uint32_t phase = 0, pphase = 0;

        // issuing warp wq takes K-steps 4wq..4wq+3 of the three 3xTF32 terms (small terms first) into its own partial
        auto issue_l1 = [&]() {
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint32_t wa = term == 0 ? TM_F_LO : TM_F_HI;
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
        auto issue_l2 = [&]() {
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KPI; kk++) {                       // W2_lo . a_hi : A from shared memory
                        const int ks = KPI * wq + kk;
                        mma_tf32(my_acc, d_w2lo + KSTEP_W * ks, d_act_hi + KSTEP_B * ks, idesc, kk > 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int term = 1; term < 3; term++) {
                        const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
#pragma unroll
                        for (int kk = 0; kk < KPI; kk++) {
                            const int ks = KPI * wq + kk;
                            mma_tf32_ts(my_acc, tmem + TM_W2_HI + 8 * ks, bd + KSTEP_B * ks, idesc, 1u);
                        }
                    }
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
        };
        auto collect = [&](float (&d)[8]) {
            if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
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
            for (int i = 0; i < 8; i += 2) {
                const psn_u64 s01 = psn_add2(psn_pack2(t0[i], t0[i + 1]), psn_pack2(t1[i], t1[i + 1]));
                if constexpr (NP == 4) {
                    const psn_u64 s23 = psn_add2(psn_pack2(t2[i], t2[i + 1]), psn_pack2(t3[i], t3[i + 1]));
                    psn_unpack2(psn_add2(s01, s23), d[i], d[i + 1]);
                } else {
                    psn_unpack2(s01, d[i], d[i + 1]);
                }
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
        // hoisted layer-1 half of step j -> pre[j & 1] (one thread; TMA bulk copy, completes on pbar[j & 1])
        auto fetch_pre = [&](int j) {
            const int k = event_of_step(j);
            const int64_t row = k >= 0 ? (int64_t)(T - 1) + k : (int64_t)(j - 1);
            mbar_expect_tx(&gs.pbar[j & 1], PRE_TILE_BYTES);
            bulk_g2s(gs.pre[j & 1], q.pre + row * q.pre_sr + (int64_t)b0 * H, PRE_TILE_BYTES, &gs.pbar[j & 1]);
        };
        auto stage_dt = [&](int j) {            // lanes 0..15 of one warp: step size of step j for trajectory `lane`
            if (lane < TN) {
                const int bb = min(b0 + lane, B - 1);
                const float* tp = q.t.p + (int64_t)bb * q.t.sb;
                gs.dts[j & 1][lane] = __fsub_rn(__ldg(tp + (int64_t)j * q.t.st), __ldg(tp + (int64_t)(j - 1) * q.t.st));
            }
        };

        // ---- initial state ------------------------------------------------------------------------------------------
        float x0[8], k1[8], k2[8], k3[8], ycur[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int b = b0 + 8 * h + i, bb = min(b, B - 1);
            x0[i] = __ldg(q.x.p + (int64_t)bb * q.x.sb + m);
            if (b < B) q.x_sol.p[(int64_t)b * q.x_sol.sb + m] = x0[i];
            ycur[i] = x0[i];
            k1[i] = k2[i] = k3[i] = 0.0f;
        }
        store_tile(x0);
        if (T > 1) {
            if (wk == 4 && lane == 0) fetch_pre(1);
            if (wk == 5) stage_dt(1);
        }
        publish();
        const float c13 = (float)(1.0 / 3.0);
        float* trec = (TAPE && q.tape) ? q.tape + (int64_t)gid * (T - 1) * NST * PSW_FWD_REC : nullptr;

        for (int j = 1; j < T; j++) {
            float dt[8];
            {
                const float4 d0 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h]);
                const float4 d1 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h + 4]);
                dt[0] = d0.x; dt[1] = d0.y; dt[2] = d0.z; dt[3] = d0.w; dt[4] = d1.x; dt[5] = d1.y; dt[6] = d1.z; dt[7] = d1.w;
            }
            const float* pre = gs.pre[j & 1] + (8 * h) * H + m;
#pragma unroll 1
            for (int e = 0; e < NST; e++) {
                float d[8], a[8];
                issue_l1();
                if (TAPE && trec) tape_block(trec + PSW_BLOCK, ycur);                 // stage input y_e
                if (e == 0) {
                    if (j + 1 < T) {                                               // next step's inputs, one step ahead
                        if (wk == 4 && lane == 0) fetch_pre(j + 1);
                        if (wk == 5) stage_dt(j + 1);
                    }
                    if (!mbar_wait(&gs.pbar[j & 1], (pphase >> (j & 1)) & 1u)) { atomicExch(q.err, 4); __trap(); }
                    pphase ^= 1u << (j & 1);
                }
                // ---- layer 1 epilogue: a1 = ELU(F_x y + pre) ----
                collect(d);
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    const psn_u64 vv = psn_add2(psn_pack2(d[i], d[i + 1]), psn_pack2(pre[i * H], pre[(i + 1) * H]));
                    float v0, v1;
                    psn_unpack2(vv, v0, v1);
                    psn_elu2(v0, v1, a[i], a[i + 1]);
                }
                store_tile(a);
                if (TAPE && trec) tape_block(trec, a);
                publish();
                // ---- layer 2 + stage algebra (reference operation order, my_fixed_grid.py:15-59) ----
                issue_l2();
                collect(d);
                const bool last = e == NST - 1;
                float xn[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float kk = __fadd_rn(d[i], bias2);
                    if (METHOD == PSNODE_EULER) {
                        xn[i] = __fadd_rn(x0[i], __fmul_rn(dt[i], kk));
                    } else if (METHOD == PSNODE_MIDPOINT) {
                        if (e == 0) xn[i] = __fadd_rn(x0[i], __fmul_rn(kk, __fmul_rn(0.5f, dt[i])));
                        else xn[i] = __fadd_rn(x0[i], __fmul_rn(dt[i], kk));
                    } else {
                        if (e == 0) { k1[i] = kk; xn[i] = __fadd_rn(x0[i], __fmul_rn(__fmul_rn(dt[i], kk), c13)); }
                        else if (e == 1) { k2[i] = kk; xn[i] = __fadd_rn(x0[i], __fmul_rn(dt[i], __fsub_rn(kk, __fmul_rn(k1[i], c13)))); }
                        else if (e == 2) { k3[i] = kk; xn[i] = __fadd_rn(x0[i], __fmul_rn(dt[i], __fadd_rn(__fsub_rn(k1[i], k2[i]), kk))); }
                        else {
                            const float ksum = __fadd_rn(__fadd_rn(k1[i], __fmul_rn(3.0f, __fadd_rn(k2[i], k3[i]))), kk);
                            xn[i] = __fadd_rn(x0[i], __fmul_rn(__fmul_rn(ksum, dt[i]), 0.125f));
                        }
                    }
                }
                store_tile(xn);
#pragma unroll
                for (int i = 0; i < 8; i++) ycur[i] = xn[i];
                if (last) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        x0[i] = xn[i];
                        const int b = b0 + 8 * h + i;
                        if (b < B) q.x_sol.p[(int64_t)j * q.x_sol.st + (int64_t)b * q.x_sol.sb + m] = xn[i];
                    }
                }
                if (TAPE && trec) trec += PSW_FWD_REC;
                publish();
            }
        }
    }
    // ---- teardown ---------------------------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

bool psn_wide_proj_view_ok(const float* p, int64_t sr, int64_t sb);

bool psn_wide_supports(const psnode_problem* p) {
    if (p->kind != PSNODE_ODE || p->teacher_x || p->teacher_i) return false;
    if (p->X != PSW_H || p->Z != PSW_H) return false;
    if (p->de.n_layers != 2 || p->de.in_dim[0] != 6 * PSW_H || p->de.out_dim[0] != PSW_H || p->de.out_dim[1] != PSW_H) return false;
    return true;
}

// workspace: [err flags 256 B][c : Bpad x 128][pre : rows x Bpad x 128 (+ one spare tile)]
static int64_t wide_c_floats(int B) { return psw_bpad(B) * PSW_H; }
int64_t psn_wide_forward_workspace(const psnode_problem* p) {
    const int E = p->event_idx ? p->E : 0;
    return 256 + 4 * (wide_c_floats(p->B) + psw_pre_floats(p->B, p->T, E) + PSW_BLOCK);
}

int psn_wide_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < psn_wide_forward_workspace(p)) return PSNODE_EWORKSPACE;
    const int E = p->event_idx ? p->E : 0;
    const int T = p->T, B = p->B;
    int* err = static_cast<int*>(ws);
    float* c = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + 256);
    float* pre = c + wide_c_floats(B);
    const int64_t bpad = psw_bpad(B);
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    const float* W1 = p->de.W[0];
    if (T > 1) {
        int st = psn_wide_const(W1, p->de.b[0], p->a0, p->a0_sb, B, c, stream);
        if (st != PSNODE_OK) return st;
        PswProjJob job;
        job.in = p->z.p; job.in_sr = p->z.st; job.in_sb = p->z.sb;
        job.R = T - 1; job.B = B;
        job.W = W1 + 2 * PSW_H + PSW_H; job.W2 = W1 + 4 * PSW_H + PSW_H; job.ldw = 6 * PSW_H; job.transpose = 0;   // (W_b + W_c)[:, X + k]
        job.add = c; job.add_sb = PSW_H;
        job.out = pre; job.out_sr = bpad * PSW_H; job.out_sb = PSW_H;
        job.zero_rows_from = job.R;
        st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<pre>");
        if (st != PSNODE_OK) return st;
        if (E > 0) {            // event rows: the held input of an event step is z_jump[:, k] (neural_base.py:59-65)
            job.in = p->z_jump; job.in_sr = p->zj_se; job.in_sb = p->zj_sb;
            job.R = E;
            job.out = pre + (int64_t)(T - 1) * bpad * PSW_H;
            st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<pre_jump>");
            if (st != PSNODE_OK) return st;
        }
    }
    WideFwdParams q;
    q.B = B; q.T = T; q.ngroups = psw_ngroups(B);
    q.t = p->t; q.x = p->x;
    q.event_idx = p->event_idx;
    q.pre = pre; q.pre_sr = bpad * PSW_H;
    q.W1 = W1; q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.x_sol = p->x_sol;
    q.tape = (p->tape && p->tape_floats >= psw_tape_floats(B, T, p->method)) ? p->tape : nullptr;
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
#define PSW_FWD(METH, TAPE_, NAME) launch(psn_wide_fwd_kernel<METH, TAPE_, 4>, NAME)
    if (q.tape) {
        switch (p->method) {
            case PSNODE_EULER: return PSW_FWD(PSNODE_EULER, true, "psn_wide_fwd_kernel<euler,tape>");
            case PSNODE_MIDPOINT: return PSW_FWD(PSNODE_MIDPOINT, true, "psn_wide_fwd_kernel<midpoint,tape>");
            default: return PSW_FWD(PSNODE_RK4, true, "psn_wide_fwd_kernel<rk4,tape>");
        }
    }
    switch (p->method) {
        case PSNODE_EULER: return PSW_FWD(PSNODE_EULER, false, "psn_wide_fwd_kernel<euler>");
        case PSNODE_MIDPOINT: return PSW_FWD(PSNODE_MIDPOINT, false, "psn_wide_fwd_kernel<midpoint>");
        default: return PSW_FWD(PSNODE_RK4, false, "psn_wide_fwd_kernel<rk4>");
    }
#undef PSW_FWD
}

// ---- encoded entry on the wide kernels (psnode_forward_encoded for the latent ODE_02 net, ODE_Model.forward,
//      neural_00_ODE_02_direct_encode.py:75-89) ------------------------------------------------------------------------------------------
// z_encoder = L(raw -> 128) . ELU . L(128 -> 128): its second Linear is folded into the hoisted half of layer 1 (A = F_z E2, constant
// F_z e2 added to c), its hidden layer is generated inside the projection kernel's shared-memory tile from the raw (T, B, <= 8) series
// (PswProjJob::gen_*), so the encoded series Zh never exists; the time loop is the TMEM-resident kernel above writing the latent
// trajectory into workspace scratch, and x_decoder runs over that scratch 64 rows at a time (psn_lg_decode).
namespace {
// A[m][k] = sum_h F_z[m][h] E2[h][k],  F_z[m][h] = (W_b + W_c)[m][X + h];  cfold[m] = sum_h F_z[m][h] e2[h]
__global__ void psn_wide_fold_enc_kernel(const float* __restrict__ W1, const float* __restrict__ E2, const float* __restrict__ e2,
                                         float* __restrict__ A, float* __restrict__ cfold) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * H) return;
    const int m = idx / H, k = idx - m * H;
    float acc = 0.0f, accb = 0.0f;
    for (int h = 0; h < H; h++) {
        const float f = __ldg(W1 + (int64_t)m * (6 * H) + 3 * H + h) + __ldg(W1 + (int64_t)m * (6 * H) + 5 * H + h);
        acc = fmaf(f, __ldg(E2 + h * H + k), acc);
        if (k == 0) accb = fmaf(f, __ldg(e2 + h), accb);
    }
    A[idx] = acc;
    if (k == 0) cfold[m] = accb;
}
__global__ void psn_wide_add_rowvec_kernel(float* __restrict__ c, const float* __restrict__ v, int64_t n) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n) c[idx] += __ldg(v + (idx % H));
}
struct EncLayout { int64_t err, c, A, cfold, pre, xs, dec, total; };
EncLayout enc_layout(const psnode_problem* p) {
    EncLayout L;
    const int E = p->event_idx ? p->E : 0;
    auto al = [](int64_t f) { return (f + 63) & ~(int64_t)63; };
    int64_t o = 64;
    L.err = 0;
    L.c = o; o += al(wide_c_floats(p->B));
    L.A = o; o += al((int64_t)H * H);
    L.cfold = o; o += al(H);
    L.pre = o; o += al(psw_pre_floats(p->B, p->T, E) + PSW_BLOCK);
    L.xs = o; o += al((int64_t)p->T * p->B * H);
    L.dec = o; o += al(psn_lg_decode_workspace(p->B, H) / 4);
    L.total = o;
    return L;
}
}  // namespace

bool psn_wide_encoded_supports(const psnode_problem* p, const psnode_codec* cd) {
    if (!p || !cd || p->kind != PSNODE_ODE || p->teacher_x || p->X != H || p->Z != H || p->V || p->I) return false;
    if (p->de.n_layers != 2 || p->de.in_dim[0] != 6 * H || p->de.out_dim[0] != H || p->de.out_dim[1] != H) return false;
    const psnode_mlp& ze = cd->z_enc;
    const psnode_mlp& xd = cd->x_dec;
    if (ze.n_layers != 2 || cd->ZR < 1 || cd->ZR > 8 || ze.in_dim[0] != cd->ZR || ze.out_dim[0] != H || ze.in_dim[1] != H || ze.out_dim[1] != H) return false;
    if (xd.n_layers != 2 || cd->XR < 1 || cd->XR > 128 || xd.in_dim[0] != H || xd.out_dim[0] != H || xd.in_dim[1] != H || xd.out_dim[1] != cd->XR) return false;
    if (!cd->z_raw.p || !cd->x_out.p || !p->t.p || !p->a0 || !p->x.p) return false;
    if (p->event_idx && (p->E < 1 || !cd->zj_raw)) return false;
    static const bool off = std::getenv("PSNODE_ENCODED_WIDE") && std::atoi(std::getenv("PSNODE_ENCODED_WIDE")) == 0;
    return !off;
}

int64_t psn_wide_encoded_workspace(const psnode_problem* p, const psnode_codec* cd) {
    (void)cd;
    return enc_layout(p).total * 4;
}

int psn_wide_forward_encoded(const psnode_problem* p, const psnode_codec* cd, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const EncLayout L = enc_layout(p);
    if (ws == nullptr || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    float* w = static_cast<float*>(ws);
    int* err = reinterpret_cast<int*>(w);
    const int E = p->event_idx ? p->E : 0;
    const int T = p->T, B = p->B;
    const int64_t bpad = psw_bpad(B);
    float* c = w + L.c;
    float* pre = w + L.pre;
    float* xs = w + L.xs;
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    const float* W1 = p->de.W[0];
    if (T > 1) {
        int st = psn_wide_const(W1, p->de.b[0], p->a0, p->a0_sb, B, c, stream);
        if (st != PSNODE_OK) return st;
        psn_wide_fold_enc_kernel<<<(H * H + 255) / 256, 256, 0, stream>>>(W1, cd->z_enc.W[1], cd->z_enc.b[1], w + L.A, w + L.cfold);
        psn_count_launch("psn_wide_fold_enc_kernel");
        psn_wide_add_rowvec_kernel<<<(int)(((int64_t)B * H + 255) / 256), 256, 0, stream>>>(c, w + L.cfold, (int64_t)B * H);
        psn_count_launch("psn_wide_add_rowvec_kernel");
        PswProjJob job;
        job.in = nullptr; job.in_sr = 0; job.in_sb = 0;
        job.R = T - 1; job.B = B;
        job.W = w + L.A; job.W2 = nullptr; job.ldw = H; job.transpose = 0;
        job.add = c; job.add_sb = H;
        job.out = pre; job.out_sr = bpad * H; job.out_sb = H;
        job.zero_rows_from = job.R;
        job.gen_raw = cd->z_raw.p; job.gen_sr = cd->z_raw.st; job.gen_sb = cd->z_raw.sb; job.gen_w = cd->ZR;
        job.gen_W = cd->z_enc.W[0]; job.gen_b = cd->z_enc.b[0];
        st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<pre,enc>");
        if (st != PSNODE_OK) return st;
        if (E > 0) {
            job.R = E;
            job.out = pre + (int64_t)(T - 1) * bpad * H;
            job.gen_raw = cd->zj_raw; job.gen_sr = cd->zjr_se; job.gen_sb = cd->zjr_sb;
            st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<pre_jump,enc>");
            if (st != PSNODE_OK) return st;
        }
    }
    WideFwdParams q;
    q.B = B; q.T = T; q.ngroups = psw_ngroups(B);
    q.t = p->t; q.x = p->x;
    q.event_idx = p->event_idx;
    q.pre = pre; q.pre_sr = bpad * H;
    q.W1 = W1; q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.x_sol.p = xs; q.x_sol.st = (int64_t)B * H; q.x_sol.sb = H;
    q.tape = nullptr;
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
    int st;
    switch (p->method) {
        case PSNODE_EULER: st = launch(psn_wide_fwd_kernel<PSNODE_EULER, false, 4>, "psn_wide_fwd_kernel<euler>"); break;
        case PSNODE_MIDPOINT: st = launch(psn_wide_fwd_kernel<PSNODE_MIDPOINT, false, 4>, "psn_wide_fwd_kernel<midpoint>"); break;
        default: st = launch(psn_wide_fwd_kernel<PSNODE_RK4, false, 4>, "psn_wide_fwd_kernel<rk4>"); break;
    }
    if (st != PSNODE_OK) return st;
    return psn_lg_decode(&cd->x_dec, cd->XR, xs, T, B, H, &cd->x_out, w + L.dec, psn_lg_decode_workspace(B, H), stream);
}
