// psnode_wide4_fwd.cu -- tensor-core forward integrator for the 4-layer ODE_01 DE_Func at hidden width 128, the argparse default
// of the reference's training script (neural_00_ODE_01_no_encode.py:245-247 `--hidden 128`; net :61-68:
// L(3S -> 128) . ELU . L(128 -> 128) . ELU . L(128 -> 128) . ELU . L(128 -> X), S = X + Z, X <= 16, Z <= 8).  Any hidden width up to 128
// runs here zero-padded to 128 neurons (a padded neuron has zero weights and bias, ELU(0) = 0: exact); `impl = auto` sends 64 < H <= 128.
// FixedGridODESolver.integrate_ODE (neural_dae/my_solvers.py:52-80) with Euler / Midpoint / RK4-3/8 steps
// (neural_dae/my_fixed_grid.py:15-59), events as neural_base.py:52-65.  Same machinery as psnode_wide_fwd.cu:
// M = 128 neurons = TMEM lanes, N = 16 trajectories per group, two groups of 8 warps per CTA, 3xTF32 products, 4 K-partials.
//
// Per stage four dependent layers:
//     a1 = ELU(F_x . y + pre)      F_x = (W_b + W_c)[:, 0:X] (128 x K = 16, zero-padded), hi / lo planes in shared memory (SS MMAs);
//                                  pre = F_z . z_held + c,  c = (W_a - W_b) . a0 + b1  (SURVEY 8d folding) is formed per step on the
//                                  CUDA cores: Z <= 8 products per element against 128 x 128 x 2 on the tensor cores
//     a2 = ELU(W2 . a1 + b2)       W2 hi | W2 lo resident in TMEM (TS MMAs)
//     a3 = ELU(W3 . a2 + b3)       W3 hi resident in TMEM, W3 lo a shared-memory A operand
//     k  = W4 . a3 + b4            W4 (X rows, padded to the M = 64 instruction shape) hi / lo planes in shared memory; an M = 64
//                                  accumulator puts rows 0..15 on TMEM lanes 0..15 (psnode_tc.cuh), i.e. on the 16 threads per
//                                  column half that own the state rows
// TMEM: [0,128) W2 hi | [128,256) W2 lo | [256,384) W3 hi | [384,512) accumulators: 2 groups x 4 K-partials x 16 columns.
// The state tile (K = 16 rows) shares the activation tile: rows 0..15 are rewritten with the next stage input after layer 4 and
// layer 1 reads K-steps 0..1 only.  Thread (lane m, half h) owns neuron m of 8 trajectories; threads m < 16 also own state row m.
// Bound: tensor pipe / dependent-layer latency (4 layers x stages x steps); HBM traffic (X + Z + 1) x 4 B read, X x 4 B written per
// trajectory-step.  The `tape` instantiation records a1 | a2 | a3 and the stage input of every stage for the reverse sweep (psnode_wide4_bwd.cu).
#include <cstddef>
#include <cstdlib>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int H = PSW_H, TN = PSW_N;
constexpr int XP = 16;                            // padded state width: K of layer 1, live rows of layer 4
constexpr int ZMAX = 8;                           // held-input width limit
constexpr int M4 = 64;                            // instruction M of layer 4
constexpr int LBO = 144;                          // K-chunk stride of the activation tiles (padded: conflict-free stores)
constexpr int SBO_ACT = (H / 4) * LBO;            // 4608
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;      // 9216
constexpr int LBO_W = 128, SBO_W = (H / 4) * LBO_W;   // K = 128 weight tiles in shared memory (W3 lo: 128 rows, W4: 64 rows)
constexpr int SBO_F = (XP / 4) * LBO_W;           // K = 16 tile of F_x (128 rows)
constexpr int TM_W2_HI = 0, TM_W2_LO = 128, TM_W3_HI = 256, TM_ACC = 384;
constexpr int GROUP_THREADS = PSW_GROUP_THREADS;
constexpr int NP = 4, KPI = 16 / NP;              // K-partials (one issuing warp each), K-steps of 8 per issuer

struct Wide4Params {
    int B, T, ngroups, X, Z, Hh;                  // Hh: the net's hidden width (<= 128; narrower nets are zero-padded to 128 neurons)
    psnode_series t, x, z;
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    const float* a0; int64_t a0_sb;
    const float* W1; const float* b1;
    const float* W2; const float* b2;
    const float* W3; const float* b3;
    const float* W4; const float* b4;
    psnode_series_out x_sol;
    float* tape;                                  // psw4_tape_floats(B, T, method) floats, or NULL
    int* err;
};

struct __align__(128) GroupSmem {
    unsigned char act_hi[ACT_TILE];
    unsigned char act_lo[ACT_TILE];
    float zh[2][TN * ZMAX];
    float dts[2][TN];
    uint64_t bar;                                 // layers 2..4: NP commits per phase
    uint64_t bar1;                                // layer 1: 2 commits per phase
};
struct __align__(128) CtaSmem {
    float w3lo[H * H];
    float w4hi[M4 * H];
    float w4lo[M4 * H];
    float fxhi[H * XP];
    float fxlo[H * XP];
    GroupSmem g[PSW_GROUPS_PER_CTA];
    uint32_t tmem_base;
};
static_assert(sizeof(CtaSmem) + 128 <= 227 * 1024, "one CTA per SM: the tiles must fit the 227 KB opt-in shared memory");

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }

template <int METHOD, bool TAPE>
__global__ void __launch_bounds__(PSW_GROUPS_PER_CTA * GROUP_THREADS, 1) psn_wide4_fwd_kernel(const __grid_constant__ Wide4Params q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform by construction (descriptors stay in uniform registers)
    const int g = cw >> 3, wk = cw & 7, wq = wk & 3, h = wk >> 2;
    const bool issuer = h == 0;                                // warps 0..3 of a group: one K-partial each
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, S = q.X + q.Z, Hh = q.Hh;
    const int gid = blockIdx.x * PSW_GROUPS_PER_CTA + g;
    const int b0 = gid * TN;
    const bool live = gid < q.ngroups;
    const int m = 32 * wq + lane;                              // the neuron (and, below 16, the state row) this thread owns
    const bool mh = m < Hh;                                    // a real neuron (padded ones have zero weights and biases: ELU(0) = 0)
    const float* w1row = q.W1 + (int64_t)(mh ? m : 0) * (3 * S);   // W1 = [W_a | W_b | W_c], blocks of S columns

    // ---- one-time setup ---------------------------------------------------------------------------------------------
    if (tid == 0) {
        for (int gg = 0; gg < PSW_GROUPS_PER_CTA; gg++) {
            mbar_init(&sm.g[gg].bar, NP);
            mbar_init(&sm.g[gg].bar1, 2);
        }
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    {   // resident weights: warp (sub-partition wq, column quarter cc) handles rows 32wq.., columns 32cc..32cc+31 of each matrix
        const int cc = cw >> 2;
        for (int ch = 0; ch < 4; ch++) {
            const int k0 = 32 * cc + 8 * ch;
            float w2h[8], w2l[8], w3h[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = k0 + i;
                const bool in = mh && k < Hh;
                split_tf32(in ? __ldg(q.W2 + m * Hh + k) : 0.0f, w2h[i], w2l[i]);
                float lo;
                split_tf32(in ? __ldg(q.W3 + m * Hh + k) : 0.0f, w3h[i], lo);
                sm.w3lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = lo;
                if (m < M4) {                                   // warp-uniform (wq < 2): output layer, rows >= X are zero
                    float v4h, v4l;
                    split_tf32(m < X && k < Hh ? __ldg(q.W4 + m * Hh + k) : 0.0f, v4h, v4l);
                    sm.w4hi[tile_byte(m, k, LBO_W, SBO_W) >> 2] = v4h;
                    sm.w4lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = v4l;
                }
            }
            tmem_st_32x32b_x8(tmem + lane_base + TM_W2_HI + k0, w2h);
            tmem_st_32x32b_x8(tmem + lane_base + TM_W2_LO + k0, w2l);
            tmem_st_32x32b_x8(tmem + lane_base + TM_W3_HI + k0, w3h);
        }
        tmem_st_wait();
        if (cc == 0) {                                          // folded layer 1, state part: (W_b + W_c)[m][k], k < X, zero-padded to 16
            for (int k = 0; k < XP; k++) {
                float fh, fl;
                split_tf32(mh && k < X ? __ldg(w1row + S + k) + __ldg(w1row + 2 * S + k) : 0.0f, fh, fl);
                sm.fxhi[tile_byte(m, k, LBO_W, SBO_F) >> 2] = fh;
                sm.fxlo[tile_byte(m, k, LBO_W, SBO_F) >> 2] = fl;
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (live) {
        const float bias2 = mh ? __ldg(q.b2 + m) : 0.0f, bias3 = mh ? __ldg(q.b3 + m) : 0.0f;
        const bool state_row = m < XP;                          // lanes 0..15 of the sub-partition-0 warps (both column halves)
        const bool live_x = m < X;
        const float bias4 = live_x ? __ldg(q.b4 + m) : 0.0f;
        // held-input half of the folded layer 1 and the per-trajectory constant (fp32 FMA, once per call)
        float fz[ZMAX], cst[8];
#pragma unroll
        for (int k = 0; k < ZMAX; k++) fz[k] = mh && k < Z ? __ldg(w1row + S + X + k) + __ldg(w1row + 2 * S + X + k) : 0.0f;
        {
            const float b1m = mh ? __ldg(q.b1 + m) : 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) cst[i] = b1m;
            for (int k = 0; k < S; k++) {
                const float wd = mh ? __ldg(w1row + k) - __ldg(w1row + S + k) : 0.0f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int bb = min(b0 + 8 * h + i, B - 1);
                    cst[i] = fmaf(wd, __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), cst[i]);
                }
            }
        }
        const int off0 = h * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4;       // element i of this thread: off0 + 16 i
        const uint32_t idesc = make_idesc_tf32(H, TN), idesc4 = make_idesc_tf32(M4, TN);
        static_assert(offsetof(GroupSmem, act_lo) - offsetof(GroupSmem, act_hi) == ACT_TILE, "act_lo must follow act_hi");
        const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = d_act_hi + (uint64_t)(ACT_TILE >> 4);
        const uint64_t d_w3lo = make_desc(smem_u32(sm.w3lo), LBO_W, SBO_W);
        const uint64_t d_w4hi = make_desc(smem_u32(sm.w4hi), LBO_W, SBO_W), d_w4lo = make_desc(smem_u32(sm.w4lo), LBO_W, SBO_W);
        const uint64_t d_fxhi = make_desc(smem_u32(sm.fxhi), LBO_W, SBO_F), d_fxlo = make_desc(smem_u32(sm.fxlo), LBO_W, SBO_F);
        constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
        const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * NP * TN);
        const uint32_t my_acc = acc_base + (uint32_t)(wq * TN);           // the K-partial this (issuing) warp accumulates
        uint32_t phase = 0, phase1 = 0;

        // layer 1: K = 16 = two K-steps, issuing warps 0 and 1 take one each (small terms first) into partials 0 and 1
        auto issue_l1 = [&]() {
            if (issuer && wq < 2) {
                if (elect_one()) {
                    tc_fence_after();
                    const uint64_t ka = KSTEP_W * (uint64_t)wq, kb = KSTEP_B * (uint64_t)wq;
                    mma_tf32(my_acc, d_fxlo + ka, d_act_hi + kb, idesc, 0u);
                    mma_tf32(my_acc, d_fxhi + ka, d_act_lo + kb, idesc, 1u);
                    mma_tf32(my_acc, d_fxhi + ka, d_act_hi + kb, idesc, 1u);
                    mma_commit(&gs.bar1);
                }
                __syncwarp();
            }
        };
        // layer 2: both planes of W2 in TMEM; issuing warp wq takes K-steps 4wq..4wq+3 of the three terms into its own partial
        auto issue_l2 = [&]() {
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint32_t wa = term == 0 ? TM_W2_LO : TM_W2_HI;
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
        // layer 3: W3 lo from shared memory, W3 hi from TMEM
        auto issue_l3 = [&]() {
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KPI; kk++) {
                        const int ks = KPI * wq + kk;
                        mma_tf32(my_acc, d_w3lo + KSTEP_W * ks, d_act_hi + KSTEP_B * ks, idesc, kk > 0 ? 1u : 0u);
                    }
#pragma unroll
                    for (int term = 1; term < 3; term++) {
                        const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
#pragma unroll
                        for (int kk = 0; kk < KPI; kk++) {
                            const int ks = KPI * wq + kk;
                            mma_tf32_ts(my_acc, tmem + TM_W3_HI + 8 * ks, bd + KSTEP_B * ks, idesc, 1u);
                        }
                    }
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
        };
        // layer 4: M = 64 instruction shape, both planes of W4 from shared memory
        auto issue_l4 = [&]() {
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint64_t ad = term == 0 ? d_w4lo : d_w4hi;
                        const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
#pragma unroll
                        for (int kk = 0; kk < KPI; kk++) {
                            const int ks = KPI * wq + kk;
                            mma_tf32(my_acc, ad + KSTEP_W * ks, bd + KSTEP_B * ks, idesc4, accumulate);
                            accumulate = 1;
                        }
                    }
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
        };
        auto collect4 = [&](float (&d)[8]) {
            if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
            phase ^= 1;
            tc_fence_after();
            float t0[8], t1[8], t2[8], t3[8];
            const uint32_t a = acc_base + lane_base + 8 * h;
            tmem_ld_32x32b_x8(a, t0);
            tmem_ld_32x32b_x8(a + TN, t1);
            tmem_ld_32x32b_x8(a + 2 * TN, t2);
            tmem_ld_32x32b_x8(a + 3 * TN, t3);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const psn_u64 s01 = psn_add2(psn_pack2(t0[i], t0[i + 1]), psn_pack2(t1[i], t1[i + 1]));
                const psn_u64 s23 = psn_add2(psn_pack2(t2[i], t2[i + 1]), psn_pack2(t3[i], t3[i + 1]));
                psn_unpack2(psn_add2(s01, s23), d[i], d[i + 1]);
            }
        };
        auto collect2 = [&](float (&d)[8]) {
            if (!mbar_wait(&gs.bar1, phase1)) { atomicExch(q.err, 2); __trap(); }
            phase1 ^= 1;
            tc_fence_after();
            float t0[8], t1[8];
            const uint32_t a = acc_base + lane_base + 8 * h;
            tmem_ld_32x32b_x8(a, t0);
            tmem_ld_32x32b_x8(a + TN, t1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                psn_unpack2(psn_add2(psn_pack2(t0[i], t0[i + 1]), psn_pack2(t1[i], t1[i + 1])), d[i], d[i + 1]);
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
        // hidden-layer epilogue: a = ELU(acc + bias) -> the next layer's operand tile
        const int toff = psw_block_off(m, 8 * h);                          // tape block: float4 at toff, float4 at toff + 32
        auto tape_block = [&](float* blk, const float (&a)[8]) {
            __stcs(reinterpret_cast<float4*>(blk + toff), make_float4(a[0], a[1], a[2], a[3]));
            __stcs(reinterpret_cast<float4*>(blk + toff + 32), make_float4(a[4], a[5], a[6], a[7]));
        };
        auto hidden_epilogue = [&](const float (&d)[8], float bias, float* trec_blk) {
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const psn_u64 vv = psn_add2(psn_pack2(d[i], d[i + 1]), psn_pack2(bias, bias));
                float v0, v1;
                psn_unpack2(vv, v0, v1);
                psn_elu2(v0, v1, a[i], a[i + 1]);
            }
            store_tile(a);
            if (TAPE && trec_blk) tape_block(trec_blk, a);
        };
        auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };
        // one warp, one step ahead: the held input of step j (z[j-1], or z_jump[:, k] when event k fires at t[j-1]:
        // neural_base.py:59-65) for trajectory lane % 16, feature quarter lane / 16; and the step sizes
        auto stage_held = [&](int j) {
            const int n = lane & 15, kh = (lane >> 4) * 4;
            const int bb = min(b0 + n, B - 1);
            const int ek = event_of_step(j);
            const float* src = ek >= 0 ? q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)ek * q.zj_se
                                       : q.z.p + (int64_t)(j - 1) * q.z.st + (int64_t)bb * q.z.sb;
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
                if (kh + kk < Z) gs.zh[j & 1][n * ZMAX + kh + kk] = __ldg(src + kh + kk);
        };
        auto stage_dt = [&](int j) {
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
            x0[i] = 0.0f;
            k1[i] = k2[i] = k3[i] = 0.0f;
        }
        if (state_row) {
            if (live_x) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int b = b0 + 8 * h + i, bb = min(b, B - 1);
                    x0[i] = __ldg(q.x.p + (int64_t)bb * q.x.sb + m);
                    if (b < B) q.x_sol.p[(int64_t)b * q.x_sol.sb + m] = x0[i];
                }
            }
            store_tile(x0);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) ycur[i] = x0[i];
        // tape record per (group, step, stage): a1 | a2 | a3 as 128 x 16 blocks, then the stage input y as [trajectory][16 state rows]
        float* trec = (TAPE && q.tape) ? q.tape + (int64_t)gid * (T - 1) * NST * PSW4_FWD_REC : nullptr;
        if (T > 1) {
            if (wk == 4) stage_held(1);
            if (wk == 5) stage_dt(1);
        }
        publish();
        const float c13 = (float)(1.0 / 3.0);

        for (int j = 1; j < T; j++) {
            // hoisted half of layer 1 for this step: pre = c + F_z . z_held (constant over the stages)
            float pre[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float* zr = &gs.zh[j & 1][(8 * h + i) * ZMAX];
                float acc = cst[i];
#pragma unroll
                for (int k = 0; k < ZMAX; k++)
                    if (k < Z) acc = fmaf(fz[k], zr[k], acc);
                pre[i] = acc;
            }
#pragma unroll 1
            for (int e = 0; e < NST; e++) {
                float d[8];
                issue_l1();
                if (TAPE && trec && state_row) {
#pragma unroll
                    for (int i = 0; i < 8; i++) trec[3 * PSW_BLOCK + (8 * h + i) * XP + m] = ycur[i];
                }
                if (e == 0 && j + 1 < T) {                                         // next step's inputs, one step ahead
                    if (wk == 4) stage_held(j + 1);
                    if (wk == 5) stage_dt(j + 1);
                }
                // ---- layer 1 epilogue: a1 = ELU(F_x y + pre) ----
                collect2(d);
                {
                    float a[8];
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const psn_u64 vv = psn_add2(psn_pack2(d[i], d[i + 1]), psn_pack2(pre[i], pre[i + 1]));
                        float v0, v1;
                        psn_unpack2(vv, v0, v1);
                        psn_elu2(v0, v1, a[i], a[i + 1]);
                    }
                    store_tile(a);
                    if (TAPE && trec) tape_block(trec, a);
                }
                publish();
                // ---- layers 2 and 3 ----
                issue_l2();
                collect4(d);
                hidden_epilogue(d, bias2, trec ? trec + PSW_BLOCK : nullptr);
                publish();
                issue_l3();
                collect4(d);
                hidden_epilogue(d, bias3, trec ? trec + 2 * PSW_BLOCK : nullptr);
                publish();
                // ---- layer 4 + stage algebra (reference operation order, my_fixed_grid.py:15-59) on the 16 state rows ----
                issue_l4();
                collect4(d);
                if (state_row) {
                    const bool last = e == NST - 1;
                    float dt[8], xn[8];
                    {
                        const float4 d0 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h]);
                        const float4 d1 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h + 4]);
                        dt[0] = d0.x; dt[1] = d0.y; dt[2] = d0.z; dt[3] = d0.w; dt[4] = d1.x; dt[5] = d1.y; dt[6] = d1.z; dt[7] = d1.w;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float kk = __fadd_rn(d[i], bias4);
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
                        if (!live_x) xn[i] = 0.0f;                                   // padded state rows stay exactly zero
                    }
                    store_tile(xn);
#pragma unroll
                    for (int i = 0; i < 8; i++) ycur[i] = xn[i];
                    if (last) {
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            x0[i] = xn[i];
                            const int b = b0 + 8 * h + i;
                            if (live_x && b < B) q.x_sol.p[(int64_t)j * q.x_sol.st + (int64_t)b * q.x_sol.sb + m] = xn[i];
                        }
                    }
                }
                if (TAPE && trec) trec += PSW4_FWD_REC;
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

// PSNODE_WIDE4=0 keeps `impl = auto` off this kernel (the generic CUDA-core kernel takes the shape); `impl = wide` always reaches it
bool psn_wide4_auto(const psnode_problem* p) {
    static const bool on = [] {
        const char* e = getenv("PSNODE_WIDE4");
        return !(e && e[0] == '0');
    }();
    return on && p->de.out_dim[0] > 64;      // narrower nets cost the same 128-neuron time here: they stay on the CUDA-core kernels
}

bool psn_wide4_supports(const psnode_problem* p) {
    if (p->kind != PSNODE_ODE || p->teacher_x || p->teacher_i) return false;
    if (p->X < 1 || p->X > XP || p->Z < 0 || p->Z > ZMAX || p->V != 0 || p->I != 0) return false;
    const psnode_mlp& n = p->de;
    if (n.n_layers != 4 || n.in_dim[0] != 3 * (p->X + p->Z) || n.out_dim[3] != p->X) return false;
    const int hh = n.out_dim[0];
    if (hh < 1 || hh > H) return false;
    for (int l = 0; l < 3; l++)
        if (n.out_dim[l] != hh || n.in_dim[l + 1] != hh) return false;
    if (p->event_idx && p->Z > 0 && !p->z_jump) return false;
    return true;
}

int64_t psn_wide4_forward_workspace(const psnode_problem*) { return 256; }

int psn_wide4_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < psn_wide4_forward_workspace(p)) return PSNODE_EWORKSPACE;
    int* err = static_cast<int*>(ws);
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    Wide4Params q;
    q.B = p->B; q.T = p->T; q.ngroups = psw_ngroups(p->B); q.X = p->X; q.Z = p->Z; q.Hh = p->de.out_dim[0];
    q.t = p->t; q.x = p->x; q.z = p->z;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.W1 = p->de.W[0]; q.b1 = p->de.b[0];
    q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.W3 = p->de.W[2]; q.b3 = p->de.b[2];
    q.W4 = p->de.W[3]; q.b4 = p->de.b[3];
    q.x_sol = p->x_sol;
    q.tape = (p->tape && p->tape_floats >= psw4_tape_floats(p->B, p->T, p->method)) ? p->tape : nullptr;
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
    if (q.tape) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_wide4_fwd_kernel<PSNODE_EULER, true>, "psn_wide4_fwd_kernel<euler,tape>");
            case PSNODE_MIDPOINT: return launch(psn_wide4_fwd_kernel<PSNODE_MIDPOINT, true>, "psn_wide4_fwd_kernel<midpoint,tape>");
            default: return launch(psn_wide4_fwd_kernel<PSNODE_RK4, true>, "psn_wide4_fwd_kernel<rk4,tape>");
        }
    }
    switch (p->method) {
        case PSNODE_EULER: return launch(psn_wide4_fwd_kernel<PSNODE_EULER, false>, "psn_wide4_fwd_kernel<euler>");
        case PSNODE_MIDPOINT: return launch(psn_wide4_fwd_kernel<PSNODE_MIDPOINT, false>, "psn_wide4_fwd_kernel<midpoint>");
        default: return launch(psn_wide4_fwd_kernel<PSNODE_RK4, false>, "psn_wide4_fwd_kernel<rk4>");
    }
}
