// psnode_wide4_bwd.cu -- tensor-core reverse sweep (discrete adjoint) for the 4-layer ODE_01 DE_Func at hidden widths <= 128
// (neural_00_ODE_01_no_encode.py:61-68; the script trains by loss.backward() through the unrolled loop, :350-359).  Exact reverse mode of
// psnode_wide4_fwd.cu's step over that kernel's tape (the post-ELU activations a1, a2, a3 and the stage input y of every stage -- what
// autograd keeps alive).  Per stage the four forward layers transposed, same machinery (M = 128 lanes, N = 16 trajectories, 3xTF32,
// 4 K-partials, two groups of 8 warps per CTA):
//     g3 = W4^T . dk_e        (K = 16, SS)           delta3 = g3 * ELU'(a3)        W4^T hi / lo: 128 x 16 tiles in shared memory
//     g2 = W3^T . delta3      (TS)                   delta2 = g2 * ELU'(a2)        W3^T hi | lo resident in TMEM
//     g1 = W2^T . delta2      (lo SS, hi TS)         delta1 = g1 * ELU'(a1)        W2^T hi in TMEM, lo in shared memory
//     dy_e = F_x^T . delta1   (M = 64, SS)           -> Runge-Kutta adjoint        F_x^T hi / lo: 64 x 128 tiles (rows >= X zero); an M = 64
//                                                                                  accumulator puts rows 0..15 on the state threads' lanes
// Weight gradients: the two 128 x 128 ones, dW2 = sum delta2 . a1^T and dW3 = sum delta3 . a2^T, come from psn_wide_grad_kernel (block GEMMs over
// the recorded delta2 / delta3 blocks and the forward tape); the narrow ones are formed here on the CUDA cores from registers and two small
// fp32 tiles -- dW4 = sum dk . a3^T (16 x 128), dF_x = sum delta1 . y^T (128 x 16), dF_z = sum (sum_e delta1) . z_held^T (128 x Z), the
// per-trajectory constant's gradient dc = sum delta1 (-> dW_a, dW_b, db1, d_a0) and the bias gradients -- and leave the kernel as one slab per
// (group, column half), summed in a fixed order by psn_wide4_assemble_kernel (deterministic, no floating-point atomics).
// Input-series / jump gradients are not produced (the `*_01` scripts feed raw data there): such calls take the generic sweep.
#include <cstddef>
#include <cstdlib>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int H = PSW_H, TN = PSW_N;
constexpr int XP = 16, ZMAX = 8, M4 = 64;
constexpr int LBO = 144;
constexpr int SBO_ACT = (H / 4) * LBO;
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;
constexpr int LBO_W = 128, SBO_W = (H / 4) * LBO_W;
constexpr int SBO_F = (XP / 4) * LBO_W;
constexpr int TM_A_HI = 0, TM_A_LO = 128, TM_B_HI = 256, TM_ACC = 384;     // A = W3^T, B = W2^T
constexpr int GROUP_THREADS = PSW_GROUP_THREADS;
constexpr int NP = 4, KPI = 16 / NP;
constexpr int SMAX = XP + ZMAX;                   // width of all_initial: at most 24
// slab of one (group, column half): fields of 128 floats (one per neuron / lane m)
constexpr int F_DFX = 0, F_DFZ = 16, F_DCA = 24, F_DB1 = 48, F_DB2 = 49, F_DB3 = 50, F_DW4 = 51, F_DB4 = 67, SLAB_FIELDS = 68;

struct Wide4BwdParams {
    int B, T, ngroups, X, Z, Hh;
    psnode_series t, z, gx;
    PsnFuse fx;                                   // fused masked-MSE upstream gradient (replaces gx when its target is set)
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    const float* a0; int64_t a0_sb;
    const float* W1; const float* W2; const float* W3; const float* W4;
    const float* tape;
    float* btape;
    float* slabs;                                 // [2 * ngroups][SLAB_FIELDS][128]
    float* d_x0; int64_t d_x0_sb;
    float* d_a0; int64_t d_a0_sb;
    int pf;                                       // stage y one stage ahead with cp.async + L2-prefetch the next record (PSNODE_WIDE4_PF=0: off)
    int* err;
};

struct __align__(128) GroupSmem {
    unsigned char act_hi[ACT_TILE];
    unsigned char act_lo[ACT_TILE];
    float dkt[TN * XP];                           // dk_e of the stage, fp32, [trajectory][state row]
    float yt[2][TN * XP];                         // stage input y_e, fp32, same layout (from the forward tape); double-buffered by stage
    float rk[4][TN * XP];                         // state threads' private adjoint state: lambda | dyA | dyB | sum_e dy_e, same layout
    float zh[2][TN * ZMAX];
    float dts[2][TN];
    uint64_t bar;
    uint64_t bar1;
};
struct __align__(128) CtaSmem {
    float w2t_lo[H * H];
    float fxt_hi[M4 * H];
    float fxt_lo[M4 * H];
    float w4t_hi[H * XP];
    float w4t_lo[H * XP];
    GroupSmem g[PSW_GROUPS_PER_CTA];
    uint32_t tmem_base;
};
static_assert(sizeof(CtaSmem) + 128 <= 227 * 1024, "one CTA per SM: the tiles must fit the 227 KB opt-in shared memory");
static_assert(ACT_TILE >= TN * H * 4, "the activation tile doubles as the dc scratch of the d_a0 epilogue");

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int METHOD>
__global__ void __launch_bounds__(PSW_GROUPS_PER_CTA * GROUP_THREADS, 1) psn_wide4_bwd_kernel(const __grid_constant__ Wide4BwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = cw >> 3, wk = cw & 7, wq = wk & 3, h = wk >> 2;
    const int gt = tid & (GROUP_THREADS - 1);
    const bool issuer = h == 0;
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, S = q.X + q.Z, Hh = q.Hh;
    const int gid = blockIdx.x * PSW_GROUPS_PER_CTA + g;
    const int b0 = gid * TN;
    const bool live = gid < q.ngroups;
    const int m = 32 * wq + lane;
    const bool mh = m < Hh;

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
    {   // transposed weights: row m = INPUT index of the forward layer, column k = its OUTPUT neuron (zero beyond the net's width)
        const int cc = cw >> 2;
        for (int ch = 0; ch < 4; ch++) {
            const int k0 = 32 * cc + 8 * ch;
            float ah[8], al[8], bh[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = k0 + i;
                const bool in = mh && k < Hh;
                split_tf32(in ? __ldg(q.W3 + k * Hh + m) : 0.0f, ah[i], al[i]);
                float lo;
                split_tf32(in ? __ldg(q.W2 + k * Hh + m) : 0.0f, bh[i], lo);
                sm.w2t_lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = lo;
                if (m < M4) {                                   // warp-uniform: F_x^T[m][k] = (W_b + W_c)[k][m], state rows m < X
                    float fh, fl;
                    const float* wr = q.W1 + (int64_t)(k < Hh ? k : 0) * (3 * S);
                    split_tf32(m < X && k < Hh ? __ldg(wr + S + m) + __ldg(wr + 2 * S + m) : 0.0f, fh, fl);
                    sm.fxt_hi[tile_byte(m, k, LBO_W, SBO_W) >> 2] = fh;
                    sm.fxt_lo[tile_byte(m, k, LBO_W, SBO_W) >> 2] = fl;
                }
            }
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_HI + k0, ah);
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_LO + k0, al);
            tmem_st_32x32b_x8(tmem + lane_base + TM_B_HI + k0, bh);
        }
        tmem_st_wait();
        if (cc == 0) {                                          // W4^T[m][k] = W4[k][m], k < X, zero-padded to K = 16
            for (int k = 0; k < XP; k++) {
                float wh, wl;
                split_tf32(mh && k < X ? __ldg(q.W4 + k * Hh + m) : 0.0f, wh, wl);
                sm.w4t_hi[tile_byte(m, k, LBO_W, SBO_F) >> 2] = wh;
                sm.w4t_lo[tile_byte(m, k, LBO_W, SBO_F) >> 2] = wl;
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (live) {
        const bool state_row = m < XP;                          // lanes 0..15 of the sub-partition-0 warps (both column halves)
        const bool live_x = m < X;
        const int off0 = h * SBO_ACT + (m >> 2) * LBO + (m & 3) * 4;
        const int toff = psw_block_off(m, 8 * h);
        const uint32_t idesc = make_idesc_tf32(H, TN), idesc4 = make_idesc_tf32(M4, TN);
        static_assert(offsetof(GroupSmem, act_lo) - offsetof(GroupSmem, act_hi) == ACT_TILE, "act_lo must follow act_hi");
        const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = d_act_hi + (uint64_t)(ACT_TILE >> 4);
        const uint64_t d_w2tlo = make_desc(smem_u32(sm.w2t_lo), LBO_W, SBO_W);
        const uint64_t d_fxthi = make_desc(smem_u32(sm.fxt_hi), LBO_W, SBO_W), d_fxtlo = make_desc(smem_u32(sm.fxt_lo), LBO_W, SBO_W);
        const uint64_t d_w4thi = make_desc(smem_u32(sm.w4t_hi), LBO_W, SBO_F), d_w4tlo = make_desc(smem_u32(sm.w4t_lo), LBO_W, SBO_F);
        constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
        const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * NP * TN);
        const uint32_t my_acc = acc_base + (uint32_t)(wq * TN);
        uint32_t phase = 0, phase1 = 0;

        auto issue_w4t = [&]() {            // g3 = W4^T dk: K = 16, issuing warps 0 and 1 take one K-step each
            if (issuer && wq < 2) {
                if (elect_one()) {
                    tc_fence_after();
                    const uint64_t ka = KSTEP_W * (uint64_t)wq, kb = KSTEP_B * (uint64_t)wq;
                    mma_tf32(my_acc, d_w4tlo + ka, d_act_hi + kb, idesc, 0u);
                    mma_tf32(my_acc, d_w4thi + ka, d_act_lo + kb, idesc, 1u);
                    mma_tf32(my_acc, d_w4thi + ka, d_act_hi + kb, idesc, 1u);
                    mma_commit(&gs.bar1);
                }
                __syncwarp();
            }
        };
        auto issue_w3t = [&]() {            // g2 = W3^T delta3: both planes in TMEM
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
        auto issue_w2t = [&]() {            // g1 = W2^T delta2: lo plane from shared memory, hi plane from TMEM
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < KPI; kk++) {
                        const int ks = KPI * wq + kk;
                        mma_tf32(my_acc, d_w2tlo + KSTEP_W * ks, d_act_hi + KSTEP_B * ks, idesc, kk > 0 ? 1u : 0u);
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
        auto issue_fxt = [&]() {            // dy = F_x^T delta1: M = 64 instruction shape, both planes from shared memory
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint64_t ad = term == 0 ? d_fxtlo : d_fxthi;
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
            if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 5); __trap(); }
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
            for (int i = 0; i < 8; i++) d[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
        };
        auto collect2 = [&](float (&d)[8]) {
            if (!mbar_wait(&gs.bar1, phase1)) { atomicExch(q.err, 6); __trap(); }
            phase1 ^= 1;
            tc_fence_after();
            float t0[8], t1[8];
            const uint32_t a = acc_base + lane_base + 8 * h;
            tmem_ld_32x32b_x8(a, t0);
            tmem_ld_32x32b_x8(a + TN, t1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i++) d[i] = t0[i] + t1[i];
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
        auto load_block = [&](const float* blk, float (&a)[8]) {
            const float4 u0 = __ldcs(reinterpret_cast<const float4*>(blk + toff));
            const float4 u1 = __ldcs(reinterpret_cast<const float4*>(blk + toff + 32));
            a[0] = u0.x; a[1] = u0.y; a[2] = u0.z; a[3] = u0.w; a[4] = u1.x; a[5] = u1.y; a[6] = u1.z; a[7] = u1.w;
        };
        auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };
        auto stage_held = [&](int j) {      // one warp: the held input of step j (as the forward kernel staged it)
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
        const bool fused = q.fx.term.target.p != nullptr;
        const float fscale = fused ? psn_fuse_scale(q.fx) : 0.0f;
        auto load_gx = [&](int j, float (&v)[8]) {              // state threads only; padded rows and trajectories carry zero
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int b = b0 + 8 * h + i;
                if (b >= B || !live_x) v[i] = 0.0f;
                else if (fused) v[i] = psn_fuse_grad(q.fx, fscale, j, b, m);
                else v[i] = __ldg(q.gx.p + (int64_t)j * q.gx.st + (int64_t)b * q.gx.sb + m);
            }
        };

        // state threads: adjoint of the state and the Runge-Kutta bookkeeping; every thread: its slice of the narrow gradients
        // (the adjoint state lives in shared memory, one private slot per state thread and trajectory: 32 registers fewer, no spills)
        float* const lam = &gs.rk[0][8 * h * XP + m];           // element i at [i * XP]
        float* const dyA = &gs.rk[1][8 * h * XP + m];
        float* const dyB = &gs.rk[2][8 * h * XP + m];
        float* const dysum = &gs.rk[3][8 * h * XP + m];
        float dw4[XP], dfx[XP], dfz[ZMAX], dcs[8], sum1[8];
        float db2 = 0.0f, db3 = 0.0f, db4 = 0.0f;
#pragma unroll
        for (int k = 0; k < XP; k++) { dw4[k] = 0.0f; dfx[k] = 0.0f; }
#pragma unroll
        for (int k = 0; k < ZMAX; k++) dfz[k] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) dcs[i] = 0.0f;
        if (state_row) {
            float g0[8];
            load_gx(T - 1, g0);
#pragma unroll
            for (int i = 0; i < 8; i++) { lam[i * XP] = g0[i]; dyA[i * XP] = 0.0f; dyB[i * XP] = 0.0f; dysum[i * XP] = 0.0f; }
        }
        if (T > 1) {
            if (wk == 5) stage_dt(T - 1);
            if (wk == 6) stage_held(T - 1);
        }
        const int64_t grec = (int64_t)gid * (T - 1);
        const bool pf = q.pf != 0;
        // one warp: the 1 KB stage-input block of a record -> yt[buf], asynchronously (no registers, nothing on the state threads' path)
        auto stage_y = [&](const float* rec, int buf) {
            const float* src = rec + 3 * PSW_BLOCK;
            cp_async16(&gs.yt[buf][4 * lane], src + 4 * lane);
            cp_async16(&gs.yt[buf][128 + 4 * lane], src + 128 + 4 * lane);
        };
        if (pf && T > 1 && wk == 7) {
            stage_y(q.tape + ((grec + (T - 2)) * NST + (NST - 1)) * PSW4_FWD_REC, 0);
            cp_async_wait_all();
        }
        publish();
        int sc = 0;                                             // stage counter: yt[sc & 1] is this stage's buffer

        for (int j = T - 1; j >= 1; j--) {
            const float* frec = q.tape + (grec + (j - 1)) * NST * PSW4_FWD_REC;
            float* brec = q.btape + (grec + (j - 1)) * NST * PSW4_BWD_REC;
            if (j > 1 && wk == 5) stage_dt(j - 1);
#pragma unroll
            for (int i = 0; i < 8; i++) sum1[i] = 0.0f;
            if (state_row) {
#pragma unroll
                for (int i = 0; i < 8; i++) { dysum[i * XP] = 0.0f; dyA[i * XP] = 0.0f; dyB[i * XP] = 0.0f; }
            }
#pragma unroll 1
            for (int e = NST - 1; e >= 0; e--) {
                const float* fr = frec + (int64_t)e * PSW4_FWD_REC;
                float* br = brec + (int64_t)e * PSW4_BWD_REC;
                float dt[8], d[8], a[8];
                if (state_row) {
                    const float4 d0 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h]);
                    const float4 d1 = *reinterpret_cast<const float4*>(&gs.dts[j & 1][8 * h + 4]);
                    dt[0] = d0.x; dt[1] = d0.y; dt[2] = d0.z; dt[3] = d0.w; dt[4] = d1.x; dt[5] = d1.y; dt[6] = d1.z; dt[7] = d1.w;
                    // ---- dL/dk_e from the Runge-Kutta adjoint algebra (my_fixed_grid.py:15-59 reversed, as psnode_wide_bwd.cu) ----
                    float dk[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float lm = lam[i * XP];
                        if (METHOD == PSNODE_EULER) dk[i] = lm * dt[i];
                        else if (METHOD == PSNODE_MIDPOINT) dk[i] = e == 1 ? lm * dt[i] : 0.5f * dt[i] * dyA[i * XP];   // dyA = dy_2
                        else {
                            const float l8 = lm * (dt[i] * 0.125f);
                            if (e == 3) dk[i] = l8;
                            else if (e == 2) dk[i] = fmaf(dt[i], dyA[i * XP], 3.0f * l8);                    // dyA = dy_4
                            else if (e == 1) dk[i] = fmaf(dt[i], dyB[i * XP] - dyA[i * XP], 3.0f * l8);      // dyB = dy_3
                            else dk[i] = l8 + dyB[i * XP];                                                   // dyB = dt/3 (dy_2 - dy_3) + dt dy_4
                        }
                        db4 += dk[i];
                        gs.dkt[(8 * h + i) * XP + m] = dk[i];
                        if (!pf) gs.yt[sc & 1][(8 * h + i) * XP + m] = __ldcs(fr + 3 * PSW_BLOCK + (8 * h + i) * XP + m);
                    }
                    store_tile(dk);
                }
                publish();
                // the next step's held inputs: only now, behind a group barrier, has every thread finished reading this buffer at the end
                // of step j + 1 (the step sizes are read at stage tops only, so they are staged at the top of the step)
                if (e == NST - 1 && j > 1 && wk == 6) stage_held(j - 1);
                const bool has_next = !(j == 1 && e == 0);      // the next stage's record is the one right below this one
                if (pf && has_next) {
                    const float* nx = fr - PSW4_FWD_REC;
                    if (wk == 7) stage_y(nx, (sc + 1) & 1);     // last read two barriers before the previous stage ended
#pragma unroll
                    for (int blk = 0; blk < 3; blk++) {
                        prefetch_l2(nx + blk * PSW_BLOCK + toff);
                        prefetch_l2(nx + blk * PSW_BLOCK + toff + 32);
                    }
                }
                // ---- delta3 = (W4^T dk) * ELU'(a3);  dW4 += dk . a3^T ----
                issue_w4t();
                load_block(fr + 2 * PSW_BLOCK, a);
#pragma unroll
                for (int i = 0; i < 8; i++) {
#pragma unroll
                    for (int kq = 0; kq < XP / 4; kq++) {
                        const float4 v = *reinterpret_cast<const float4*>(&gs.dkt[(8 * h + i) * XP + 4 * kq]);
                        dw4[4 * kq + 0] = fmaf(v.x, a[i], dw4[4 * kq + 0]);
                        dw4[4 * kq + 1] = fmaf(v.y, a[i], dw4[4 * kq + 1]);
                        dw4[4 * kq + 2] = fmaf(v.z, a[i], dw4[4 * kq + 2]);
                        dw4[4 * kq + 3] = fmaf(v.w, a[i], dw4[4 * kq + 3]);
                    }
                }
                collect2(d);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    d[i] *= psn_elu_grad_from_out(a[i]);
                    db3 += d[i];
                }
                store_tile(d);
                tape_block(br + PSW_BLOCK, d);
                publish();
                // ---- delta2 = (W3^T delta3) * ELU'(a2) ----
                issue_w3t();
                load_block(fr + PSW_BLOCK, a);
                collect4(d);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    d[i] *= psn_elu_grad_from_out(a[i]);
                    db2 += d[i];
                }
                store_tile(d);
                tape_block(br, d);
                if (pf && wk == 7) cp_async_wait_all();           // the next stage's y block has landed: visible behind this barrier
                publish();
                // ---- delta1 = (W2^T delta2) * ELU'(a1);  dF_x += delta1 . y^T ----
                issue_w2t();
                load_block(fr, a);
                collect4(d);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    d[i] *= psn_elu_grad_from_out(a[i]);
                    sum1[i] += d[i];
#pragma unroll
                    for (int kq = 0; kq < XP / 4; kq++) {
                        const float4 v = *reinterpret_cast<const float4*>(&gs.yt[sc & 1][(8 * h + i) * XP + 4 * kq]);
                        dfx[4 * kq + 0] = fmaf(v.x, d[i], dfx[4 * kq + 0]);
                        dfx[4 * kq + 1] = fmaf(v.y, d[i], dfx[4 * kq + 1]);
                        dfx[4 * kq + 2] = fmaf(v.z, d[i], dfx[4 * kq + 2]);
                        dfx[4 * kq + 3] = fmaf(v.w, d[i], dfx[4 * kq + 3]);
                    }
                }
                store_tile(d);
                publish();
                // ---- dy_e = F_x^T delta1 on the state rows ----
                issue_fxt();
                collect4(d);
                if (state_row) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        dysum[i * XP] += d[i];
                        if (METHOD == PSNODE_MIDPOINT) dyA[i * XP] = d[i];
                        if (METHOD == PSNODE_RK4) {
                            if (e == 3) dyA[i * XP] = d[i];
                            else if (e == 2) dyB[i * XP] = d[i];
                            else if (e == 1) dyB[i * XP] = fmaf(dt[i] * (float)(1.0 / 3.0), d[i] - dyB[i * XP], dt[i] * dyA[i * XP]);
                        }
                    }
                }
                sc++;
            }
            // ---- end of step j: held-input and constant gradients of the folded layer 1, lambda_{j-1} ----
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const float* zr = &gs.zh[j & 1][(8 * h + i) * ZMAX];
#pragma unroll
                for (int k = 0; k < ZMAX; k++)
                    if (k < Z) dfz[k] = fmaf(sum1[i], zr[k], dfz[k]);
                dcs[i] += sum1[i];
            }
            if (state_row) {
                float gprev[8];
                load_gx(j - 1, gprev);
#pragma unroll
                for (int i = 0; i < 8; i++) lam[i * XP] = (lam[i * XP] + dysum[i * XP]) + gprev[i];
            }
        }
        // ---- lambda_0 = dL/dx[0] ----
        if (live_x && q.d_x0) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int b = b0 + 8 * h + i;
                if (b < B) q.d_x0[(int64_t)b * q.d_x0_sb + m] = lam[i * XP];
            }
        }
        // ---- slab of this (group, column half): the narrow gradients of neuron m over its 8 trajectories ----
        {
            float* slab = q.slabs + ((int64_t)gid * 2 + h) * SLAB_FIELDS * H + m;
#pragma unroll
            for (int k = 0; k < XP; k++) { slab[(F_DFX + k) * H] = dfx[k]; slab[(F_DW4 + k) * H] = dw4[k]; }
#pragma unroll
            for (int k = 0; k < ZMAX; k++) slab[(F_DFZ + k) * H] = dfz[k];
            float db1 = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) db1 += dcs[i];
            for (int k = 0; k < SMAX; k++) {                    // dCA[m][k] = sum_n dc[m][n] a0[n][k]: the W_a / W_b split of layer 1
                float s = 0.0f;
                if (k < S) {
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int bb = min(b0 + 8 * h + i, B - 1);
                        s = fmaf(dcs[i], __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), s);
                    }
                }
                slab[(F_DCA + k) * H] = s;
            }
            slab[F_DB1 * H] = db1;
            slab[F_DB2 * H] = db2;
            slab[F_DB3 * H] = db3;
            slab[F_DB4 * H] = db4;
        }
        // ---- d_a0[b][k] = sum_m dc[b][m] (W_a - W_b)[m][k]: dc through shared memory, one (trajectory, k) pair per thread ----
        if (q.d_a0) {
            float* dct = reinterpret_cast<float*>(gs.act_hi);       // [trajectory][128 neurons]
            group_sync(g);
#pragma unroll
            for (int i = 0; i < 8; i++) dct[(8 * h + i) * H + m] = dcs[i];
            group_sync(g);
            for (int idx = gt; idx < TN * S; idx += GROUP_THREADS) {
                const int n = idx / S, k = idx - n * S;
                const int b = b0 + n;
                if (b >= B) continue;
                float s = 0.0f;
                for (int mm = 0; mm < Hh; mm++) {
                    const float* wr = q.W1 + (int64_t)mm * (3 * S);
                    s = fmaf(dct[n * H + mm], __ldg(wr + k) - __ldg(wr + S + k), s);
                }
                q.d_a0[(int64_t)b * q.d_a0_sb + k] = s;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 512);
}

// d_theta = [dW1 (Hh x 3S) | db1 | dW2 (Hh x Hh) | db2 | dW3 | db3 | dW4 (X x Hh) | db4], every element a fixed-order sum over the slabs
__global__ void psn_wide4_assemble_kernel(const float* __restrict__ slabs, int nslab4, const float* __restrict__ gslabs, int c0, int c1, int c2,
                                          int X, int Z, int Hh, float* __restrict__ d_theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int S = X + Z;
    const int nW1 = Hh * 3 * S, nWh = Hh * Hh, nW4 = X * Hh;
    auto field_sum = [&](int f, int m) {
        float s = 0.0f;
        for (int c = 0; c < nslab4; c++) s += slabs[((int64_t)c * SLAB_FIELDS + f) * H + m];
        return s;
    };
    auto gslab_sum = [&](int from, int to, int m, int k) {
        float s = 0.0f;
        for (int c = from; c < to; c++) s += gslabs[(int64_t)c * H * H + m * H + k];
        return s;
    };
    int o = idx;
    if (o < nW1) {
        const int m = o / (3 * S), col = o - m * 3 * S, blk = col / S, k = col - blk * S;
        const float a = field_sum(F_DCA + k, m);
        if (blk == 0) { d_theta[idx] = a; return; }
        const float gsum = k < X ? field_sum(F_DFX + k, m) : field_sum(F_DFZ + (k - X), m);
        d_theta[idx] = blk == 1 ? gsum - a : gsum;
        return;
    }
    o -= nW1;
    if (o < Hh) { d_theta[idx] = field_sum(F_DB1, o); return; }
    o -= Hh;
    if (o < nWh) { d_theta[idx] = gslab_sum(c0, c1, o / Hh, o % Hh); return; }
    o -= nWh;
    if (o < Hh) { d_theta[idx] = field_sum(F_DB2, o); return; }
    o -= Hh;
    if (o < nWh) { d_theta[idx] = gslab_sum(c1, c2, o / Hh, o % Hh); return; }
    o -= nWh;
    if (o < Hh) { d_theta[idx] = field_sum(F_DB3, o); return; }
    o -= Hh;
    if (o < nW4) { d_theta[idx] = field_sum(F_DW4 + o / Hh, o % Hh); return; }
    o -= nW4;
    if (o < X) d_theta[idx] = field_sum(F_DB4, o);
}

int64_t align64(int64_t floats) { return (floats + 63) & ~(int64_t)63; }

struct Bwd4Layout {
    int64_t err, btape, slabs, gslabs, total;
    int nslab;
};
Bwd4Layout bwd4_layout(const psnode_problem* p) {
    Bwd4Layout L;
    const int64_t steps = p->T > 1 ? p->T - 1 : 0;
    const int64_t ng = psw_ngroups(p->B);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    L.nslab = sms < 2 ? 2 : sms;
    int64_t o = 0;
    L.err = o; o += 64;
    L.btape = o; o += align64(ng * steps * psw_nstages(p->method) * PSW4_BWD_REC);
    L.slabs = o; o += align64(2 * ng * SLAB_FIELDS * PSW_H);
    L.gslabs = o; o += align64((int64_t)L.nslab * PSW_H * PSW_H);
    L.total = o;
    return L;
}

}  // namespace

// PSNODE_WIDE4_BWD=0 keeps the generic recomputing sweep for this shape (and no tape is recorded)
bool psn_wide4_bwd_enabled() {
    const char* e = getenv("PSNODE_WIDE4_BWD");
    return e ? e[0] != '0' : true;
}

bool psn_wide4_bwd_supports(const psnode_problem* p, const psnode_adjoint* a) {
    if (!psn_wide4_bwd_enabled() || !psn_wide4_supports(p)) return false;
    if (p->impl != PSNODE_IMPL_WIDE && !(p->impl == PSNODE_IMPL_AUTO && psn_wide4_auto(p))) return false;      // where the forward recorded this tape
    if (!p->tape || p->tape_floats < psw4_tape_floats(p->B, p->T, p->method)) return false;
    if (a->d_z.p || a->d_zjump || a->d_v.p || a->d_vjump || a->d_xteach.p || a->d_iteach.p) return false;
    if (!a->gx.p && !a->fuse_x.target.p) return false;
    return true;
}

int64_t psn_wide4_backward_workspace(const psnode_problem* p, const psnode_adjoint* a) {
    (void)a;
    return bwd4_layout(p).total * 4;
}

int psn_wide4_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const Bwd4Layout L = bwd4_layout(p);
    if (ws == nullptr || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    float* w = static_cast<float*>(ws);
    int* err = reinterpret_cast<int*>(w + L.err);
    const int X = p->X, Z = p->Z, Hh = p->de.out_dim[0], S = X + Z;
    const int64_t steps = p->T > 1 ? p->T - 1 : 0, ng = psw_ngroups(p->B);
    const int NST = psw_nstages(p->method);
    const int ntheta = Hh * 3 * S + Hh + 2 * (Hh * Hh + Hh) + X * Hh + X;
    if (a->n_theta < ntheta) return PSNODE_EINVAL;
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    Wide4BwdParams q;
    q.B = p->B; q.T = p->T; q.ngroups = (int)ng; q.X = X; q.Z = Z; q.Hh = Hh;
    q.t = p->t; q.z = p->z; q.gx = a->gx;
    q.fx = psn_make_fuse(a->fuse_x, p->x_sol);
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.W1 = p->de.W[0]; q.W2 = p->de.W[1]; q.W3 = p->de.W[2]; q.W4 = p->de.W[3];
    q.tape = p->tape; q.btape = w + L.btape;
    q.slabs = w + L.slabs;
    q.d_x0 = a->d_x0; q.d_x0_sb = a->d_x0_sb;
    q.d_a0 = a->d_a0; q.d_a0_sb = a->d_a0_sb;
    { const char* e = getenv("PSNODE_WIDE4_PF"); q.pf = e ? (e[0] != '0') : 1; }      // A/B at the cfg2 batch: 68.5 -> 57.2 ms per training step
    q.err = err;
    {
        const int grid = (int)((ng + PSW_GROUPS_PER_CTA - 1) / PSW_GROUPS_PER_CTA);
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
            case PSNODE_EULER: st = launch(psn_wide4_bwd_kernel<PSNODE_EULER>, "psn_wide4_bwd_kernel<euler>"); break;
            case PSNODE_MIDPOINT: st = launch(psn_wide4_bwd_kernel<PSNODE_MIDPOINT>, "psn_wide4_bwd_kernel<midpoint>"); break;
            default: st = launch(psn_wide4_bwd_kernel<PSNODE_RK4>, "psn_wide4_bwd_kernel<rk4>"); break;
        }
        if (st != PSNODE_OK) return st;
    }
    // ---- dW2 = sum delta2 . a1^T, dW3 = sum delta3 . a2^T: block GEMMs over the two tapes (psnode_wide_grad.cu) ----
    int cta0[3];
    {
        const int st = psn_wide_grad_pairs(w + L.btape, PSW4_BWD_REC, p->tape, PSW4_FWD_REC,                             // delta2, a1
                                           w + L.btape + PSW_BLOCK, PSW4_BWD_REC, p->tape + PSW_BLOCK, PSW4_FWD_REC,   // delta3, a2
                                           ng * steps * NST, L.nslab, w + L.gslabs, err, cta0, stream);
        if (st != PSNODE_OK) return st;
    }
    psn_wide4_assemble_kernel<<<(ntheta + 255) / 256, 256, 0, stream>>>(w + L.slabs, (int)(2 * ng), w + L.gslabs, cta0[0], cta0[1], cta0[2],
                                                                        X, Z, Hh, a->d_theta);
    psn_count_launch("psn_wide4_assemble_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
