// NEGATIVE EXPERIMENT (round 1, kept for the record; not part of the build) -- psn_tc8_kernel for integrate_ODE with the three
// 3xTF32 products of a K-step stacked into ONE tcgen05.mma of shape M = 128, N = 32:
//     A (TMEM, 128 lanes) = [ W_hi ; W_lo ]      lanes 0..15 of each sub-partition: W_hi rows, lanes 16..31: W_lo rows
//     B (shared memory, 32 rows) = [ a_hi ; a_lo ] per 8-trajectory block: rows  hi(0..7) lo(0..7) hi(8..15) lo(8..15)
//     D (TMEM, 128 lanes x 32 columns):  lanes 0..15 = [ W_hi a_hi | W_hi a_lo ],  lanes 16..31 = [ W_lo a_hi | (W_lo a_lo, unused) ]
// i.e. 8 instructions per 64 x 64 layer instead of 24.  Parity-green (all 166 GPU tests) but SLOWER on B200 at cfg2:
// 8.95 ms vs 7.18 ms.  Cycle stamps (clock64, -DPSN_EXP=20; per layer and 16-trajectory group, issuing warp):
//     this build : MMA issue 321 (2 instr)  completion wait 222   TMEM loads + 3-block sum 319   ELU + tile store 326   fence + barrier 120
//     3-MMA build: MMA issue 540 (6 instr)  completion wait 100   TMEM loads + sum          110   ELU + tile store 296   fence + barrier 107
// The tensor-pipe time of an instruction scales with its N (and the 16 KB accumulator tile it read-modify-writes), so the
// MMA phase shrinks only 640 -> 543 cycles, while the epilogue now needs 8 tcgen05.ld per layer (4 x2 + 4 x1) instead of 4.
// To build it: copy next to psnode_tc8_fwd.cu, declare psn_tc8_ode_forward in psnode_internal.cuh and call it from
// psn_tc8_forward for PSNODE_ODE.
#include <cstddef>
#include <cstdio>
#ifndef PSN_EXP
#define PSN_EXP 0
#endif
#include "psnode_internal.cuh"
#include "psnode_tc.cuh"
#include "psnode_tc_tape.cuh"

namespace {
using namespace psn_tc;

constexpr int TN = 16;                 // trajectories per group
constexpr int TH = 64, TX = 16, TU = 8;
constexpr int TK1 = TX + TU;           // layer-1 K after folding
constexpr int LBO = 144;               // K-chunk stride of the activation tiles (16 B chunk + 128 B row block, padded)
constexpr int SBO_ACT = (TH / 4) * LBO;        // one 8-row block of a K = 64 tile
constexpr int SBO_B1 = (TK1 / 4) * LBO;        // one 8-row block of the K = 24 layer-1 tile
constexpr int ACT_TILE = 4 * SBO_ACT;          // 32 rows: hi(0..7) lo(0..7) hi(8..15) lo(8..15)
constexpr int B1_TILE = 4 * SBO_B1;
constexpr int NQ = 2 * TN;             // MMA N
constexpr int NPART = 4;               // K-split partial accumulators (one per issuing warp)

// TMEM columns (all 32 lanes of every sub-partition are used: hi parts in lanes 0..15, lo parts in lanes 16..31)
constexpr int TM_ACC = 0;                          // 2 groups x NPART x 32
constexpr int TM_W2 = 2 * NPART * NQ;              // 256
constexpr int TM_W3 = TM_W2 + TH, TM_W4 = TM_W3 + TH, TM_W1 = TM_W4 + TH;      // 320, 384, 448 (+32)
constexpr int TM_COLS = 512;
constexpr uint32_t TM_UPPER = 16u << 16;
constexpr int GROUP_THREADS = 256;

struct OdeParams {
    int B, T, X, Z, S, groups;                 // X <= 16 state variables (rows / columns X..15 of the tiles are zero padding)
    psnode_series t, x, z;
    const float* a0; int64_t a0_sb;
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    psnode_series_out x_sol;
    const float* W1; const float* b1; const float* W2; const float* b2;
    const float* W3; const float* b3; const float* W4; const float* b4;
    int vec_out;
    float* tape;
    int* err;
};

struct __align__(128) GroupSmem {
    unsigned char act[ACT_TILE];
    unsigned char b1t[B1_TILE + 64];
    float ostage[TN][TX];
    float dts[2][TN];
    uint64_t bar;
};

struct __align__(128) CtaSmem {
    GroupSmem g[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }
// B-tile row of trajectory n: hi part (the lo part is 8 rows = one SBO further)
__device__ __forceinline__ int brow(int n) { return (n & 7) + 16 * (n >> 3); }

template <int METHOD, bool TAPE>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc8_ode_kernel(const __grid_constant__ OdeParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    // warp-level indices through a shuffle: ptxas then keeps descriptors / TMEM addresses in uniform registers
    const int cta_warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = cta_warp >> 3;               // group
    const int gt = tid & 255;                  // thread within the group
    const int wk = cta_warp & 7, lane = tid & 31;
    const int wq = wk & 3, h = wk >> 2;        // TMEM sub-partition, trajectory half
    const bool issuer = h == 0;
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, S = q.S;
    // tile column c of the layer-1 B tile [x (16, X used) | held inputs (8, Z used)] -> index into s = cat(x, z), or -1
    auto scol = [&](int c) { return c < TX ? (c < X ? c : -1) : (c - TX < Z ? X + (c - TX) : -1); };
    const int gid = blockIdx.x * q.groups + g;
    const int b0 = gid * TN;
    const bool live = g < q.groups && b0 < B;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&sm.g[0].bar, 4); mbar_init(&sm.g[1].bar, 4); fence_mbar_init(); }
    if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, TM_COLS);
    for (int e = gt; e < (int)(offsetof(GroupSmem, bar) / 4); e += GROUP_THREADS) reinterpret_cast<float*>(&gs)[e] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    // this thread's accumulator fragment: element i (0..3) <-> (row m0 + 8*(i>>1), trajectory 8h + c0 + (i&1))
    const int m0 = 16 * wq + (lane >> 2), c0 = 2 * (lane & 3);
    auto frag_row = [&](int i) { return m0 + (i >> 1) * 8; };
    auto frag_col = [&](int i) { return 8 * h + c0 + (i & 1); };

    // resident weights -> TMEM: hi parts into lanes 0..15, lo parts into lanes 16..31 of the same columns (warps 0..3 of
    // group 0 write).  W4 (X <= 16 rows) is replicated into every 16-row block; the folded layer 1 is
    // (Wb + Wc)[:, x | held inputs], zero padded to 32 columns.
    if (g == 0 && issuer) {
        const int K1 = 3 * S;
        for (int cb = 0; cb < 4; cb++) {               // 64 columns = 4 x 16
            float h2[8], l2[8], h3[8], l3[8], h4[8], l4[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {              // 16x256b.x2 fragment: rows m0 (+8), columns c0 (+1) (+8)
                const int row = m0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                split_tf32(__ldg(q.W2 + row * TH + col), h2[i], l2[i]);
                split_tf32(__ldg(q.W3 + row * TH + col), h3[i], l3[i]);
                h4[i] = 0.0f; l4[i] = 0.0f;
                if ((row & 15) < X) split_tf32(__ldg(q.W4 + (row & 15) * TH + col), h4[i], l4[i]);
            }
            tmem_st_16x256b_x2(tmem + lane_base + TM_W2 + 16 * cb, h2);
            tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W2 + 16 * cb, l2);
            tmem_st_16x256b_x2(tmem + lane_base + TM_W3 + 16 * cb, h3);
            tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W3 + 16 * cb, l3);
            tmem_st_16x256b_x2(tmem + lane_base + TM_W4 + 16 * cb, h4);
            tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W4 + 16 * cb, l4);
        }
        for (int cb = 0; cb < 2; cb++) {
            float h1[8], l1[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = m0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                h1[i] = 0.0f; l1[i] = 0.0f;
                const int sc = scol(col);
                if (col < TK1 && sc >= 0) split_tf32(__ldg(q.W1 + row * K1 + S + sc) + __ldg(q.W1 + row * K1 + 2 * S + sc), h1[i], l1[i]);
            }
            tmem_st_16x256b_x2(tmem + lane_base + TM_W1 + 16 * cb, h1);
            tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W1 + 16 * cb, l1);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();            // resident weights visible to both groups' MMAs
    tc_fence_after();
    // per-thread constants: biases of its two rows, c1 = b1 + (Wa - Wb) a0 of its 4 (row, trajectory) elements
    float bias2[2], bias3[2], c1[4];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        bias2[r] = __ldg(q.b2 + m0 + 8 * r);
        bias3[r] = __ldg(q.b3 + m0 + 8 * r);
    }
    {
        const int K1 = 3 * S;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = frag_row(i), bb = min(b0 + frag_col(i), B - 1);
            float acc = __ldg(q.b1 + row);
            if (live)
                for (int k = 0; k < S; k++)
                    acc = fmaf(__ldg(q.W1 + row * K1 + k) - __ldg(q.W1 + row * K1 + S + k), __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), acc);
            c1[i] = acc;
        }
    }
    // the state element this thread owns in the layer-4 epilogue: state srow of trajectory column sn
    const int srow = (lane >> 2) + 8 * h, sn = c0 + (wq & 1) + 8 * (wq >> 1);
    const float bias4 = srow < X ? __ldg(q.b4 + srow) : 0.0f;
    // activation-tile byte offsets (hi part; lo part SBO further) of this thread's 4 elements: row = trajectory, column = neuron
    int off_act[4];
#pragma unroll
    for (int i = 0; i < 4; i++) off_act[i] = tile_byte(brow(frag_col(i)), frag_row(i), LBO, SBO_ACT);
    const int off_x = (int)tile_byte(brow(sn), srow, LBO, SBO_B1);
    // descriptors
    const uint32_t idesc = make_idesc_tf32(128, NQ);
    const uint64_t d_act = make_desc(smem_u32(gs.act), LBO, SBO_ACT), d_b1 = make_desc(smem_u32(gs.b1t), LBO, SBO_B1);
    const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * NPART) * NQ;      // partial accumulators of this group
    const uint32_t my_acc = acc_base + (uint32_t)wq * NQ;                      // the one this (issuing) warp's MMAs write
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4);
    uint32_t phase = 0;

    // ---- helpers ---------------------------------------------------------------------------------------
    // 64 x 64 layer: issuing warp wq takes K-steps 2wq, 2wq + 1 (one stacked 3xTF32 instruction each)
    auto issue_layer = [&](uint32_t w_col) {
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                mma_tf32_ts(my_acc, tmem + w_col + 16 * wq, d_act + KSTEP_B * (2 * wq), idesc, 0u);
                mma_tf32_ts(my_acc, tmem + w_col + 16 * wq + 8, d_act + KSTEP_B * (2 * wq + 1), idesc, 1u);
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    // folded layer 1 (K = 24): issuing warps 0..2 take one K-step each, warp 3 only commits
    auto issue_l1 = [&]() {
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                if (wq < 3) mma_tf32_ts(my_acc, tmem + TM_W1 + 8 * wq, d_b1 + KSTEP_B * wq, idesc, 0u);
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    auto wait_mma = [&]() {
        if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
        phase ^= 1;
        tc_fence_after();
    };
    // sum the first `nacc` partial accumulators into d[4] (this thread's 2 rows x 2 trajectory columns): per partial the
    // lanes 0..15 block [W_hi a_hi | W_hi a_lo] of its 8-trajectory half (16 columns) and the lanes 16..31 block W_lo a_hi
    auto collect_nowait = [&](float (&d)[4], int nacc) {
        float lo8[NPART][8], up4[NPART][4];
        const uint32_t a = acc_base + lane_base + 16 * h;
#pragma unroll
        for (int p = 0; p < NPART; p++) {
            if (p < nacc) {
                tmem_ld_16x256b_x2(a + p * NQ, lo8[p]);
                tmem_ld_16x256b_x1(a + TM_UPPER + p * NQ, up4[p]);
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float small = 0.0f, big = 0.0f;
#pragma unroll
            for (int p = 0; p < NPART; p++) {
                if (p < nacc) {
                    small += lo8[p][4 + i] + up4[p][i];
                    big += lo8[p][i];
                }
            }
            d[i] = big + small;
        }
    };
    auto collect = [&](float (&d)[4], int nacc) {
        wait_mma();
        collect_nowait(d, nacc);
    };
    // layer 4: the slope element (state srow, trajectory sn) of this thread; every 16-row block holds the same 16 x 32 tile
    auto collect_slope = [&]() {
        wait_mma();
        float lo8[NPART][8], up4[NPART][4];
        const uint32_t a = acc_base + lane_base + 16 * (wq >> 1);
#pragma unroll
        for (int p = 0; p < NPART; p++) {
            tmem_ld_16x256b_x2(a + p * NQ, lo8[p]);
            tmem_ld_16x256b_x1(a + TM_UPPER + p * NQ, up4[p]);
        }
        tmem_ld_wait();
        const int sel = 2 * h + (wq & 1);      // element (row + 8h, column parity) of the fragment
        float s[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float small = 0.0f, big = 0.0f;
#pragma unroll
            for (int p = 0; p < NPART; p++) {
                small += lo8[p][4 + i] + up4[p][i];
                big += lo8[p][i];
            }
            s[i] = big + small;
        }
        return sel == 0 ? s[0] : (sel == 1 ? s[1] : (sel == 2 ? s[2] : s[3]));
    };
    // publish freshly written B-tile data to the tensor core and line the group up for the next layer's MMAs
    auto publish = [&]() {
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
    };
    auto store_hidden = [&](const float (&d)[4], const float (&bias)[2], const float* cadd, float* trec) {
        float a[4];
#pragma unroll
        for (int pr = 0; pr < 2; pr++) {       // the two trajectory columns of one row share the bias: packed f32x2 arithmetic
            const psn_u64 dd = psn_pack2(d[2 * pr], d[2 * pr + 1]);
            const psn_u64 vv = psn_add2(dd, cadd ? psn_pack2(cadd[2 * pr], cadd[2 * pr + 1]) : psn_dup2(bias[pr]));
            float v0, v1;
            psn_unpack2(vv, v0, v1);
            psn_elu2(v0, v1, a[2 * pr], a[2 * pr + 1]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float hi, lo;
            split_tf32_fast(a[i], hi, lo);
            st_f32(gs.act, off_act[i], hi);
            st_f32(gs.act, off_act[i] + SBO_ACT, lo);
        }
        if (TAPE && trec)      // elements 4h..4h+3 of the reverse sweep's 8-element fragment (thread 32 wq + lane)
            __stcs(reinterpret_cast<float4*>(trec + (32 * wq + lane) * 8 + 4 * h), make_float4(a[0], a[1], a[2], a[3]));
    };
    // z columns of the layer-1 B tile (warp 4, lane = trajectory): values of grid point jp, or of event k when k >= 0
    // (jump_change_fn, neural_base.py:59-65)
    auto load_z = [&](int jp, int k, float (&u)[TU]) {
        const int bb = min(b0 + (lane & 15), B - 1);
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, jp, bb, c);
        }
    };
    auto load_dt = [&](int j) {      // step that ENDS at grid point j
        const int bb = min(b0 + (lane & 15), B - 1);
        return __fsub_rn(ldser(q.t, j, bb, 0), ldser(q.t, j - 1, bb, 0));
    };
    auto store_z = [&](const float (&u)[TU]) {
        if (lane < TN) {
#pragma unroll
            for (int c = 0; c < TU; c++) {
                if (c < Z) {
                    float hi, lo;
                    split_tf32_fast(u[c], hi, lo);
                    const int o = tile_byte(brow(lane), TX + c, LBO, SBO_B1);
                    st_f32(gs.b1t, o, hi);
                    st_f32(gs.b1t, o + SBO_B1, lo);
                }
            }
        }
    };
    auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };   // step that ENDS at j
    auto store_x_row = [&](int jrow) {      // trajectory row jrow of the group: 16 x 64 B as 128-bit stores (X = 16), else X floats per trajectory
        if (q.vec_out) {
            if (gt < 64) {
                const int n = gt >> 2, c4 = gt & 3, b = b0 + n;
                if (b < B)
                    *reinterpret_cast<float4*>(q.x_sol.p + (int64_t)jrow * q.x_sol.st + (int64_t)b * q.x_sol.sb + 4 * c4) =
                        *reinterpret_cast<const float4*>(&gs.ostage[n][4 * c4]);
            }
        } else if (gt < TN * X) {
            const int n = gt / X, c = gt - n * X, b = b0 + n;
            if (b < B) q.x_sol.p[(int64_t)jrow * q.x_sol.st + (int64_t)b * q.x_sol.sb + c] = gs.ostage[n][c];
        }
    };

    if (live) {
        // ---- initial state: every thread owns one state element (state srow of trajectory column sn) -------
        float x0, k1 = 0.f, k2 = 0.f, k3 = 0.f;
        {
            const int b = b0 + sn, bb = min(b, B - 1);
            const float xv = srow < X ? ldser(q.x, 0, bb, srow) : 0.0f;
            x0 = xv;
            if (b < B && srow < X) q.x_sol.p[(int64_t)b * q.x_sol.sb + srow] = xv;
            float hi, lo;
            split_tf32_fast(xv, hi, lo);
            st_f32(gs.b1t, off_x, hi);
            st_f32(gs.b1t, off_x + SBO_B1, lo);
        }
        if (wk == 4 && T > 1) {
            float u[TU];
            load_z(0, event_of_step(1), u);
            store_z(u);
            if (lane < TN) gs.dts[1][lane] = load_dt(1);
        }
        publish();

        const float c13 = (float)(1.0 / 3.0);
        float* trec = (TAPE && q.tape) ? q.tape + (int64_t)gid * (T - 1) * NST * PSN_TAPE_STAGE : nullptr;
        float ycur = x0;                                    // input of the current stage (recorded on the tape)
#if PSN_EXP == 20
        long long prof[5] = {0, 0, 0, 0, 0}; int nprof = 0;
        const long long tstart = clock64();
#endif
        for (int j = 1; j < T; j++) {
            float un[TU] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dtn = 0.0f;   // next step's inputs, prefetched by warp 4 during stage 0
            const bool have_next = j + 1 < T;
            const float dt = gs.dts[j & 1][sn];
#pragma unroll 1
            for (int e = 0; e < NST; e++) {
                float d[4];
                if (TAPE && trec) __stcs(trec + 3 * PSN_TAPE_FRAG + (32 * wq + lane) * 2 + h, ycur);
                // ---- layer 1 ----
                issue_l1();
                if (e == 0 && wk == 4 && have_next) {       // next step's held inputs (jumped if step j+1 fires) and step size
                    load_z(j, event_of_step(j + 1), un);
                    dtn = load_dt(j + 1);
                }
                if (e == 0 && j > 1) store_x_row(j - 1);    // row staged by the previous step
                collect(d, 3);
                store_hidden(d, bias2, c1, trec);
                publish();
                // ---- layer 2 ----
#if PSN_EXP == 20
                {
                    const long long t0 = clock64();
                    issue_layer(TM_W2);
                    const long long t1 = clock64();
                    wait_mma();
                    const long long t2 = clock64();
                    collect_nowait(d, NPART);
                    const long long t3 = clock64();
                    store_hidden(d, bias2, nullptr, trec ? trec + PSN_TAPE_FRAG : nullptr);
                    const long long t4 = clock64();
                    publish();
                    const long long t5 = clock64();
                    prof[0] += t1 - t0; prof[1] += t2 - t1; prof[2] += t3 - t2; prof[3] += t4 - t3; prof[4] += t5 - t4; nprof++;
                }
#else
                issue_layer(TM_W2);
                collect(d, NPART);
                store_hidden(d, bias2, nullptr, trec ? trec + PSN_TAPE_FRAG : nullptr);
                publish();
#endif
                // ---- layer 3 ----
                issue_layer(TM_W3);
                collect(d, NPART);
                store_hidden(d, bias3, nullptr, trec ? trec + 2 * PSN_TAPE_FRAG : nullptr);
                publish();
                // ---- layer 4 + stage algebra: one state element per thread ----
                issue_layer(TM_W4);
                const float kk = collect_slope() + bias4;
                const bool last = e == NST - 1;
                float xn;
                if (METHOD == PSNODE_EULER) {
                    xn = __fadd_rn(x0, __fmul_rn(dt, kk));
                } else if (METHOD == PSNODE_MIDPOINT) {
                    if (e == 0) xn = __fadd_rn(x0, __fmul_rn(kk, __fmul_rn(0.5f, dt)));
                    else xn = __fadd_rn(x0, __fmul_rn(dt, kk));
                } else {
                    if (e == 0) { k1 = kk; xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(dt, kk), c13)); }
                    else if (e == 1) { k2 = kk; xn = __fadd_rn(x0, __fmul_rn(dt, __fsub_rn(kk, __fmul_rn(k1, c13)))); }
                    else if (e == 2) { k3 = kk; xn = __fadd_rn(x0, __fmul_rn(dt, __fadd_rn(__fsub_rn(k1, k2), kk))); }
                    else {
                        const float ksum = __fadd_rn(__fadd_rn(k1, __fmul_rn(3.0f, __fadd_rn(k2, k3))), kk);
                        xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(ksum, dt), 0.125f));
                    }
                }
                {
                    float hi, lo;
                    split_tf32_fast(xn, hi, lo);
                    st_f32(gs.b1t, off_x, hi);
                    st_f32(gs.b1t, off_x + SBO_B1, lo);
                }
                ycur = xn;
                if (last) { x0 = xn; gs.ostage[sn][srow] = xn; }
                if (TAPE && trec) trec += PSN_TAPE_STAGE;
                if (wk == 4 && last && have_next) {         // all layer-1 MMAs of this step are done
                    store_z(un);
                    if (lane < TN) gs.dts[(j + 1) & 1][lane] = dtn;
                }
                publish();
            }
        }
#if PSN_EXP == 20
        if (blockIdx.x == 3 && (gt == 0 || gt == 5 * 32) && T > 100)
            printf("g%d w%d stages %d cycles/stage %lld | L2: issue %lld wait %lld ld+sum %lld elu+st %lld pub %lld\n", g, wk, nprof,
                   (clock64() - tstart) / nprof, prof[0] / nprof, prof[1] / nprof, prof[2] / nprof, prof[3] / nprof, prof[4] / nprof);
#endif
        if (T > 1) store_x_row(T - 1);
    }
    // ---- teardown --------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

int psn_tc8_ode_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < 4) return PSNODE_EWORKSPACE;
    if (p->kind != PSNODE_ODE) return PSNODE_EUNSUPPORTED;
    OdeParams q;
    q.B = p->B; q.T = p->T; q.X = p->X; q.Z = p->Z; q.S = p->X + p->Z;
    q.t = p->t; q.x = p->x; q.z = p->z;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.x_sol = p->x_sol;
    q.W1 = p->de.W[0]; q.b1 = p->de.b[0]; q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.W3 = p->de.W[2]; q.b3 = p->de.b[2]; q.W4 = p->de.W[3]; q.b4 = p->de.b[3];
    q.vec_out = (p->X == TX && (reinterpret_cast<uintptr_t>(p->x_sol.p) & 15) == 0 && (p->x_sol.st & 3) == 0 && (p->x_sol.sb & 3) == 0) ? 1 : 0;
    const int64_t tape_need = psn_tc_tape_floats(p->B, p->T, p->method);
    q.tape = (p->tape && p->tape_floats >= tape_need) ? p->tape : nullptr;
    q.err = static_cast<int*>(ws);
    PSN_CUDA(cudaMemsetAsync(q.err, 0, 4, stream));
    const int ngroups = psn_tc_ngroups(p->B);
    q.groups = psn_tc_groups_per_cta(p->B);
    const int grid = (ngroups + q.groups - 1) / q.groups;
    const int smem = (int)sizeof(CtaSmem) + 128;
    auto launch = [&](auto kern, const char* name) -> int {
        PSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, 2 * GROUP_THREADS, smem, stream>>>(q);
        psn_count_launch(name);
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    };
    if (q.tape) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_tc8_ode_kernel<PSNODE_EULER, true>, "psn_tc8_ode_kernel<euler,tape>");
            case PSNODE_MIDPOINT: return launch(psn_tc8_ode_kernel<PSNODE_MIDPOINT, true>, "psn_tc8_ode_kernel<midpoint,tape>");
            default: return launch(psn_tc8_ode_kernel<PSNODE_RK4, true>, "psn_tc8_ode_kernel<rk4,tape>");
        }
    }
    switch (p->method) {
        case PSNODE_EULER: return launch(psn_tc8_ode_kernel<PSNODE_EULER, false>, "psn_tc8_ode_kernel<euler>");
        case PSNODE_MIDPOINT: return launch(psn_tc8_ode_kernel<PSNODE_MIDPOINT, false>, "psn_tc8_ode_kernel<midpoint>");
        default: return launch(psn_tc8_ode_kernel<PSNODE_RK4, false>, "psn_tc8_ode_kernel<rk4>");
    }
}
