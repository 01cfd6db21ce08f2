// psnode_tc8_fwd.cu -- tensor-core forward integrator, 8 warps per 16-trajectory group (impl = tc8).
//
// Same algorithm, operand layouts, TMEM plan and results as psnode_tc_fwd.cu (integrate_ODE / integrate_DAE of
// neural_dae/my_solvers.py:52-131 for the H = 64 nets, 3xTF32 on tcgen05, weights resident in TMEM); what changes is how the
// epilogue of each layer is spread over threads.  The profile of the 4-warp kernel (profiles/r01_ncu_tc_fwd_cfg2_v2.txt) shows
// the serial chain MMA issue -> MMA completion -> 8-element epilogue per thread (~290 issue slots) -> barrier at 1284 cycles
// per layer with only 44 % of the issue slots used: the kernel is latency bound and B = 4096 trajectories leave just two
// groups per SM to overlap.  Here a group is 256 threads: warps k and k + 4 share TMEM sub-partition k (a warp may only touch
// the 32 lanes of sub-partition warp_id % 4) and split the 16 trajectory columns of every accumulator row 8 / 8, so each
// thread finishes 4 elements per layer instead of 8 and the stage algebra owns ONE state element per thread.  Only warps
// 0..3 of a group issue MMAs (6 per layer each, 4 partial accumulators as before); warp 4 stages the next step's inputs.
// 16 warps per SM instead of 8 -> twice the latency hiding for the same number of MMAs.
//
// The activation tape layout (psnode_tc_tape.cuh) is unchanged: thread (quadrant w, half h, lane) writes the 16 bytes at
// float offset (32 w + lane) * 8 + 4 h of a fragment block, i.e. exactly the elements 4h..4h+3 of the 8-element fragment
// the reverse sweep's thread (w, lane) reads.
#include <cstddef>
#include "psnode_internal.cuh"
#include "psnode_tc.cuh"
#include "psnode_tc_tape.cuh"

namespace {
using namespace psn_tc;

constexpr int TN = 16;                 // trajectories per group (MMA N)
constexpr int TH = 64, TX = 16, TU = 8;
constexpr int TK1 = TX + TU;           // layer-1 K after folding
constexpr int LBO = 144;               // K-chunk stride of the activation tiles (16 B chunk + 128 B row block, padded)
constexpr int SBO_ACT = (TH / 4) * LBO;
constexpr int SBO_B1 = (TK1 / 4) * LBO;
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;
constexpr int B1_TILE = (TN / 8) * SBO_B1;
constexpr int LBO_W = 128, SBO_W = (TK1 / 4) * LBO_W;    // folded layer-1 weight tiles in shared memory (64 rows x K = 24)
constexpr int W1_TILE = (TH / 8) * SBO_W;

// TMEM columns: accumulators first (2 groups x 4 issuing warps x 16), then the resident weights
constexpr int TM_ACC = 0;
constexpr int TM_W2 = 128, TM_W3 = 256, TM_W4 = 384;     // hi at +0, lo at +64
constexpr int TM_COLS = 512;
// An M = 64 operand / accumulator occupies lanes 0..15 of each TMEM sub-partition.  DAE: the AE net's layers 2..4 and their
// accumulators use lanes 16..31 of the SAME columns (a TS MMA needs A and D at the same lanes).
constexpr uint32_t TM_UPPER = 16u << 16;
// ODE: lanes 16..31 are free, so the folded layer 1 (K = 24: hi at column 128, lo at 160) and its accumulators live there too.
constexpr int TM_W1 = 128;
constexpr int GROUP_THREADS = 256;

struct Tc8Params {
    int B, T, X, Z, S, groups;                 // X <= 16 state variables (rows / columns X..15 of the tiles are zero padding)
    int V, I, E;                               // DAE only (0 for an ODE); U = Z + V + I <= 8 held-input columns; E events
    psnode_series t, x, z, v;
    const float* x_init; int64_t x_init_sb;
    const float* v_jump; int64_t vj_sb, vj_se;
    psnode_series_out i_sol;
    const float* A1; const float* ab1; const float* A2; const float* ab2;     // AE net (DAE)
    const float* A3; const float* ab3; const float* A4; const float* ab4;
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
    unsigned char act_hi[ACT_TILE];
    unsigned char act_lo[ACT_TILE];
    unsigned char b1_hi[B1_TILE + 64];
    unsigned char b1_lo[B1_TILE + 64];
    float ostage[TN][TX];
    float istage[TN][TU];
    float dts[2][TN];
    uint64_t bar;
};

struct __align__(128) CtaSmem {
    float w1_hi[W1_TILE / 4];
    float w1_lo[W1_TILE / 4];
    GroupSmem g[2];
    uint32_t tmem_base;
};
struct __align__(128) CtaSmemDae {            // DAE: the AE net's folded layer 1 follows (A operand read from shared memory)
    CtaSmem base;
    float wa1_hi[W1_TILE / 4], wa1_lo[W1_TILE / 4];
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }

template <int METHOD, bool DAE, bool TAPE>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc8_kernel(const __grid_constant__ Tc8Params q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmemDae& smd = *reinterpret_cast<CtaSmemDae*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));   // AE part only if DAE
    CtaSmem& sm = smd.base;
    const int tid = threadIdx.x;
    // warp-level indices come from a shuffle so that ptxas knows they are warp-uniform: descriptors, TMEM addresses and
    // role branches then live in uniform registers instead of being moved there (R2UR) in front of every MMA
    const int cta_warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = cta_warp >> 3;               // group
    const int gt = tid & 255;                  // thread within the group
    const int wk = cta_warp & 7, lane = tid & 31;   // warp within the group
    const int wq = wk & 3, h = wk >> 2;        // TMEM sub-partition (== CTA warp index % 4), column half
    const bool issuer = h == 0;
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, S = q.S;
    const int ZV = Z + (DAE ? q.V : 0);                // columns of the B tile fed from the z / v series
    const int U = ZV + (DAE ? q.I : 0);                // held-input columns (i columns are written by the AE epilogue)
    // tile column c of the layer-1 B tile [x (16, X used) | held inputs (8, U used)] -> index into s = cat(x, z, v, i), or -1
    auto scol = [&](int c) { return c < TX ? (c < X ? c : -1) : (c - TX < U ? X + (c - TX) : -1); };
    const int gid = blockIdx.x * q.groups + g;
    const int b0 = gid * TN;
    const bool live = g < q.groups && b0 < B;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&sm.g[0].bar, 4); mbar_init(&sm.g[1].bar, 4); fence_mbar_init(); }
    if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, TM_COLS);
    {   // folded layer-1 weights (Wb + Wc restricted to [x | held inputs], zero padded to K = 24) -> shared-memory tiles
        const int K1 = 3 * S;
        for (int e = tid; e < TH * TK1; e += 2 * GROUP_THREADS) {
            const int m = e / TK1, c = e - m * TK1;
            float hi = 0.0f, lo = 0.0f;
            const int sc = scol(c);
            if (sc >= 0) split_tf32(__ldg(q.W1 + m * K1 + S + sc) + __ldg(q.W1 + m * K1 + 2 * S + sc), hi, lo);
            sm.w1_hi[tile_byte(m, c, LBO_W, SBO_W) >> 2] = hi;
            sm.w1_lo[tile_byte(m, c, LBO_W, SBO_W) >> 2] = lo;
        }
    }
    if constexpr (DAE) {
        // AE layer 1 folded onto the same B tile: columns [x | z v] carry W[:, S + c], the i columns (and the padding) zero
        const int KA = S + X + ZV;
        for (int e = tid; e < TH * TK1; e += 2 * GROUP_THREADS) {
            const int m = e / TK1, c = e - m * TK1;
            float hi = 0.0f, lo = 0.0f;
            const int sc = scol(c);
            if (sc >= 0 && sc < X + ZV) split_tf32(__ldg(q.A1 + m * KA + S + sc), hi, lo);
            smd.wa1_hi[tile_byte(m, c, LBO_W, SBO_W) >> 2] = hi;
            smd.wa1_lo[tile_byte(m, c, LBO_W, SBO_W) >> 2] = lo;
        }
    }
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

    // resident weights -> TMEM (warps 0..3 of group 0 write; both groups read them through the tensor core only).  W4 (16 rows)
    // is replicated into every 16-row block: row r of the operand holds W4[r & 15].
    if (g == 0 && issuer) {
        for (int half = 0; half < 2; half++) {
            for (int cb = 0; cb < 4; cb++) {               // 64 columns = 4 x 16
                float w2[8], w3[8], w4[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {              // 16x256b.x2 fragment: rows m0 (+8), columns c0 (+1) (+8)
                    const int row = m0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                    float hi, lo;
                    split_tf32(__ldg(q.W2 + row * TH + col), hi, lo); w2[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W3 + row * TH + col), hi, lo); w3[i] = half ? lo : hi;
                    hi = 0.0f; lo = 0.0f;
                    if ((row & 15) < X) split_tf32(__ldg(q.W4 + (row & 15) * TH + col), hi, lo);
                    w4[i] = half ? lo : hi;
                }
                tmem_st_16x256b_x2(tmem + lane_base + TM_W2 + 64 * half + 16 * cb, w2);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W3 + 64 * half + 16 * cb, w3);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W4 + 64 * half + 16 * cb, w4);
            }
        }
        if constexpr (!DAE) {      // folded layer 1 -> lanes 16..31: A[m][c] = (Wb + Wc)[m][c] for c < 16 + U, zero up to column 32
            const int K1 = 3 * S;
            for (int half = 0; half < 2; half++) {
                for (int cb = 0; cb < 2; cb++) {
                    float w1[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int row = m0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                        float hi = 0.0f, lo = 0.0f;
                        const int sc = scol(col);
                        if (col < TK1 && sc >= 0) split_tf32(__ldg(q.W1 + row * K1 + S + sc) + __ldg(q.W1 + row * K1 + 2 * S + sc), hi, lo);
                        w1[i] = half ? lo : hi;
                    }
                    tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W1 + 32 * half + 16 * cb, w1);
                }
            }
        }
        if constexpr (DAE) {       // AE layers 2..4 -> lanes 16..31 (layer 4: I rows replicated into every 16-row block, rest zero)
            for (int half = 0; half < 2; half++) {
                for (int cb = 0; cb < 4; cb++) {
                    float w2[8], w3[8], w4[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int row = m0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                        float hi, lo;
                        split_tf32(__ldg(q.A2 + row * TH + col), hi, lo); w2[i] = half ? lo : hi;
                        split_tf32(__ldg(q.A3 + row * TH + col), hi, lo); w3[i] = half ? lo : hi;
                        hi = 0.0f; lo = 0.0f;
                        if ((row & 15) < q.I) split_tf32(__ldg(q.A4 + (row & 15) * TH + col), hi, lo);
                        w4[i] = half ? lo : hi;
                    }
                    tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W2 + 64 * half + 16 * cb, w2);
                    tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W3 + 64 * half + 16 * cb, w3);
                    tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W4 + 64 * half + 16 * cb, w4);
                }
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();            // resident weights visible to both groups' MMAs
    tc_fence_after();
    // per-thread constants: biases of its two rows, c1 of its 4 (row, trajectory) elements
    float bias2[2], bias3[2], c1[4];           // bias2 doubles as a dummy for layer 1 (c1 carries b1)
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
    float biasA2[2] = {0.f, 0.f}, biasA3[2] = {0.f, 0.f}, biasA4 = 0.0f, c1a[4] = {0.f, 0.f, 0.f, 0.f};   // AE net (DAE)
    if constexpr (DAE) {
        const int KA = S + X + ZV;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            biasA2[r] = __ldg(q.ab2 + m0 + 8 * r);
            biasA3[r] = __ldg(q.ab3 + m0 + 8 * r);
        }
        biasA4 = srow < q.I ? __ldg(q.ab4 + srow) : 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) {       // AE layer 1: the all_initial block is a per-(neuron, trajectory) constant
            const int row = frag_row(i), bb = min(b0 + frag_col(i), B - 1);
            float acc = __ldg(q.ab1 + row);
            if (live)
                for (int k = 0; k < S; k++) acc = fmaf(__ldg(q.A1 + row * KA + k), __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), acc);
            c1a[i] = acc;
        }
    }
    // activation-tile byte offsets of this thread's 4 elements (row = K index of the next layer, column = trajectory)
    int off_act[4];
#pragma unroll
    for (int i = 0; i < 4; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    const int off_x = (int)tile_byte(sn, srow, LBO, SBO_B1);
    // DAE: the algebraic output (row srow of the replicated layer-4 tile) this thread writes into the i columns
    const int off_i = (int)tile_byte(sn, min(TX + ZV + srow, TK1 - 1), LBO, SBO_B1);
    // descriptors
    const uint32_t idesc = make_idesc_tf32(TH, TN);
    // every lo tile directly follows its hi tile: the lo descriptor is the hi one plus a constant in the start-address field,
    // which halves the descriptors ptxas has to keep in uniform registers (it spilled them to vector registers + R2UR chains
    // in front of the MMAs of some instantiations)
    static_assert(offsetof(GroupSmem, act_lo) - offsetof(GroupSmem, act_hi) == ACT_TILE, "act_lo must follow act_hi");
    static_assert(offsetof(GroupSmem, b1_lo) - offsetof(GroupSmem, b1_hi) == B1_TILE + 64, "b1_lo must follow b1_hi");
    static_assert(offsetof(CtaSmem, w1_lo) - offsetof(CtaSmem, w1_hi) == W1_TILE, "w1_lo must follow w1_hi");
    const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = d_act_hi + (uint64_t)(ACT_TILE >> 4);
    const uint64_t d_b1_hi = make_desc(smem_u32(gs.b1_hi), LBO, SBO_B1), d_b1_lo = d_b1_hi + (uint64_t)((B1_TILE + 64) >> 4);
    const uint64_t d_w1_hi = make_desc(smem_u32(sm.w1_hi), LBO_W, SBO_W), d_w1_lo = d_w1_hi + (uint64_t)(W1_TILE >> 4);
    const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4) * TN;      // 4 partial accumulators of this group
    const uint32_t my_acc = acc_base + (uint32_t)wq * TN;                  // the one this (issuing) warp's MMAs write
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
    uint32_t phase = 0;

    // ---- helpers ---------------------------------------------------------------------------------------
    // TS layers (weights in TMEM): issuing warp wq takes K-steps 2wq, 2wq+1 of the three 3xTF32 terms, small terms first
    auto issue_ts = [&](uint32_t w_hi, uint32_t w_lo, uint32_t up = 0u) {      // up = TM_UPPER: the AE net's lanes
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int term = 0; term < 3; term++) {
                    const uint32_t wa = term == 0 ? w_lo : w_hi;
                    const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
                    for (int kk = 0; kk < 2; kk++) {
                        const int ks = 2 * wq + kk;
                        mma_tf32_ts(my_acc + up, tmem + up + wa + 8 * ks, bd + KSTEP_B * ks, idesc, accumulate);
                        accumulate = 1;
                    }
                }
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    // layer 1 (A from shared memory, K = 24): issuing warps 0..2 take one K-step each, warp 3 only commits
    auto issue_l1 = [&](uint64_t a_hi, uint64_t a_lo, bool l1_tmem = false) {
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                if (wq < 3) {
                    uint32_t accumulate = 0;
                    for (int term = 0; term < 3; term++) {
                        const uint64_t bd = term == 1 ? d_b1_lo : d_b1_hi;
                        if (l1_tmem) {          // ODE: folded layer 1 resident in TMEM lanes 16..31
                            const uint32_t wa = TM_W1 + (term == 0 ? 32 : 0);
                            mma_tf32_ts(my_acc + TM_UPPER, tmem + TM_UPPER + wa + 8 * wq, bd + KSTEP_B * wq, idesc, accumulate);
                        } else {
                            const uint64_t ad = term == 0 ? a_lo : a_hi;
                            mma_tf32(my_acc, ad + KSTEP_W * wq, bd + KSTEP_B * wq, idesc, accumulate);
                        }
                        accumulate = 1;
                    }
                }
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    // wait for the group's 4 commits
    auto wait_mma = [&]() {
        if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
        phase ^= 1;
        tc_fence_after();
    };
    // sum the first `nacc` partial accumulators into d[4] (this thread's 2 rows x 2 trajectory columns)
    auto collect = [&](float (&d)[4], int nacc, uint32_t up = 0u) {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc_base + up + lane_base + 8 * h;
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        if (nacc == 4) tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
#pragma unroll
        for (int pr = 0; pr < 2; pr++) {
            const psn_u64 s01 = psn_add2(psn_pack2(t0[2 * pr], t0[2 * pr + 1]), psn_pack2(t1[2 * pr], t1[2 * pr + 1]));
            const psn_u64 s23 = nacc == 4 ? psn_add2(psn_pack2(t2[2 * pr], t2[2 * pr + 1]), psn_pack2(t3[2 * pr], t3[2 * pr + 1]))
                                          : psn_pack2(t2[2 * pr], t2[2 * pr + 1]);
            psn_unpack2(psn_add2(s01, s23), d[2 * pr], d[2 * pr + 1]);
        }
    };
    // layer 4: the slope element (state srow, trajectory sn) of this thread; every 16-row block holds the same 16 x 16 tile
    auto collect_slope = [&](uint32_t up = 0u) {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc_base + up + lane_base + 8 * (wq >> 1);
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
        const int sel = 2 * h + (wq & 1);      // element (row + 8h, column parity) of the 16x256b.x1 fragment
        float s[4];
#pragma unroll
        for (int i = 0; i < 4; i++) s[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
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
            st_f32(gs.act_hi, off_act[i], hi);
            st_f32(gs.act_lo, off_act[i], lo);
        }
        if (TAPE && trec)      // elements 4h..4h+3 of the reverse sweep's 8-element fragment (thread 32 wq + lane)
            __stcs(reinterpret_cast<float4*>(trec + (32 * wq + lane) * 8 + 4 * h), make_float4(a[0], a[1], a[2], a[3]));
    };
    // z / v columns of the layer-1 B tile (warp 4, lane = trajectory): values of grid point jp, or of event k when k >= 0
    // (jump_change_fn, neural_base.py:59-65 / :187-196)
    auto load_zv = [&](int jp, int k, float (&u)[TU]) {
        const int bb = min(b0 + (lane & 15), B - 1);
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, jp, bb, c);
            else if (DAE && c < ZV)
                u[c] = k >= 0 ? __ldg(q.v_jump + (int64_t)bb * q.vj_sb + (int64_t)k * q.vj_se + (c - Z)) : ldser(q.v, jp, bb, c - Z);
        }
    };
    auto store_zv = [&](const float (&u)[TU]) {
        if (lane < TN) {
#pragma unroll
            for (int c = 0; c < TU; c++) {
                if (c < ZV) {
                    float hi, lo;
                    split_tf32_fast(u[c], hi, lo);
                    const int o = tile_byte(lane, TX + c, LBO, SBO_B1);
                    st_f32(gs.b1_hi, o, hi);
                    st_f32(gs.b1_lo, o, lo);
                }
            }
        }
    };
    auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };   // step that ENDS at j
    // DAE: i = ae(x, z, v) on the current layer-1 B tile (AE_Func.forward, neural_01_DAE_01_no_encode.py:74-83); the result goes
    // into the tile's i columns (held input of the next DE evaluations) and, with `stage_out`, into istage (-> i_sol row).
    auto ae_eval = [&](bool stage_out, float* rec = nullptr, bool rec_i0 = false) {     // rec: tape record of this evaluation
        if constexpr (DAE) {
            const uint64_t d_wa1_hi = make_desc(smem_u32(smd.wa1_hi), LBO_W, SBO_W), d_wa1_lo = d_wa1_hi + (uint64_t)(W1_TILE >> 4);
            float d[4];
            issue_l1(d_wa1_hi, d_wa1_lo);
            collect(d, 3);
            store_hidden(d, biasA2, c1a, rec);
            publish();
            issue_ts(TM_W2, TM_W2 + 64, TM_UPPER);
            collect(d, 4, TM_UPPER);
            store_hidden(d, biasA2, nullptr, rec ? rec + PSN_TAPE_FRAG : nullptr);
            publish();
            issue_ts(TM_W3, TM_W3 + 64, TM_UPPER);
            collect(d, 4, TM_UPPER);
            store_hidden(d, biasA3, nullptr, rec ? rec + 2 * PSN_TAPE_FRAG : nullptr);
            publish();
            issue_ts(TM_W4, TM_W4 + 64, TM_UPPER);
            const float kv = collect_slope(TM_UPPER);
            if (srow < q.I) {
                const float iv = kv + biasA4;
                float hi, lo;
                split_tf32_fast(iv, hi, lo);
                st_f32(gs.b1_hi, off_i, hi);
                st_f32(gs.b1_lo, off_i, lo);
                if (stage_out) gs.istage[sn][srow] = iv;
                if (TAPE && rec && rec_i0) __stcs(rec + 3 * PSN_TAPE_FRAG + (32 * wq + lane) * 2 + h, iv);    // re-evaluated i_0
            }
            publish();
        }
    };
    // i_sol row jrow <- istage (staged by the last ae_eval(true))
    auto flush_i = [&](int jrow) {
        if constexpr (DAE) {
            if (gt < TN * q.I) {
                const int n = gt / q.I, c = gt - n * q.I, b = b0 + n;
                if (b < B) q.i_sol.p[(int64_t)jrow * q.i_sol.st + (int64_t)b * q.i_sol.sb + c] = gs.istage[n][c];
            }
        }
    };
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
            const float xv = srow < X ? (DAE ? __ldg(q.x_init + (int64_t)bb * q.x_init_sb + srow) : ldser(q.x, 0, bb, srow)) : 0.0f;
            x0 = xv;
            if (b < B && srow < X) q.x_sol.p[(int64_t)b * q.x_sol.sb + srow] = xv;
            float hi, lo;
            split_tf32_fast(xv, hi, lo);
            st_f32(gs.b1_hi, off_x, hi);
            st_f32(gs.b1_lo, off_x, lo);
        }
        float t_prev = 0.0f;        // warp 4: time of the grid point the staged step size ends at
        if (wk == 4) {
            float u[TU];
            if (DAE) load_zv(0, -1, u);                        // i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95)
            else if (T > 1) load_zv(0, event_of_step(1), u);
            if (DAE || T > 1) store_zv(u);
            if (T > 1) {
                const int bb = min(b0 + (lane & 15), B - 1);
                t_prev = ldser(q.t, 1, bb, 0);
                if (lane < TN) gs.dts[1][lane] = __fsub_rn(t_prev, ldser(q.t, 0, bb, 0));
            }
        }
        // event index of the NEXT step, fetched one step ahead so that no thread ever waits on it
        int k_next = T > 2 ? event_of_step(2) : -1;
        int k_cur = (DAE && T > 1) ? event_of_step(1) : -1;
        publish();
        // DAE tape: records of this group (psnode_tc_tape.cuh)
        float* dbase = (DAE && TAPE && q.tape) ? q.tape + (int64_t)gid * psn_dae_group_recs(T, NST, q.E) * PSN_TAPE_STAGE : nullptr;
        if constexpr (DAE) {
            ae_eval(true, dbase ? dbase + psn_dae_rec_point(0, T, NST) * PSN_TAPE_STAGE : nullptr);
            flush_i(0);
        }

        const float c13 = (float)(1.0 / 3.0);
        float* trec = (TAPE && q.tape) ? (DAE ? dbase : q.tape + (int64_t)gid * (T - 1) * NST * PSN_TAPE_STAGE) : nullptr;
        float ycur = x0;                                    // input of the current stage (recorded on the tape)
        for (int j = 1; j < T; j++) {
            float un[TU] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, tn = 0.0f;   // next step's inputs: raw loads issued by warp 4 during stage 0, first used in the last stage
            const bool have_next = j + 1 < T;
            const float dt = gs.dts[j & 1][sn];
            const int k_after = k_next;                         // event of step j + 1
            k_next = j + 2 < T ? event_of_step(j + 2) : -1;     // in flight until the next iteration reads it
            if constexpr (DAE) {
                const int k = k_cur;
                k_cur = k_after;
                if (k >= 0) {                                  // event: jumped z / v replace the held inputs and i_0 is re-evaluated
                    if (wk == 4) { float uj[TU]; load_zv(j - 1, k, uj); store_zv(uj); }
                    publish();
                    ae_eval(false, dbase ? dbase + psn_dae_rec_event(k, T, NST) * PSN_TAPE_STAGE : nullptr, true);
                }
            }
            int e = 0;
            // Next step's held inputs and step size go into the tile / dts once the step's last layer-1 MMA has completed, in
            // DAE: in the shadow of the layer-2 MMAs of the last stage (warp 4 issues none; cfg3 9.00 -> 8.76 ms).  The ODE kernel and
            // Euler (whose only stage also issued the loads) keep them at the end of the stage (moving them cost the ODE 1.7 %).
            auto stage_next = [&]() {
                if (wk == 4 && e == NST - 1 && (DAE || have_next)) {
                    store_zv(un);
                    if (have_next) {
                        if (lane < TN) gs.dts[(j + 1) & 1][lane] = __fsub_rn(tn, t_prev);
                        t_prev = tn;
                    }
                }
            };
#pragma unroll 1
            for (e = 0; e < NST; e++) {
                float d[4];
                if (TAPE && trec) __stcs(trec + 3 * PSN_TAPE_FRAG + (32 * wq + lane) * 2 + h, ycur);
                // ---- layer 1 (shared-memory weights) ----
                issue_l1(d_w1_hi, d_w1_lo, !DAE);
                if (e == 0 && wk == 4) {
                    // next step's held inputs; DAE: the un-jumped z[j], v[j] (they feed i_j first), ODE: jumped if step j+1 fires
                    if (DAE) load_zv(j, -1, un);
                    else if (have_next) load_zv(j, k_after, un);
                    if (have_next) tn = ldser(q.t, j + 1, min(b0 + (lane & 15), B - 1), 0);
                }
                if (e == 0 && j > 1) { store_x_row(j - 1); flush_i(j - 1); }    // rows staged by the previous step
                collect(d, 3, DAE ? 0u : TM_UPPER);
                store_hidden(d, bias2, c1, trec);
                publish();
                // ---- layer 2 ----
                issue_ts(TM_W2, TM_W2 + 64);
                if (DAE && NST > 1) stage_next();
                collect(d, 4);
                store_hidden(d, bias2, nullptr, trec ? trec + PSN_TAPE_FRAG : nullptr);
                publish();
                // ---- layer 3 ----
                issue_ts(TM_W3, TM_W3 + 64);
                collect(d, 4);
                store_hidden(d, bias3, nullptr, trec ? trec + 2 * PSN_TAPE_FRAG : nullptr);
                publish();
                // ---- layer 4 + stage algebra: one state element per thread ----
                issue_ts(TM_W4, TM_W4 + 64);
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
                    st_f32(gs.b1_hi, off_x, hi);
                    st_f32(gs.b1_lo, off_x, lo);
                }
                ycur = xn;
                if (last) { x0 = xn; gs.ostage[sn][srow] = xn; }
                if (TAPE && trec) trec += PSN_TAPE_STAGE;
                if (!DAE || NST == 1) stage_next();
                publish();
            }
            if constexpr (DAE) {                                // i_j = ae(x_j, z[j], v[j])  (my_solvers.py:121)
                ae_eval(true, (TAPE && trec) ? trec : nullptr);         // its record follows the step's stage records
                if (TAPE && trec) trec += PSN_TAPE_STAGE;
            }
        }
        if (T > 1) { store_x_row(T - 1); flush_i(T - 1); }
    }
    // ---- teardown --------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

bool psn_tc_supports(const psnode_problem* p);

int psn_tc8_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < 4) return PSNODE_EWORKSPACE;
    const bool dae = p->kind == PSNODE_DAE;
    Tc8Params q;
    q.B = p->B; q.T = p->T; q.X = p->X; q.Z = p->Z; q.S = p->X + p->Z + p->V + p->I;
    q.V = p->V; q.I = p->I; q.E = p->event_idx ? p->E : 0;
    q.t = p->t; q.x = p->x; q.z = p->z; q.v = p->v;
    q.x_init = p->x_init; q.x_init_sb = p->x_init_sb;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.v_jump = p->v_jump; q.vj_sb = p->vj_sb; q.vj_se = p->vj_se;
    q.x_sol = p->x_sol;
    q.i_sol = p->i_sol;
    q.W1 = p->de.W[0]; q.b1 = p->de.b[0]; q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.W3 = p->de.W[2]; q.b3 = p->de.b[2]; q.W4 = p->de.W[3]; q.b4 = p->de.b[3];
    q.A1 = p->ae.W[0]; q.ab1 = p->ae.b[0]; q.A2 = p->ae.W[1]; q.ab2 = p->ae.b[1];
    q.A3 = p->ae.W[2]; q.ab3 = p->ae.b[2]; q.A4 = p->ae.W[3]; q.ab4 = p->ae.b[3];
    q.vec_out = (p->X == TX && (reinterpret_cast<uintptr_t>(p->x_sol.p) & 15) == 0 && (p->x_sol.st & 3) == 0 && (p->x_sol.sb & 3) == 0) ? 1 : 0;
    const int64_t tape_need = dae ? psn_tc_dae_tape_floats(p->B, p->T, p->method, q.E) : psn_tc_tape_floats(p->B, p->T, p->method);
    q.tape = (p->tape && p->tape_floats >= tape_need) ? p->tape : nullptr;
    q.err = static_cast<int*>(ws);
    PSN_CUDA(cudaMemsetAsync(q.err, 0, 4, stream));
    const int ngroups = psn_tc_ngroups(p->B);
    q.groups = psn_tc_groups_per_cta(p->B);
    const int grid = (ngroups + q.groups - 1) / q.groups;
    const int smem = (int)(dae ? sizeof(CtaSmemDae) : sizeof(CtaSmem)) + 128;
    auto launch = [&](auto kern, const char* name) -> int {
        PSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, 2 * GROUP_THREADS, smem, stream>>>(q);
        psn_count_launch(name);
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    };
    if (dae && q.tape) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_tc8_kernel<PSNODE_EULER, true, true>, "psn_tc8_dae_kernel<euler,tape>");
            case PSNODE_MIDPOINT: return launch(psn_tc8_kernel<PSNODE_MIDPOINT, true, true>, "psn_tc8_dae_kernel<midpoint,tape>");
            default: return launch(psn_tc8_kernel<PSNODE_RK4, true, true>, "psn_tc8_dae_kernel<rk4,tape>");
        }
    }
    if (dae) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_tc8_kernel<PSNODE_EULER, true, false>, "psn_tc8_dae_kernel<euler>");
            case PSNODE_MIDPOINT: return launch(psn_tc8_kernel<PSNODE_MIDPOINT, true, false>, "psn_tc8_dae_kernel<midpoint>");
            default: return launch(psn_tc8_kernel<PSNODE_RK4, true, false>, "psn_tc8_dae_kernel<rk4>");
        }
    }
    if (q.tape) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_tc8_kernel<PSNODE_EULER, false, true>, "psn_tc8_ode_kernel<euler,tape>");
            case PSNODE_MIDPOINT: return launch(psn_tc8_kernel<PSNODE_MIDPOINT, false, true>, "psn_tc8_ode_kernel<midpoint,tape>");
            default: return launch(psn_tc8_kernel<PSNODE_RK4, false, true>, "psn_tc8_ode_kernel<rk4,tape>");
        }
    }
    switch (p->method) {
        case PSNODE_EULER: return launch(psn_tc8_kernel<PSNODE_EULER, false, false>, "psn_tc8_ode_kernel<euler>");
        case PSNODE_MIDPOINT: return launch(psn_tc8_kernel<PSNODE_MIDPOINT, false, false>, "psn_tc8_ode_kernel<midpoint>");
        default: return launch(psn_tc8_kernel<PSNODE_RK4, false, false>, "psn_tc8_ode_kernel<rk4>");
    }
}
