// psnode_tc_fwd.cu -- tensor-core forward integrator for the reference's H = 64 nets (BASELINE configs[1] and [2]):
// integrate_ODE (neural_dae/my_solvers.py:52-80) with the 4-layer DE_Func of neural_00_ODE_01_no_encode.py:58-68
// (3S -> 64 -> 64 -> 64 -> 16, ELU) and integrate_DAE (my_solvers.py:82-131) with DE_Func + AE_Func of
// neural_01_DAE_01_no_encode.py:61-83 (one explicit algebraic evaluation per step, one more on event steps),
// Euler / Midpoint / RK4-3/8 (neural_dae/my_fixed_grid.py:15-59), event jumps.
//
// Every stage MLP layer is a genuine 64 x 16 x K GEMM per group of 16 trajectories, so it runs on the 5th-generation
// tensor cores: tcgen05.mma kind::tf32 with fp32 accumulation in TMEM.  The reference is fp32 and the parity tolerance
// (rtol 1e-5 / atol 1e-6 over 1000 steps) leaves no room for plain TF32 (measured 3e-4 per dot product), so every
// product is formed as 3xTF32:  W a ~= W_lo a_hi + W_hi a_lo + W_hi a_hi  with round-to-nearest hi/lo splits -- measured
// 1.4e-7 max error per 64-term dot product on B200, the same as an fp32 FMA chain (bench_micro/tc_probe.cu).
//
// Mapping (measured design points in profiles/r01_tc_probe_*.log):
//   * D[neuron m][trajectory n] = W[m][:] . act[n][:]: the WEIGHTS are the A operand and stay resident in TMEM for the whole
//     kernel (hi and lo copies of layers 2, 3 and 4: 384 of the 512 columns, the other 128 are the accumulators); only the
//     16-row activation tile (B operand, 512 B per MMA) is fetched from shared memory.  With A in shared memory the MMA rate was
//     bound by the 2 KB operand fetch (27-50 cycles per MMA); from TMEM the issue overhead (~60 cycles per MMA and warp)
//     dominates, so the 24 MMAs of a layer are issued by all 4 warps of the group in parallel, each into its own
//     accumulator (6 MMAs per warp, K-split), and the 4 partial accumulators are summed in the epilogue.
//   * A group = 128 threads = 16 trajectories.  Warp w reads accumulator rows 16w..16w+15 with tcgen05.ld.16x256b (thread t:
//     rows 16w + t/4 (+8), trajectories 2(t%4) (+1) (+8)), adds the bias, applies ELU, splits hi/lo and writes the next
//     layer's B tile (K-major, no swizzle, 144-byte K-chunk stride -> conflict-free transposed stores).
//   * Two groups per CTA (one CTA per SM) run staggered, so one group's epilogue overlaps the other's MMA round trip.
//   * Layer 1 is folded as in the CUDA-core kernel: W1 [a0; s-a0; s] + b1 = (Wb+Wc) [x; u] + c1, c1 = (Wa-Wb) a0 + b1 is a
//     per-(neuron, trajectory) register constant; the B tile of layer 1 is [x (16) | held inputs (8)].
//     The folded layer 1 (K = 24, 9 MMAs) keeps its weights in shared memory.
//   * Layer 4 (64 -> 16): W4 is replicated into all four 16-row blocks of the M = 64 operand, so every warp receives the
//     whole 16 x 16 slope tile and processes one quarter of it (2 elements per thread: rows t/4 and t/4+8 of one trajectory
//     column) -- x, k1..k3 live in registers, the stage algebra follows the reference's operation order, the next stage's
//     x columns are written straight into the layer-1 B tile; trajectory rows leave as 1 KB contiguous 128-bit stores.
//   * DAE: the algebraic net (S+X+Z+V -> 64 -> 64 -> 64 -> I) is evaluated once per step on the SAME layer-1 B tile
//     [x | z v i] (folded the same way: its a0 part is a per-trajectory constant, its i columns carry zero weights); TMEM is
//     full with the DE weights, so the AE weights are shared-memory A operands.  i_j is written straight into the i columns
//     of the B tile (the held input of the next step).  The tile holds the UN-jumped z[j], v[j] while i_j is evaluated
//     (my_solvers.py:121); on an event step the jumped values replace them and i_0 is re-evaluated first (:108-110).
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
constexpr int SBO_W64 = (TH / 4) * LBO_W, W64_TILE = (TH / 8) * SBO_W64;   // 64 x 64 weight tiles in shared memory (DAE: AE layers 2..4)
// TMEM columns: accumulators first (2 groups x 4 warps x 16), then the resident weights
constexpr int TM_ACC = 0;
constexpr int TM_W2 = 128, TM_W3 = 256, TM_W4 = 384;     // hi at +0, lo at +64
constexpr int TM_COLS = 512;
constexpr int GROUP_THREADS = 128;

struct TcParams {
    int B, T, Z, S, groups;
    int V, I;                                  // DAE only (0 for an ODE); U = Z + V + I <= 8 held-input columns
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
    float* tape;            // activation tape for the tensor-core reverse sweep (psnode_tc_tape.cuh) or nullptr
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
struct __align__(128) CtaSmemDae {            // DAE: the AE net's weights follow (A operands read from shared memory)
    CtaSmem base;
    float wa1_hi[W1_TILE / 4], wa1_lo[W1_TILE / 4];
    unsigned char wa2_hi[W64_TILE], wa2_lo[W64_TILE];
    unsigned char wa3_hi[W64_TILE], wa3_lo[W64_TILE];
    unsigned char wa4_hi[W64_TILE], wa4_lo[W64_TILE];
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }

template <int METHOD, bool DAE>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc_ode_kernel(const __grid_constant__ TcParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmemDae& smd = *reinterpret_cast<CtaSmemDae*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));   // AE part only if DAE
    CtaSmem& sm = smd.base;
    const int tid = threadIdx.x;
    const int g = tid >> 7;                    // group
    const int gt = tid & 127;                  // thread within the group
    const int warp = gt >> 5, lane = gt & 31;  // warp within the group == TMEM sub-partition (CTA warp index % 4)
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, Z = q.Z, S = q.S;
    const int ZV = Z + (DAE ? q.V : 0);                // columns of the B tile fed from the z / v series
    const int U = ZV + (DAE ? q.I : 0);                // held-input columns (i columns are written by the AE epilogue)
    const int b0 = (blockIdx.x * q.groups + g) * TN;
    const bool live = g < q.groups && b0 < B;          // whole group has at least one trajectory

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&sm.g[0].bar, 4); mbar_init(&sm.g[1].bar, 4); fence_mbar_init(); }
    if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, TM_COLS);
    // folded layer-1 weights (Wb + Wc restricted to [x | held inputs], zero padded to K = 24) -> shared-memory tiles
    {
        const int K1 = 3 * S;
        for (int e = tid; e < TH * TK1; e += 2 * GROUP_THREADS) {
            const int m = e / TK1, c = e - m * TK1;
            const int k = c < TX ? c : (c - TX < U ? c : -1);
            float hi = 0.0f, lo = 0.0f;
            if (k >= 0) split_tf32(__ldg(q.W1 + m * K1 + S + k) + __ldg(q.W1 + m * K1 + 2 * S + k), hi, lo);
            sm.w1_hi[tile_byte(m, c, LBO_W, SBO_W) >> 2] = hi;
            sm.w1_lo[tile_byte(m, c, LBO_W, SBO_W) >> 2] = lo;
        }
    }
    if constexpr (DAE) {
        // AE layer 1 folded onto the same B tile: columns [x | z v] carry W[:, S + c], the i columns (and the padding) zero
        const int KA = S + TX + ZV;
        for (int e = tid; e < TH * TK1; e += 2 * GROUP_THREADS) {
            const int m = e / TK1, c = e - m * TK1;
            float hi = 0.0f, lo = 0.0f;
            if (c < TX + ZV) split_tf32(__ldg(q.A1 + m * KA + S + c), hi, lo);
            smd.wa1_hi[tile_byte(m, c, LBO_W, SBO_W) >> 2] = hi;
            smd.wa1_lo[tile_byte(m, c, LBO_W, SBO_W) >> 2] = lo;
        }
        for (int e = tid; e < TH * TH; e += 2 * GROUP_THREADS) {
            const int m = e >> 6, k = e & 63;
            const int o = tile_byte(m, k, LBO_W, SBO_W64);
            float hi, lo;
            split_tf32(__ldg(q.A2 + e), hi, lo);
            *reinterpret_cast<float*>(smd.wa2_hi + o) = hi; *reinterpret_cast<float*>(smd.wa2_lo + o) = lo;
            split_tf32(__ldg(q.A3 + e), hi, lo);
            *reinterpret_cast<float*>(smd.wa3_hi + o) = hi; *reinterpret_cast<float*>(smd.wa3_lo + o) = lo;
            hi = 0.0f; lo = 0.0f;                              // layer 4 (I rows) replicated into every 16-row block
            if ((m & 15) < q.I) split_tf32(__ldg(q.A4 + (m & 15) * TH + k), hi, lo);
            *reinterpret_cast<float*>(smd.wa4_hi + o) = hi; *reinterpret_cast<float*>(smd.wa4_lo + o) = lo;
        }
    }
    for (int e = gt; e < (int)(offsetof(GroupSmem, bar) / 4); e += GROUP_THREADS) reinterpret_cast<float*>(&gs)[e] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    // this thread's accumulator fragment: element i <-> (row m0 + 8*((i>>1)&1), trajectory c0 + (i&1) + 8*(i>>2))
    const int m0 = 16 * warp + (lane >> 2), c0 = 2 * (lane & 3);
    auto frag_row = [&](int i) { return m0 + ((i >> 1) & 1) * 8; };
    auto frag_col = [&](int i) { return c0 + (i & 1) + (i >> 2) * 8; };

    // resident weights -> TMEM (group 0 writes; both groups read them through the tensor core only).  W4 (16 rows) is
    // replicated into every 16-row block: row r of the operand holds W4[r & 15].
    if (g == 0) {
        for (int half = 0; half < 2; half++) {
            for (int cb = 0; cb < 4; cb++) {               // 64 columns = 4 x 16
                float w2[8], w3[8], w4[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = frag_row(i), col = 16 * cb + frag_col(i);
                    float hi, lo;
                    split_tf32(__ldg(q.W2 + row * TH + col), hi, lo); w2[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W3 + row * TH + col), hi, lo); w3[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W4 + (row & 15) * TH + col), hi, lo); w4[i] = half ? lo : hi;
                }
                tmem_st_16x256b_x2(tmem + lane_base + TM_W2 + 64 * half + 16 * cb, w2);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W3 + 64 * half + 16 * cb, w3);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W4 + 64 * half + 16 * cb, w4);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();            // resident weights visible to both groups' MMAs
    tc_fence_after();
    // per-thread constants: biases of its two rows, c1 of its 8 (row, trajectory) elements
    float bias2[2], bias3[2], bias4[2], c1[8];   // bias2 doubles as a dummy for layer 1 (c1 carries b1)
#pragma unroll
    for (int r = 0; r < 2; r++) {
        bias2[r] = __ldg(q.b2 + m0 + 8 * r);
        bias3[r] = __ldg(q.b3 + m0 + 8 * r);
        bias4[r] = __ldg(q.b4 + ((m0 + 8 * r) & 15));
    }
    {
        const int K1 = 3 * S;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int row = frag_row(i), bb = min(b0 + frag_col(i), B - 1);
            float acc = __ldg(q.b1 + row);
            if (live)
                for (int k = 0; k < S; k++)
                    acc = fmaf(__ldg(q.W1 + row * K1 + k) - __ldg(q.W1 + row * K1 + S + k), __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), acc);
            c1[i] = acc;
        }
    }
    float biasA2[2] = {0.f, 0.f}, biasA3[2] = {0.f, 0.f}, biasA4[2] = {0.f, 0.f}, c1a[8];   // AE net (DAE)
    if constexpr (DAE) {
        const int KA = S + TX + ZV;
#pragma unroll
        for (int r = 0; r < 2; r++) {
            biasA2[r] = __ldg(q.ab2 + m0 + 8 * r);
            biasA3[r] = __ldg(q.ab3 + m0 + 8 * r);
            const int ci = (m0 + 8 * r) & 15;
            biasA4[r] = ci < q.I ? __ldg(q.ab4 + ci) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {       // AE layer 1: the all_initial block is a per-(neuron, trajectory) constant
            const int row = frag_row(i), bb = min(b0 + frag_col(i), B - 1);
            float acc = __ldg(q.ab1 + row);
            if (live)
                for (int k = 0; k < S; k++) acc = fmaf(__ldg(q.A1 + row * KA + k), __ldg(q.a0 + (int64_t)bb * q.a0_sb + k), acc);
            c1a[i] = acc;
        }
    }
    // activation-tile byte offsets of this thread's 8 elements (row = K index of the next layer, column = trajectory)
    int off_act[8];
#pragma unroll
    for (int i = 0; i < 8; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    // the two state elements this thread owns in the layer-4 epilogue: states sm0, sm0 + 8 of trajectory column sn
    const int sm0 = lane >> 2, sn = c0 + (warp & 1) + 8 * (warp >> 1);
    const int off_x[2] = {(int)tile_byte(sn, sm0, LBO, SBO_B1), (int)tile_byte(sn, sm0 + 8, LBO, SBO_B1)};
    // DAE: the algebraic outputs (rows sm0, sm0 + 8 of the replicated layer-4 tile) this thread writes into the i columns
    const int off_i[2] = {(int)tile_byte(sn, min(TX + ZV + sm0, TK1 - 1), LBO, SBO_B1), (int)tile_byte(sn, min(TX + ZV + sm0 + 8, TK1 - 1), LBO, SBO_B1)};
    // descriptors
    const uint32_t idesc = make_idesc_tf32(TH, TN);
    const uint64_t d_act_hi = make_desc(smem_u32(gs.act_hi), LBO, SBO_ACT), d_act_lo = make_desc(smem_u32(gs.act_lo), LBO, SBO_ACT);
    const uint64_t d_b1_hi = make_desc(smem_u32(gs.b1_hi), LBO, SBO_B1), d_b1_lo = make_desc(smem_u32(gs.b1_lo), LBO, SBO_B1);
    const uint64_t d_w1_hi = make_desc(smem_u32(sm.w1_hi), LBO_W, SBO_W), d_w1_lo = make_desc(smem_u32(sm.w1_lo), LBO_W, SBO_W);
    const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4) * TN;      // 4 partial accumulators of this group
    const uint32_t my_acc = acc_base + (uint32_t)warp * TN;                // the one this warp's MMAs write
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
    uint32_t phase = 0;

    // ---- helpers ---------------------------------------------------------------------------------------
    // TS layers (weights in TMEM): this warp's K-steps [ks0, ks0 + nks) for the three 3xTF32 terms, small terms first
    auto issue_ts = [&](uint32_t w_hi, uint32_t w_lo, uint64_t b_hi, uint64_t b_lo, int ks0, int nks) {
        if (elect_one()) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int term = 0; term < 3; term++) {
                const uint32_t wa = term == 0 ? w_lo : w_hi;
                const uint64_t bd = term == 1 ? b_lo : b_hi;
                for (int kk = 0; kk < nks; kk++) {
                    const int ks = ks0 + kk;
                    mma_tf32_ts(my_acc, tmem + wa + 8 * ks, bd + KSTEP_B * ks, idesc, accumulate);
                    accumulate = 1;
                }
            }
            mma_commit(&gs.bar);
        }
        __syncwarp();
    };
    auto issue_ss = [&](uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, int ks0, int nks) {
        if (elect_one()) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int term = 0; term < 3; term++) {
                const uint64_t ad = term == 0 ? a_lo : a_hi;
                const uint64_t bd = term == 1 ? b_lo : b_hi;
                for (int kk = 0; kk < nks; kk++) {
                    const int ks = ks0 + kk;
                    mma_tf32(my_acc, ad + KSTEP_W * ks, bd + KSTEP_B * ks, idesc, accumulate);
                    accumulate = 1;
                }
            }
            mma_commit(&gs.bar);
        }
        __syncwarp();
    };
    // wait for the group's 4 commits
    auto wait_mma = [&]() {
        if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
        phase ^= 1;
        tc_fence_after();
    };
    // sum the first `nacc` partial accumulators into d[8] (this thread's 2 rows x 4 trajectory columns)
    auto collect = [&](float (&d)[8], int nacc) {
        wait_mma();
        float t0[8], t1[8], t2[8], t3[8];
        tmem_ld_16x256b_x2(acc_base + lane_base + 0 * TN, t0);
        tmem_ld_16x256b_x2(acc_base + lane_base + 1 * TN, t1);
        tmem_ld_16x256b_x2(acc_base + lane_base + 2 * TN, t2);
        if (nacc == 4) tmem_ld_16x256b_x2(acc_base + lane_base + 3 * TN, t3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = nacc == 4 ? (t0[i] + t1[i]) + (t2[i] + t3[i]) : (t0[i] + t1[i]) + t2[i];
    };
    // layer 4: the slope elements (state sm0 / sm0 + 8, trajectory sn) of this thread; every 16-row block holds the same tile
    auto collect_slopes = [&](float (&kv)[2]) {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc_base + lane_base + 8 * (warp >> 1);
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
        const bool o = (warp & 1) != 0;
        kv[0] = ((o ? t0[1] : t0[0]) + (o ? t1[1] : t1[0])) + ((o ? t2[1] : t2[0]) + (o ? t3[1] : t3[0]));
        kv[1] = ((o ? t0[3] : t0[2]) + (o ? t1[3] : t1[2])) + ((o ? t2[3] : t2[2]) + (o ? t3[3] : t3[2]));
    };
    // publish freshly written B-tile data to the tensor core and line the group up for the next layer's MMAs
    auto publish = [&]() {
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
    };
    auto store_hidden = [&](const float (&d)[8], const float (&bias)[2], const float* cadd, float* trec) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float v = d[i] + bias[(i >> 1) & 1];
            if (cadd) v = d[i] + cadd[i];
            a[i] = psn_elu(v);
            float hi, lo;
            split_tf32_fast(a[i], hi, lo);
            *reinterpret_cast<float*>(gs.act_hi + off_act[i]) = hi;
            *reinterpret_cast<float*>(gs.act_lo + off_act[i]) = lo;
        }
        if (trec) {   // thread-private, coalesced: 32 contiguous bytes per thread
            __stcs(reinterpret_cast<float4*>(trec + gt * 8), make_float4(a[0], a[1], a[2], a[3]));
            __stcs(reinterpret_cast<float4*>(trec + gt * 8) + 1, make_float4(a[4], a[5], a[6], a[7]));
        }
    };
    // z / v columns of the layer-1 B tile (warp 1, lane = trajectory): values of grid point jp, or of event k when k >= 0
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
    auto load_dt = [&](int j) {      // step that ENDS at grid point j
        const int bb = min(b0 + (lane & 15), B - 1);
        return __fsub_rn(ldser(q.t, j, bb, 0), ldser(q.t, j - 1, bb, 0));
    };
    auto store_zv = [&](const float (&u)[TU]) {
        if (lane < TN) {
#pragma unroll
            for (int c = 0; c < TU; c++) {
                if (c < ZV) {
                    float hi, lo;
                    split_tf32_fast(u[c], hi, lo);
                    const int o = tile_byte(lane, TX + c, LBO, SBO_B1);
                    *reinterpret_cast<float*>(gs.b1_hi + o) = hi;
                    *reinterpret_cast<float*>(gs.b1_lo + o) = lo;
                }
            }
        }
    };
    auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };   // step that ENDS at j
    // A operand from shared memory, K = 64 (AE layers 2..4): this warp's two K-steps of the three 3xTF32 terms
    auto issue_ss64 = [&](uint64_t a_hi, uint64_t a_lo) {
        if (elect_one()) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int term = 0; term < 3; term++) {
                const uint64_t ad = term == 0 ? a_lo : a_hi;
                const uint64_t bd = term == 1 ? d_act_lo : d_act_hi;
                for (int kk = 0; kk < 2; kk++) {
                    const int ks = 2 * warp + kk;
                    mma_tf32(my_acc, ad + KSTEP_W * ks, bd + KSTEP_B * ks, idesc, accumulate);
                    accumulate = 1;
                }
            }
            mma_commit(&gs.bar);
        }
        __syncwarp();
    };
    // DAE: i = ae(x, z, v) on the current layer-1 B tile (AE_Func.forward, neural_01_DAE_01_no_encode.py:74-83); the result goes
    // into the tile's i columns (held input of the next DE evaluations) and, with `stage_out`, into istage (-> i_sol row).
    auto ae_eval = [&](bool stage_out) {
        if constexpr (DAE) {
            const uint64_t d_wa1_hi = make_desc(smem_u32(smd.wa1_hi), LBO_W, SBO_W), d_wa1_lo = make_desc(smem_u32(smd.wa1_lo), LBO_W, SBO_W);
            const uint64_t d_wa2_hi = make_desc(smem_u32(smd.wa2_hi), LBO_W, SBO_W64), d_wa2_lo = make_desc(smem_u32(smd.wa2_lo), LBO_W, SBO_W64);
            const uint64_t d_wa3_hi = make_desc(smem_u32(smd.wa3_hi), LBO_W, SBO_W64), d_wa3_lo = make_desc(smem_u32(smd.wa3_lo), LBO_W, SBO_W64);
            const uint64_t d_wa4_hi = make_desc(smem_u32(smd.wa4_hi), LBO_W, SBO_W64), d_wa4_lo = make_desc(smem_u32(smd.wa4_lo), LBO_W, SBO_W64);
            float d[8];
            issue_ss(d_wa1_hi, d_wa1_lo, d_b1_hi, d_b1_lo, warp, warp < 3 ? 1 : 0);
            collect(d, 3);
            store_hidden(d, biasA2, c1a, nullptr);
            publish();
            issue_ss64(d_wa2_hi, d_wa2_lo);
            collect(d, 4);
            store_hidden(d, biasA2, nullptr, nullptr);
            publish();
            issue_ss64(d_wa3_hi, d_wa3_lo);
            collect(d, 4);
            store_hidden(d, biasA3, nullptr, nullptr);
            publish();
            issue_ss64(d_wa4_hi, d_wa4_lo);
            float kv[2];
            collect_slopes(kv);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int ci = sm0 + 8 * r;
                if (ci < q.I) {
                    const float iv = kv[r] + biasA4[r];
                    float hi, lo;
                    split_tf32_fast(iv, hi, lo);
                    *reinterpret_cast<float*>(gs.b1_hi + off_i[r]) = hi;
                    *reinterpret_cast<float*>(gs.b1_lo + off_i[r]) = lo;
                    if (stage_out) gs.istage[sn][ci] = iv;
                }
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

    if (live) {
        // ---- initial state: every thread owns two state elements (states sm0, sm0 + 8 of trajectory column sn) -------
        float x0[2], k1[2] = {0.f, 0.f}, k2[2] = {0.f, 0.f}, k3[2] = {0.f, 0.f};
        {
            const int b = b0 + sn, bb = min(b, B - 1);
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const float xv = DAE ? __ldg(q.x_init + (int64_t)bb * q.x_init_sb + sm0 + 8 * r) : ldser(q.x, 0, bb, sm0 + 8 * r);
                x0[r] = xv;
                if (b < B) q.x_sol.p[(int64_t)b * q.x_sol.sb + sm0 + 8 * r] = xv;
                float hi, lo;
                split_tf32_fast(xv, hi, lo);
                *reinterpret_cast<float*>(gs.b1_hi + off_x[r]) = hi;
                *reinterpret_cast<float*>(gs.b1_lo + off_x[r]) = lo;
            }
        }
        if (warp == 1) {
            float u[TU];
            if (DAE) load_zv(0, -1, u);                        // i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95)
            else if (T > 1) load_zv(0, event_of_step(1), u);
            if (DAE || T > 1) store_zv(u);
            if (T > 1 && lane < TN) gs.dts[1][lane] = load_dt(1);
        }
        publish();
        if constexpr (DAE) {
            ae_eval(true);
            flush_i(0);
        }

        const float c13 = (float)(1.0 / 3.0);
        float* trec = q.tape ? q.tape + (int64_t)(blockIdx.x * q.groups + g) * (T - 1) * NST * PSN_TAPE_STAGE : nullptr;
        float ycur[2] = {x0[0], x0[1]};                     // input of the current stage (recorded on the tape)
        for (int j = 1; j < T; j++) {
            float un[TU] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dtn = 0.0f;   // next step's inputs, prefetched by warp 1 during stage 0
            const bool have_next = j + 1 < T;
            const float dt = gs.dts[j & 1][sn];
            if constexpr (DAE) {
                const int k = event_of_step(j);
                if (k >= 0) {                                  // event: jumped z / v replace the held inputs and i_0 is re-evaluated
                    if (warp == 1) { float uj[TU]; load_zv(j - 1, k, uj); store_zv(uj); }
                    publish();
                    ae_eval(false);
                }
            }
#pragma unroll 1
            for (int e = 0; e < NST; e++) {
                float d[8];
                if (trec) __stcs(reinterpret_cast<float2*>(trec + 3 * PSN_TAPE_FRAG + gt * 2), make_float2(ycur[0], ycur[1]));
                // ---- layer 1 (shared-memory weights): K = 24 -> warps 0..2 take one K-step each; warp 3 only commits ----
                issue_ss(d_w1_hi, d_w1_lo, d_b1_hi, d_b1_lo, warp, warp < 3 ? 1 : 0);
                if (e == 0 && warp == 1) {
                    // next step's held inputs; DAE: the un-jumped z[j], v[j] (they feed i_j first), ODE: jumped if step j+1 fires
                    if (DAE) load_zv(j, -1, un);
                    else if (have_next) load_zv(j, event_of_step(j + 1), un);
                    if (have_next) dtn = load_dt(j + 1);
                }
                if (DAE && e == 0 && j > 1) flush_i(j - 1);
                if (e == 0 && j > 1 && gt < 64) {           // trajectory row j-1 (staged by the previous step's last stage)
                    const int n = gt >> 2, c4 = gt & 3, b = b0 + n;
                    if (b < B) {
                        float* dst = q.x_sol.p + (int64_t)(j - 1) * q.x_sol.st + (int64_t)b * q.x_sol.sb + 4 * c4;
                        const float4 v = *reinterpret_cast<const float4*>(&gs.ostage[n][4 * c4]);
                        if (q.vec_out) *reinterpret_cast<float4*>(dst) = v;
                        else { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
                    }
                }
                collect(d, 3);
                store_hidden(d, bias2, c1, trec);
                publish();
                // ---- layer 2 ----
                issue_ts(TM_W2, TM_W2 + 64, d_act_hi, d_act_lo, 2 * warp, 2);
                collect(d, 4);
                store_hidden(d, bias2, nullptr, trec ? trec + PSN_TAPE_FRAG : nullptr);
                publish();
                // ---- layer 3 ----
                issue_ts(TM_W3, TM_W3 + 64, d_act_hi, d_act_lo, 2 * warp, 2);
                collect(d, 4);
                store_hidden(d, bias3, nullptr, trec ? trec + 2 * PSN_TAPE_FRAG : nullptr);
                publish();
                // ---- layer 4 + stage algebra: 2 state elements per thread ----
                issue_ts(TM_W4, TM_W4 + 64, d_act_hi, d_act_lo, 2 * warp, 2);
                float kv[2];
                collect_slopes(kv);
                const bool last = e == NST - 1;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const float kk = kv[r] + bias4[r];
                    float xn;
                    if (METHOD == PSNODE_EULER) {
                        xn = __fadd_rn(x0[r], __fmul_rn(dt, kk));
                    } else if (METHOD == PSNODE_MIDPOINT) {
                        if (e == 0) xn = __fadd_rn(x0[r], __fmul_rn(kk, __fmul_rn(0.5f, dt)));
                        else xn = __fadd_rn(x0[r], __fmul_rn(dt, kk));
                    } else {
                        if (e == 0) { k1[r] = kk; xn = __fadd_rn(x0[r], __fmul_rn(__fmul_rn(dt, kk), c13)); }
                        else if (e == 1) { k2[r] = kk; xn = __fadd_rn(x0[r], __fmul_rn(dt, __fsub_rn(kk, __fmul_rn(k1[r], c13)))); }
                        else if (e == 2) { k3[r] = kk; xn = __fadd_rn(x0[r], __fmul_rn(dt, __fadd_rn(__fsub_rn(k1[r], k2[r]), kk))); }
                        else {
                            const float ksum = __fadd_rn(__fadd_rn(k1[r], __fmul_rn(3.0f, __fadd_rn(k2[r], k3[r]))), kk);
                            xn = __fadd_rn(x0[r], __fmul_rn(__fmul_rn(ksum, dt), 0.125f));
                        }
                    }
                    float hi, lo;
                    split_tf32_fast(xn, hi, lo);
                    *reinterpret_cast<float*>(gs.b1_hi + off_x[r]) = hi;
                    *reinterpret_cast<float*>(gs.b1_lo + off_x[r]) = lo;
                    ycur[r] = xn;
                    if (last) { x0[r] = xn; gs.ostage[sn][sm0 + 8 * r] = xn; }
                }
                if (trec) trec += PSN_TAPE_STAGE;
                if (warp == 1 && last && (DAE || have_next)) {      // all layer-1 MMAs of this step are done
                    store_zv(un);
                    if (have_next && lane < TN) gs.dts[(j + 1) & 1][lane] = dtn;
                }
                publish();
            }
            if constexpr (DAE) ae_eval(true);                   // i_j = ae(x_j, z[j], v[j])  (my_solvers.py:121)
        }
        if (DAE && T > 1) flush_i(T - 1);
        if (T > 1 && gt < 64) {                             // last trajectory row
            const int n = gt >> 2, c4 = gt & 3, b = b0 + n;
            if (b < B) {
                float* dst = q.x_sol.p + (int64_t)(T - 1) * q.x_sol.st + (int64_t)b * q.x_sol.sb + 4 * c4;
                const float4 v = *reinterpret_cast<const float4*>(&gs.ostage[n][4 * c4]);
                if (q.vec_out) *reinterpret_cast<float4*>(dst) = v;
                else { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
            }
        }
    }
    // ---- teardown --------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tmem, TM_COLS);
}

}  // namespace

static bool h64_net(const psnode_mlp& m, int in0, int out_last) {
    return m.n_layers == 4 && m.in_dim[0] == in0 && m.out_dim[0] == TH && m.out_dim[1] == TH && m.out_dim[2] == TH &&
           m.out_dim[3] == out_last;
}

// shapes of the tensor-core kernels: 4-layer H = 64 nets, up to 16 state variables, up to 8 held-input columns
bool psn_tc_supports(const psnode_problem* p) {
    if (p->teacher_x || p->teacher_i || p->X < 1 || p->X > TX) return false;
    const int S = p->X + p->Z + p->V + p->I;
    if (S - p->X > TU) return false;
    if (!h64_net(p->de, 3 * S, p->X)) return false;
    if (p->kind == PSNODE_DAE) return p->I >= 1 && h64_net(p->ae, S + p->X + p->Z + p->V, p->I);
    return p->kind == PSNODE_ODE;
}

int64_t psn_tc_forward_workspace(const psnode_problem*) { return 256; }

int psn_tc_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < 4) return PSNODE_EWORKSPACE;
    const bool dae = p->kind == PSNODE_DAE;
    TcParams q;
    q.B = p->B; q.T = p->T; q.Z = p->Z; q.S = p->X + p->Z + p->V + p->I;
    q.V = p->V; q.I = p->I;
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
    q.vec_out = ((reinterpret_cast<uintptr_t>(p->x_sol.p) & 15) == 0 && (p->x_sol.st & 3) == 0 && (p->x_sol.sb & 3) == 0) ? 1 : 0;
    q.tape = (!dae && p->tape && p->tape_floats >= psn_tc_tape_floats(p->B, p->T, p->method)) ? p->tape : nullptr;
    q.err = static_cast<int*>(ws);
    PSN_CUDA(cudaMemsetAsync(q.err, 0, 4, stream));
    // two groups of 16 trajectories per CTA once there are enough trajectories to give every SM a CTA
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
    if (dae) {
        switch (p->method) {
            case PSNODE_EULER: return launch(psn_tc_ode_kernel<PSNODE_EULER, true>, "psn_tc_dae_kernel<euler>");
            case PSNODE_MIDPOINT: return launch(psn_tc_ode_kernel<PSNODE_MIDPOINT, true>, "psn_tc_dae_kernel<midpoint>");
            default: return launch(psn_tc_ode_kernel<PSNODE_RK4, true>, "psn_tc_dae_kernel<rk4>");
        }
    }
    switch (p->method) {
        case PSNODE_EULER: return launch(psn_tc_ode_kernel<PSNODE_EULER, false>, "psn_tc_ode_kernel<euler>");
        case PSNODE_MIDPOINT: return launch(psn_tc_ode_kernel<PSNODE_MIDPOINT, false>, "psn_tc_ode_kernel<midpoint>");
        default: return launch(psn_tc_ode_kernel<PSNODE_RK4, false>, "psn_tc_ode_kernel<rk4>");
    }
}
