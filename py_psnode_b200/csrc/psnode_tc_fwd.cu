// psnode_tc_fwd.cu -- tensor-core forward integrator for the reference's H = 64 ODE nets (BASELINE configs[1]):
// integrate_ODE (neural_dae/my_solvers.py:52-80) with the 4-layer DE_Func of neural_00_ODE_01_no_encode.py:58-68
// (3S -> 64 -> 64 -> 64 -> 16, ELU), Euler / Midpoint / RK4-3/8 (neural_dae/my_fixed_grid.py:15-59), event jumps.
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
// TMEM columns: accumulators first (2 groups x 4 warps x 16), then the resident weights
constexpr int TM_ACC = 0;
constexpr int TM_W2 = 128, TM_W3 = 256, TM_W4 = 384;     // hi at +0, lo at +64
constexpr int TM_COLS = 512;
constexpr int GROUP_THREADS = 128;

struct TcParams {
    int B, T, Z, S, groups;
    psnode_series t, x, z;
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
    float dts[2][TN];
    uint64_t bar;
};

struct __align__(128) CtaSmem {
    float w1_hi[W1_TILE / 4];
    float w1_lo[W1_TILE / 4];
    GroupSmem g[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }

template <int METHOD>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc_ode_kernel(const __grid_constant__ TcParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    CtaSmem& sm = *reinterpret_cast<CtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    const int g = tid >> 7;                    // group
    const int gt = tid & 127;                  // thread within the group
    const int warp = gt >> 5, lane = gt & 31;  // warp within the group == TMEM sub-partition (CTA warp index % 4)
    GroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, Z = q.Z, S = q.S;
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
            const int k = c < TX ? c : (c - TX < Z ? c : -1);
            float hi = 0.0f, lo = 0.0f;
            if (k >= 0) split_tf32(__ldg(q.W1 + m * K1 + S + k) + __ldg(q.W1 + m * K1 + 2 * S + k), hi, lo);
            sm.w1_hi[tile_byte(m, c, LBO_W, SBO_W) >> 2] = hi;
            sm.w1_lo[tile_byte(m, c, LBO_W, SBO_W) >> 2] = lo;
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
    // activation-tile byte offsets of this thread's 8 elements (row = K index of the next layer, column = trajectory)
    int off_act[8];
#pragma unroll
    for (int i = 0; i < 8; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    // the two state elements this thread owns in the layer-4 epilogue: states sm0, sm0 + 8 of trajectory column sn
    const int sm0 = lane >> 2, sn = c0 + (warp & 1) + 8 * (warp >> 1);
    const int off_x[2] = {(int)tile_byte(sn, sm0, LBO, SBO_B1), (int)tile_byte(sn, sm0 + 8, LBO, SBO_B1)};
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
    // held inputs / dt of the step that ENDS at grid point j -> B1 tile columns 16.., dts[j & 1]; done by warp 1, lane = trajectory
    auto load_step_inputs = [&](int j, float (&u)[TU], float& dt) {
        const int bb = min(b0 + (lane & 15), B - 1);
        dt = __fsub_rn(ldser(q.t, j, bb, 0), ldser(q.t, j - 1, bb, 0));
        const int k = q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1;
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, j - 1, bb, c);
        }
    };
    auto store_step_inputs = [&](int j, const float (&u)[TU], float dt) {
        if (lane < TN) {
            gs.dts[j & 1][lane] = dt;
#pragma unroll
            for (int c = 0; c < TU; c++) {
                float hi, lo;
                split_tf32_fast(u[c], hi, lo);
                const int o = tile_byte(lane, TX + c, LBO, SBO_B1);
                *reinterpret_cast<float*>(gs.b1_hi + o) = hi;
                *reinterpret_cast<float*>(gs.b1_lo + o) = lo;
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
                const float xv = ldser(q.x, 0, bb, sm0 + 8 * r);
                x0[r] = xv;
                if (b < B) q.x_sol.p[(int64_t)b * q.x_sol.sb + sm0 + 8 * r] = xv;
                float hi, lo;
                split_tf32_fast(xv, hi, lo);
                *reinterpret_cast<float*>(gs.b1_hi + off_x[r]) = hi;
                *reinterpret_cast<float*>(gs.b1_lo + off_x[r]) = lo;
            }
        }
        if (warp == 1 && T > 1) {
            float u[TU], dt;
            load_step_inputs(1, u, dt);
            store_step_inputs(1, u, dt);
        }
        publish();

        const float c13 = (float)(1.0 / 3.0);
        float* trec = q.tape ? q.tape + (int64_t)(blockIdx.x * q.groups + g) * (T - 1) * NST * PSN_TAPE_STAGE : nullptr;
        float ycur[2] = {x0[0], x0[1]};                     // input of the current stage (recorded on the tape)
        for (int j = 1; j < T; j++) {
            float un[TU], dtn = 0.0f;                       // next step's inputs, prefetched by warp 1 during stage 0
            const bool have_next = j + 1 < T;
            const float dt = gs.dts[j & 1][sn];
#pragma unroll 1
            for (int e = 0; e < NST; e++) {
                float d[8];
                if (trec) __stcs(reinterpret_cast<float2*>(trec + 3 * PSN_TAPE_FRAG + gt * 2), make_float2(ycur[0], ycur[1]));
                // ---- layer 1 (shared-memory weights): K = 24 -> warps 0..2 take one K-step each; warp 3 only commits ----
                issue_ss(d_w1_hi, d_w1_lo, d_b1_hi, d_b1_lo, warp, warp < 3 ? 1 : 0);
                if (e == 0 && warp == 1 && have_next) load_step_inputs(j + 1, un, dtn);
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
                if (warp == 1 && last && have_next) store_step_inputs(j + 1, un, dtn);   // all layer-1 MMAs of this step are done
                publish();
            }
        }
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

bool psn_tc_supports(const psnode_problem* p) {
    if (p->kind != PSNODE_ODE || p->teacher_x) return false;
    if (p->X != TX || p->Z < 0 || p->Z > TU) return false;
    const psnode_mlp& m = p->de;
    if (m.n_layers != 4) return false;
    const int S = p->X + p->Z;
    return m.in_dim[0] == 3 * S && m.out_dim[0] == TH && m.out_dim[1] == TH && m.out_dim[2] == TH && m.out_dim[3] == TX;
}

int64_t psn_tc_forward_workspace(const psnode_problem*) { return 256; }

int psn_tc_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < 4) return PSNODE_EWORKSPACE;
    TcParams q;
    q.B = p->B; q.T = p->T; q.Z = p->Z; q.S = p->X + p->Z;
    q.t = p->t; q.x = p->x; q.z = p->z;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.x_sol = p->x_sol;
    q.W1 = p->de.W[0]; q.b1 = p->de.b[0]; q.W2 = p->de.W[1]; q.b2 = p->de.b[1];
    q.W3 = p->de.W[2]; q.b3 = p->de.b[2]; q.W4 = p->de.W[3]; q.b4 = p->de.b[3];
    q.vec_out = ((reinterpret_cast<uintptr_t>(p->x_sol.p) & 15) == 0 && (p->x_sol.st & 3) == 0 && (p->x_sol.sb & 3) == 0) ? 1 : 0;
    q.tape = (p->tape && p->tape_floats >= psn_tc_tape_floats(p->B, p->T, p->method)) ? p->tape : nullptr;
    q.err = static_cast<int*>(ws);
    PSN_CUDA(cudaMemsetAsync(q.err, 0, 4, stream));
    // two groups of 16 trajectories per CTA once there are enough trajectories to give every SM a CTA
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
    switch (p->method) {
        case PSNODE_EULER: return launch(psn_tc_ode_kernel<PSNODE_EULER>, "psn_tc_ode_kernel<euler>");
        case PSNODE_MIDPOINT: return launch(psn_tc_ode_kernel<PSNODE_MIDPOINT>, "psn_tc_ode_kernel<midpoint>");
        default: return launch(psn_tc_ode_kernel<PSNODE_RK4>, "psn_tc_ode_kernel<rk4>");
    }
}
