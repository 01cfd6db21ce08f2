// psnode_tc_bwd.cu -- tensor-core reverse sweep (discrete adjoint) for the reference's H = 64 ODE nets (BASELINE configs[1]).
//
// Exact reverse mode of FixedGridODESolver.integrate_ODE (neural_dae/my_solvers.py:52-80) with the Euler / Midpoint /
// RK4-3/8 step functions (neural_dae/my_fixed_grid.py:15-59) and the 4-layer DE_Func of neural_00_ODE_01_no_encode.py:58-68,
// i.e. what the reference obtains from autograd at loss.backward() (neural_00_ODE_01_no_encode.py:359).  Like autograd -- and
// unlike the generic reverse sweep (psnode_generic_bwd.cu), which recomputes every stage from x_sol -- it consumes the
// activations the forward pass recorded (psnode_tc_tape.cuh), so a stage costs four dependent GEMMs instead of eight.
//
// Per stage and 16-trajectory group (n = trajectory, all products 3xTF32 = fp32-accurate, fp32 accumulation in TMEM):
//     g3 = W4^T dk            d3 = g3 * elu'(a3)        dW4^T += a3 dk^T        db4 += dk
//     g2 = W3^T d3            d2 = g2 * elu'(a2)        dW3   += d3 a2^T        db3 += d3
//     g1 = W2^T d2            d1 = g1 * elu'(a1)        dW2   += d2 a1^T        db2 += d2
//     gy = (Wb+Wc)_x^T d1                               dW1f  += d1 [y;u]^T     D1  += d1   (per trajectory)
//   * data path (left column): the TRANSPOSED weights are the A operand (M = 64), resident in TMEM for the whole kernel like
//     the forward kernel's (W3^T, W2^T, (Wb+Wc)_x^T, hi + lo: 384 columns; W4^T, K = 16, stays in shared memory), the 16-row
//     delta tile is the B operand; 24 MMAs per layer issued by the 4 warps of the group in parallel into 4 partial
//     accumulators.  With the weights in shared memory the kernel was bound by the operand fetch (2 KB of A per MMA:
//     1300 of 2330 cycles per layer pair, profiles/r01_ncu_tc_bwd_cfg2.txt).
//   * TMEM plan: an M = 64 operand or accumulator occupies only 16 of the 32 lanes of each sub-partition.  The data path
//     (TS MMAs: A and D must sit at the same lanes) uses lanes 0..15: 128 accumulator + 384 weight columns = all 512.  The
//     weight-gradient accumulators (SS MMAs) live in lanes 16..31 of the same columns: 2 x 176.  (Wb+Wc)_x^T (16 rows) is replicated into every 16-row block so each warp receives the whole
//     dL/dy tile and handles two state elements per thread: the Runge-Kutta adjoint algebra stays in registers.
//   * weight gradients (right column): one MMA chain per layer with K = the 16 trajectories (A = delta tile [m][n],
//     B = activation tile [k][n], N = 64), accumulated in TMEM over ALL stages and steps of the launch: 176 columns per group
//     hold dW2, dW3, dW4^T and dW1f.  They are issued after the data MMAs' commit, off the critical path, and never read
//     until the kernel ends.  Bias gradients and D1 = sum of d1 per (neuron, trajectory) live in registers.
//   * layer 1 is folded in the forward pass (W1 [a0; s-a0; s] + b1 = (Wb+Wc) s + c1(a0)); its gradient is unfolded once per
//     group at the end: dWc = G, dWb = G - D1 a0^T, dWa = D1 a0^T, db1 = sum_n D1, d_a0 = (Wa-Wb)^T D1  (G = dW1f accumulator).
//   * the tensor core's fp32 accumulate TRUNCATES (measured: a chain of N accumulations into one TMEM accumulator is biased
//     towards zero by ~N * 2^-24), so the TMEM gradient accumulators are flushed into the group's slab with round-to-nearest
//     adds every 16 / NST steps (<= 16 accumulations per chain) instead of once at the end.
//   * every group writes its gradient to a private slab; psn_tc_grad_reduce_kernel sums the slabs in a fixed order
//     (deterministic, no floating-point atomics).
#include <cstddef>
#include "psnode_internal.cuh"
#include "psnode_tc.cuh"
#include "psnode_tc_tape.cuh"

namespace {
using namespace psn_tc;

constexpr int TN = PSN_TC_TN;          // trajectories per group
constexpr int TH = 64, TX = 16, TU = 8;
constexpr int TK1 = TX + TU;
constexpr int LBO = 144;               // delta tile as B operand of the data MMAs: rows = trajectory, K = neuron (as the forward act tile)
constexpr int SBO_ACT = (TH / 4) * LBO;
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;
constexpr int LBO_W = 128;             // every other tile: contiguous 8 x 16 B core matrices

constexpr int SBO_K16 = (TN / 4) * LBO_W, K16_TILE = (TH / 8) * SBO_K16;     // 64 rows x K = 16   ( 4 KB)
constexpr int DKB_TILE = (TX / 8) * SBO_K16;                                 // 16 rows x K = 16   ( 1 KB)
// TMEM columns of one group
constexpr int TM_ACC = 0;                                       // lanes 0..15: 2 groups x 4 partial accumulators x 16
constexpr int TM_W3T = 128, TM_W2T = 256, TM_W1T = 384;         // lanes 0..15: resident transposed weights, hi at +0, lo at +64
constexpr uint32_t TM_UPPER = 16u << 16;                        // lanes 16..31 of every sub-partition
constexpr int TM_DW2 = 0, TM_DW3 = 64, TM_DW4T = 128, TM_DW1F = 144, TM_DWGROUP = 176;   // per group, upper half-lanes
constexpr int TM_W4T = 352;       // lanes 16..31: W4^T (K = 16), hi at +0, lo at +16 -- with its own accumulators (TS: same lanes)
constexpr int TM_ACC_M1 = 384;    // lanes 16..31: 2 groups x 4 partial accumulators x 16 for g3 = W4^T dk
constexpr int TM_COLS = 512;
constexpr int GROUP_THREADS = 256;     // 8 warps per 16-trajectory group: warps k and k + 4 share TMEM sub-partition k
constexpr int PSN_DW_CHAIN = 16;       // accumulations into a TMEM weight-gradient accumulator between two flushes (16 / NST steps)
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }   // slabs stay 16-byte aligned when X is odd
constexpr int G_AREA = TH * TK1;       // floats of the dW1f (folded layer 1) accumulator kept behind each group's slab

struct TcBwdParams {
    int B, T, X, Z, S, groups, n_theta;        // X <= 16 state variables (rows / columns X..15 of the tiles are zero padding)
    psnode_series t, z, gx;
    PsnFuse fx;                                 // fused masked-MSE upstream gradient (fx.term.target.p != NULL: gx is ignored)
    const float* a0; int64_t a0_sb;
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    const float* W1; const float* W2; const float* W3; const float* W4;
    const float* tape;
    float* slab;
    float* d_x0; int64_t d_x0_sb;
    float* d_a0; int64_t d_a0_sb;
    int* err;
};

struct __align__(128) BwdGroupSmem {
    unsigned char dT_hi[ACT_TILE];          // delta (or dk in K columns 0..15) [n][m]: B operand of the data MMAs
    unsigned char dT_lo[ACT_TILE];
    unsigned char dA_hi[2][K16_TILE];       // delta [m][n]: A operand of the weight-gradient MMAs (double buffered)
    unsigned char dA_lo[2][K16_TILE];
    unsigned char aB_hi[2][K16_TILE];       // activation [k][n]: B operand of the weight-gradient MMAs (A operand for dW4^T)
    unsigned char aB_lo[2][K16_TILE];
    unsigned char dkB_hi[DKB_TILE];         // dk [state][n]: B operand of dW4^T
    unsigned char dkB_lo[DKB_TILE];
    uint64_t bar;
};

struct __align__(128) BwdCtaSmem {
    BwdGroupSmem g[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }
__device__ __forceinline__ void st_f32x2(unsigned char* base, int off, float a, float b) { *reinterpret_cast<float2*>(base + off) = make_float2(a, b); }

// FUSED: the upstream gradient is the masked-MSE gradient formed from the stored trajectory (psnode_adjoint.fuse_x); a separate
// instantiation so that the plain-gx kernel keeps its register allocation (tools/sass_r2ur_check.py)
template <int METHOD, bool FUSED>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc_bwd_kernel(const __grid_constant__ TcBwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    BwdCtaSmem& sm = *reinterpret_cast<BwdCtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    // warp-level indices come from a shuffle so that ptxas knows they are warp-uniform (descriptors stay in uniform registers)
    const int cta_warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int g = cta_warp >> 3, gt = tid & 255;
    const int wk = cta_warp & 7, lane = tid & 31; // warp within the group
    const int wq = wk & 3, h = wk >> 2;           // TMEM sub-partition (== CTA warp index % 4), trajectory-column half
    const bool issuer = h == 0;                   // warps 0..3 issue the MMAs (4 partial accumulators)
    BwdGroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, S = q.S, K1 = 3 * q.S;
    const int gid = blockIdx.x * q.groups + g;
    const int b0 = gid * TN;
    const bool live = g < q.groups && b0 < B;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&sm.g[0].bar, 4); mbar_init(&sm.g[1].bar, 4); fence_mbar_init(); }
    if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, TM_COLS);
    for (int e = gt; e < (int)(offsetof(BwdGroupSmem, bar) / 4); e += GROUP_THREADS) reinterpret_cast<float*>(&gs)[e] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    const uint32_t tm_dw = tmem + TM_UPPER + (uint32_t)(g * TM_DWGROUP);     // this group's weight-gradient accumulators
    // resident transposed weights -> TMEM lanes 0..15 (warps 0..3 of group 0 write; read through the tensor core only):
    //   A[k][m] = W3[m][k], W2[m][k];  A[r][m] = (Wb+Wc)[m][r & 15]  (x columns of the folded layer 1, replicated 4 times)
    if (g == 0 && issuer) {
        const int r0 = 16 * wq + (lane >> 2), cc0 = 2 * (lane & 3);
        for (int half = 0; half < 2; half++) {
            for (int cb = 0; cb < 4; cb++) {
                float w3[8], w2[8], w1[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = r0 + ((i >> 1) & 1) * 8, col = 16 * cb + cc0 + (i & 1) + (i >> 2) * 8;   // A[row][col]
                    float hi, lo;
                    split_tf32(__ldg(q.W3 + col * TH + row), hi, lo); w3[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W2 + col * TH + row), hi, lo); w2[i] = half ? lo : hi;
                    hi = 0.0f; lo = 0.0f;
                    if ((row & 15) < X) split_tf32(__ldg(q.W1 + col * K1 + S + (row & 15)) + __ldg(q.W1 + col * K1 + 2 * S + (row & 15)), hi, lo);
                    w1[i] = half ? lo : hi;
                }
                tmem_st_16x256b_x2(tmem + lane_base + TM_W3T + 64 * half + 16 * cb, w3);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W2T + 64 * half + 16 * cb, w2);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W1T + 64 * half + 16 * cb, w1);
            }
        }
        for (int half = 0; half < 2; half++) {      // W4^T -> lanes 16..31: A[k][m] = W4[m][k], m < 16
            float w4[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = r0 + ((i >> 1) & 1) * 8, col = cc0 + (i & 1) + (i >> 2) * 8;
                float hi, lo;
                hi = 0.0f; lo = 0.0f;
                if (col < X) split_tf32(__ldg(q.W4 + col * TH + row), hi, lo);
                w4[i] = half ? lo : hi;
            }
            tmem_st_16x256b_x2(tmem + TM_UPPER + lane_base + TM_W4T + 16 * half, w4);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // fragment maps (as in the tc8 forward kernel): element i (0..3) <-> (row m0 + 8*(i>>1), trajectory 8h + c0 + (i&1))
    const int m0 = 16 * wq + (lane >> 2), c0 = 2 * (lane & 3);
    auto frag_row = [&](int i) { return m0 + (i >> 1) * 8; };
    auto frag_col = [&](int i) { return 8 * h + c0 + (i & 1); };
    int off_act[4];      // delta tile [n][m]
#pragma unroll
    for (int i = 0; i < 4; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    // [m][n] tiles: pair p = elements (2p, 2p + 1) = row m0 + 8p, trajectories 8h + c0, 8h + c0 + 1
    const int off2[2] = {(int)tile_byte(m0, 8 * h + c0, LBO_W, SBO_K16), (int)tile_byte(m0 + 8, 8 * h + c0, LBO_W, SBO_K16)};
    // the state element this thread owns in the stage algebra: state srow of trajectory column sn
    const int srow = (lane >> 2) + 8 * h, sn = c0 + (wq & 1) + 8 * (wq >> 1);
    const int off_dk = (int)tile_byte(sn, srow, LBO, SBO_ACT);
    const int off_sq = (int)tile_byte(srow, sn, LBO_W, SBO_K16);   // [state][n]
    const int ftape = (32 * wq + lane) * 8 + 4 * h;                // this thread's 4 floats of a tape fragment block
    const int ytape = 3 * PSN_TAPE_FRAG + (32 * wq + lane) * 2 + h;

    // descriptors
    const uint32_t idesc16 = make_idesc_tf32(TH, TN), idesc24 = make_idesc_tf32(TH, TK1), idesc64 = make_idesc_tf32(TH, TH);
    // two base descriptors; every other tile shares one of the two (LBO, SBO) pairs and sits at a constant byte offset, so its
    // descriptor is base + (offset >> 4) in the start-address field: ptxas keeps 2 descriptors in uniform registers instead of
    // rebuilding 14 of them (shift / mask / or chains) in front of every MMA batch
    const uint64_t d_dT_hi = make_desc(smem_u32(gs.dT_hi), LBO, SBO_ACT);
    const uint64_t d_k16 = make_desc(smem_u32(gs.dA_hi[0]), LBO_W, SBO_K16);
    auto at = [](uint64_t base, size_t from, size_t to) { return base + (uint64_t)((to - from) >> 4); };
    constexpr size_t o_dA = offsetof(BwdGroupSmem, dA_hi);
    const uint64_t d_dT_lo = at(d_dT_hi, offsetof(BwdGroupSmem, dT_hi), offsetof(BwdGroupSmem, dT_lo));
    const uint64_t d_dA_hi0 = d_k16, d_dA_hi1 = at(d_k16, 0, K16_TILE);
    const uint64_t d_dA_lo0 = at(d_k16, o_dA, offsetof(BwdGroupSmem, dA_lo)), d_dA_lo1 = at(d_k16, o_dA, offsetof(BwdGroupSmem, dA_lo) + K16_TILE);
    const uint64_t d_aB_hi0 = at(d_k16, o_dA, offsetof(BwdGroupSmem, aB_hi)), d_aB_hi1 = at(d_k16, o_dA, offsetof(BwdGroupSmem, aB_hi) + K16_TILE);
    const uint64_t d_aB_lo0 = at(d_k16, o_dA, offsetof(BwdGroupSmem, aB_lo)), d_aB_lo1 = at(d_k16, o_dA, offsetof(BwdGroupSmem, aB_lo) + K16_TILE);
    const uint64_t d_dkB_hi = at(d_k16, o_dA, offsetof(BwdGroupSmem, dkB_hi)), d_dkB_lo = at(d_k16, o_dA, offsetof(BwdGroupSmem, dkB_lo));
    const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4) * TN;
    const uint32_t my_acc = acc_base + (uint32_t)wq * TN;
    const uint32_t acc_m1 = tmem + TM_UPPER + TM_ACC_M1 + (uint32_t)(g * 4) * TN;      // accumulators of g3 = W4^T dk (lanes 16..31)
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
    uint32_t phase = 0;

    // ---- helpers ---------------------------------------------------------------------------------------
    // data MMA with K = 64 (A resident in TMEM): issuing warp wq takes K-steps 2wq, 2wq+1 of the three 3xTF32 terms, then commits
    auto issue_data = [&](uint32_t w_tm) {          // w_tm: TMEM column of the resident operand (hi; lo at +64)
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int term = 0; term < 3; term++) {
                    const uint32_t wa = term == 0 ? w_tm + 64 : w_tm;
                    const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
                    for (int kk = 0; kk < 2; kk++) {
                        const int ks = 2 * wq + kk;
                        mma_tf32_ts(my_acc, tmem + wa + 8 * ks, bd + KSTEP_B * ks, idesc16, accumulate);
                        accumulate = 1;
                    }
                }
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    // g3 = W4^T dk (K = 16): six MMAs, entry e = 2*term + kstep, issuing warp wq takes e = wq and e = wq + 4
    auto issue_m1 = [&]() {
        if (issuer) {
            if (elect_one()) {
                tc_fence_after();
                uint32_t accumulate = 0;
                for (int e = wq; e < 6; e += 4) {
                    const int term = e >> 1, ks = e & 1;
                    const uint32_t wa = TM_W4T + (term == 0 ? 16 : 0);
                    const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
                    mma_tf32_ts(acc_m1 + (uint32_t)wq * TN, tmem + TM_UPPER + wa + 8 * ks, bd + KSTEP_B * ks, idesc16, accumulate);
                    accumulate = 1;
                }
                mma_commit(&gs.bar);
            }
            __syncwarp();
        }
    };
    // weight-gradient chain (K = 16 trajectories), accumulated in TMEM; issued AFTER this warp's commit: off the critical path
    // (`fresh`: first chain after a flush overwrites the accumulator)
    auto issue_dw = [&](uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t d_tmem, uint32_t idesc, bool fresh) {
        if (elect_one()) {
            uint32_t accumulate = fresh ? 0u : 1u;
            for (int term = 0; term < 3; term++) {
                const uint64_t ad = term == 0 ? a_lo : a_hi;
                const uint64_t bd = term == 1 ? b_lo : b_hi;
                for (int ks = 0; ks < 2; ks++) {
                    mma_tf32(d_tmem, ad + KSTEP_W * ks, bd + KSTEP_W * ks, idesc, accumulate);
                    accumulate = 1u;
                }
            }
        }
        __syncwarp();
    };
    auto wait_mma = [&]() {
        if (!mbar_wait(&gs.bar, phase)) { atomicExch(q.err, 1); __trap(); }
        phase ^= 1;
        tc_fence_after();
    };
    auto collect = [&](float (&d)[4], bool m1 = false) {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = (m1 ? acc_m1 : acc_base) + lane_base + 8 * h;
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
    };
    // dL/dy element (state srow, trajectory sn) of this thread; every 16-row block holds the same 16 x 16 tile
    auto collect_gy = [&]() {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc_base + lane_base + 8 * (wq >> 1);
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
        const int sel = 2 * h + (wq & 1);
        float sv[4];
#pragma unroll
        for (int i = 0; i < 4; i++) sv[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
        return sel == 0 ? sv[0] : (sel == 1 ? sv[1] : (sel == 2 ? sv[2] : sv[3]));
    };
    auto publish = [&]() {
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
    };
    // fragment -> [row][n] tile (hi / lo), 8-byte stores of the two adjacent trajectories
    auto store_pairs = [&](unsigned char* hi_t, unsigned char* lo_t, const float (&v)[4]) {
#pragma unroll
        for (int p = 0; p < 2; p++) {
            float h0, l0, h1, l1;
            split_tf32_fast(v[2 * p], h0, l0);
            split_tf32_fast(v[2 * p + 1], h1, l1);
            st_f32x2(hi_t, off2[p], h0, h1);
            st_f32x2(lo_t, off2[p], l0, l1);
        }
    };
    // delta = g * elu'(a); delta -> dT tile [n][m] and dA tile [m][n]; running sums for the bias gradient
    auto make_delta = [&](const float (&gsum)[4], const float (&act)[4], float (&bsum)[4], unsigned char* dA_hi, unsigned char* dA_lo) {
        float d[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            d[i] = gsum[i] * psn_elu_grad_from_out(act[i]);
            bsum[i] += d[i];
            float hi, lo;
            split_tf32_fast(d[i], hi, lo);
            st_f32(gs.dT_hi, off_act[i], hi);
            st_f32(gs.dT_lo, off_act[i], lo);
        }
        store_pairs(dA_hi, dA_lo, d);
    };
    auto ld_frag = [&](const float* src, float (&v)[4]) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src + ftape));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    };
    // held inputs of the step that ends at grid point j (warp 4, lane = trajectory), as in the forward kernel
    auto load_held = [&](int j, float (&u)[TU]) {
        const int bb = min(b0 + (lane & 15), B - 1);
        const int k = q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1;
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, j - 1, bb, c);
        }
    };

    float* sl = q.slab + (int64_t)gid * (pad4(q.n_theta) + G_AREA);
    float* garea = sl + pad4(q.n_theta);        // [64][24] running dW1f
    const int oW1 = 0, ob1 = TH * K1, oW2 = ob1 + TH, ob2 = oW2 + TH * TH, oW3 = ob2 + TH, ob3 = oW3 + TH * TH, oW4 = ob3 + TH,
              ob4 = oW4 + X * TH;

    // drain the tensor pipe and add the TMEM weight-gradient accumulators into the slab (round-to-nearest fp32 adds; every
    // element is owned by one thread, the slab was zeroed by the launcher).  Warp (wq, h) owns rows 16wq.. and the 8-column
    // blocks of parity h.
    auto flush_dw = [&]() {
        if (issuer) {
            if (elect_one()) { tc_fence_after(); mma_commit(&gs.bar); }
            __syncwarp();
        }
        wait_mma();
        // the slab values are fetched first, all at once (independent L2 round trips), then combined with the accumulators
        const int r0 = m0, cc = c0;
        float2 o2[4][2], o3[4][2], og[2][2];
        float o4[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {          // dW2 / dW3: 64 columns = 8 blocks of 8, this warp takes blocks h, h+2, h+4, h+6
            const int cb = h + 2 * q4;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                o2[q4][r] = *reinterpret_cast<const float2*>(sl + oW2 + (r0 + 8 * r) * TH + 8 * cb + cc);
                o3[q4][r] = *reinterpret_cast<const float2*>(sl + oW3 + (r0 + 8 * r) * TH + 8 * cb + cc);
            }
        }
        const int mc = 8 * h + cc;                // dW4^T: rows = hidden k, columns = output m (16): block h
        o4[0] = mc < X ? sl[oW4 + mc * TH + r0] : 0.0f; o4[1] = mc + 1 < X ? sl[oW4 + (mc + 1) * TH + r0] : 0.0f;
        o4[2] = mc < X ? sl[oW4 + mc * TH + r0 + 8] : 0.0f; o4[3] = mc + 1 < X ? sl[oW4 + (mc + 1) * TH + r0 + 8] : 0.0f;
#pragma unroll
        for (int q2 = 0; q2 < 2; q2++) {          // dW1f: 24 columns = 3 blocks of 8: blocks h and h + 2 (< 3)
            const int cb = h + 2 * q2;
#pragma unroll
            for (int r = 0; r < 2; r++)
                og[q2][r] = cb < 3 ? *reinterpret_cast<const float2*>(garea + (r0 + 8 * r) * TK1 + 8 * cb + cc) : make_float2(0.f, 0.f);
        }
        float v[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            const int cb = h + 2 * q4;
            tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW2 + 8 * cb, v);
            tmem_ld_wait();
            *reinterpret_cast<float2*>(sl + oW2 + r0 * TH + 8 * cb + cc) = make_float2(o2[q4][0].x + v[0], o2[q4][0].y + v[1]);
            *reinterpret_cast<float2*>(sl + oW2 + (r0 + 8) * TH + 8 * cb + cc) = make_float2(o2[q4][1].x + v[2], o2[q4][1].y + v[3]);
            tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW3 + 8 * cb, v);
            tmem_ld_wait();
            *reinterpret_cast<float2*>(sl + oW3 + r0 * TH + 8 * cb + cc) = make_float2(o3[q4][0].x + v[0], o3[q4][0].y + v[1]);
            *reinterpret_cast<float2*>(sl + oW3 + (r0 + 8) * TH + 8 * cb + cc) = make_float2(o3[q4][1].x + v[2], o3[q4][1].y + v[3]);
        }
        tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW4T + 8 * h, v);
        tmem_ld_wait();
        if (mc < X) { sl[oW4 + mc * TH + r0] = o4[0] + v[0]; sl[oW4 + mc * TH + r0 + 8] = o4[2] + v[2]; }
        if (mc + 1 < X) { sl[oW4 + (mc + 1) * TH + r0] = o4[1] + v[1]; sl[oW4 + (mc + 1) * TH + r0 + 8] = o4[3] + v[3]; }
#pragma unroll
        for (int q2 = 0; q2 < 2; q2++) {
            const int cb = h + 2 * q2;
            if (cb < 3) {                          // warp-uniform
                tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW1F + 8 * cb, v);
                tmem_ld_wait();
                *reinterpret_cast<float2*>(garea + r0 * TK1 + 8 * cb + cc) = make_float2(og[q2][0].x + v[0], og[q2][0].y + v[1]);
                *reinterpret_cast<float2*>(garea + (r0 + 8) * TK1 + 8 * cb + cc) = make_float2(og[q2][1].x + v[2], og[q2][1].y + v[3]);
            }
        }
    };

    if (live) {
        bool fresh = true;
        const int bown = b0 + sn, bbown = min(bown, B - 1);
        const bool valid = bown < B && srow < X;
        float lam, D1[4], dB2[4], dB3[4], dB4 = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) { D1[i] = 0.f; dB2[i] = 0.f; dB3[i] = 0.f; }
        auto up_x = [&](int j) -> float {        // dL/dx_sol[j] of this thread's state element (operands re-read from the parameters)
            if (!valid) return 0.0f;
            if constexpr (FUSED) return psn_fuse_grad(q.fx, psn_fuse_scale(q.fx), j, bown, srow);
            else return q.gx.p ? ldser(q.gx, j, bown, srow) : 0.0f;
        };
        lam = up_x(T - 1);

        const float c13 = (float)(1.0 / 3.0);
        const float* tp = q.tape + ((int64_t)gid * (T - 1) * NST + (int64_t)(T - 1) * NST - 1) * PSN_TAPE_STAGE;   // last record
        float a1[4], a2[4], a3[4], yv = 0.0f;
        float tan = 0.0f, tbn = 0.0f, un[TU];       // grid times of the next step to process (subtracted when it starts)
#pragma unroll
        for (int c = 0; c < TU; c++) un[c] = 0.0f;
        if (T > 1) {
            ld_frag(tp, a1); ld_frag(tp + PSN_TAPE_FRAG, a2); ld_frag(tp + 2 * PSN_TAPE_FRAG, a3);
            yv = __ldcs(tp + ytape);
            tan = ldser(q.t, T - 1, bbown, 0); tbn = ldser(q.t, T - 2, bbown, 0);
            if (wk == 4) load_held(T - 1, un);
        }

        for (int j = T - 1; j >= 1; j--) {
            const float dt = __fsub_rn(tan, tbn);
            float u[TU];
#pragma unroll
            for (int c = 0; c < TU; c++) u[c] = un[c];
            const float gxn = up_x(j - 1);
            if (j > 1) {
                tan = tbn; tbn = ldser(q.t, j - 2, bbown, 0);
                if (wk == 4) load_held(j - 1, un);
            }
            // x_j = x_{j-1} + dt * sum_s bw[s] k_s : dL/dk_s starts at lam * dt * bw[s]; dxs collects dL/dx_{j-1}
            float dxs = lam, d1 = 0.f, d2 = 0.f, d3 = 0.f, dcur;
            {
                const float ld = lam * dt;
                if (METHOD == PSNODE_RK4) { d1 = ld * 0.125f; d2 = ld * 0.375f; d3 = ld * 0.375f; dcur = ld * 0.125f; }
                else dcur = ld;
            }
#pragma unroll 1
            for (int e = NST - 1; e >= 0; e--) {
                const bool has_next = !(j == 1 && e == 0);
                const float* tpn = tp - PSN_TAPE_STAGE;
                float gsum[4];
                // ---- P0: dk tiles, a3 -> aB[0]; g3 = W4^T dk ; dW4^T += a3 dk^T ----
                {
                    float hi, lo;
                    split_tf32_fast(dcur, hi, lo);
                    st_f32(gs.dT_hi, off_dk, hi); st_f32(gs.dT_lo, off_dk, lo);
                    st_f32(gs.dkB_hi, off_sq, hi); st_f32(gs.dkB_lo, off_sq, lo);
                    dB4 += dcur;
                }
                store_pairs(gs.aB_hi[0], gs.aB_lo[0], a3);
                publish();
                issue_m1();
                if (wk == 0) issue_dw(d_aB_hi0, d_aB_lo0, d_dkB_hi, d_dkB_lo, tm_dw + TM_DW4T, idesc16, fresh);
                // ---- P1: d3 ; a2 -> aB[1]; g2 = W3^T d3 ; dW3 += d3 a2^T ----
                collect(gsum, true);
                make_delta(gsum, a3, dB3, gs.dA_hi[1], gs.dA_lo[1]);
                store_pairs(gs.aB_hi[1], gs.aB_lo[1], a2);
                if (has_next) ld_frag(tpn + 2 * PSN_TAPE_FRAG, a3);
                publish();
                issue_data(TM_W3T);
                if (wk == 1) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW3, idesc64, fresh);
                // ---- P2: d2 ; a1 -> aB[0]; g1 = W2^T d2 ; dW2 += d2 a1^T ----
                collect(gsum);
                make_delta(gsum, a2, dB2, gs.dA_hi[0], gs.dA_lo[0]);
                store_pairs(gs.aB_hi[0], gs.aB_lo[0], a1);
                if (has_next) ld_frag(tpn + PSN_TAPE_FRAG, a2);
                publish();
                issue_data(TM_W2T);
                if (wk == 2) issue_dw(d_dA_hi0, d_dA_lo0, d_aB_hi0, d_aB_lo0, tm_dw + TM_DW2, idesc64, fresh);
                // ---- P3: d1 ; [y; u] -> aB[1] rows 0..23; gy = (Wb+Wc)_x^T d1 ; dW1f += d1 [y;u]^T ----
                collect(gsum);
                make_delta(gsum, a1, D1, gs.dA_hi[1], gs.dA_lo[1]);
                {
                    float hi, lo;
                    split_tf32_fast(yv, hi, lo);
                    st_f32(gs.aB_hi[1], off_sq, hi); st_f32(gs.aB_lo[1], off_sq, lo);
                }
                if (wk == 4 && lane < TN) {
#pragma unroll
                    for (int c = 0; c < TU; c++) {
                        float hi, lo;
                        split_tf32_fast(u[c], hi, lo);
                        const int o = tile_byte(TX + c, lane, LBO_W, SBO_K16);
                        st_f32(gs.aB_hi[1], o, hi); st_f32(gs.aB_lo[1], o, lo);
                    }
                }
                if (has_next) {
                    ld_frag(tpn, a1);
                    yv = __ldcs(tpn + ytape);
                }
                publish();
                issue_data(TM_W1T);
                if (wk == 3) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW1F, idesc24, fresh);
                // ---- P4: dL/dy of this stage -> Runge-Kutta adjoint algebra (my_fixed_grid.py:15-59 reversed) ----
                const float gq = collect_gy();
                dxs += gq;
                if (METHOD == PSNODE_RK4) {
                    const float tg = dt * gq;
                    if (e == 3) { d3 += tg; d2 -= tg; d1 += tg; dcur = d3; }
                    else if (e == 2) { d2 += tg; d1 -= tg * c13; dcur = d2; }
                    else if (e == 1) { d1 += tg * c13; dcur = d1; }
                } else if (METHOD == PSNODE_MIDPOINT) {
                    if (e == 1) dcur = (0.5f * dt) * gq;
                }
                tp = tpn;
                fresh = false;
            }
            lam = dxs + gxn;
            if (((T - j) % (PSN_DW_CHAIN / NST)) == 0 || j == 1) { flush_dw(); fresh = true; }
        }
        if (q.d_x0 && valid) q.d_x0[(int64_t)bown * q.d_x0_sb + srow] = lam;

        // ---- the weight gradients are in the slab already (last flush); bias gradients and layer-1 unfolding follow ----
        group_sync(g);
        float* scr = reinterpret_cast<float*>(gs.dA_hi[0]);          // 32 KB of dead tiles: scratch
        float* D1s = scr;                 // [64][17]
        float* B2s = D1s + TH * 17;       // [64][17]
        float* B3s = B2s + TH * 17;       // [64][17]
        float* Gs = B3s + TH * 17;        // [64][25]
        float* a0s = Gs + TH * 25;        // [16][25]
        float* dks = a0s + TN * 25;       // [16][17]   dB4 per (state, trajectory)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int o = frag_row(i) * 17 + frag_col(i);
            D1s[o] = D1[i]; B2s[o] = dB2[i]; B3s[o] = dB3[i];
        }
        dks[srow * 17 + sn] = dB4;
        for (int e = gt; e < TH * TK1; e += GROUP_THREADS) Gs[(e / TK1) * 25 + (e % TK1)] = garea[e];   // flushed before the group_sync above
        for (int e = gt; e < TN * S; e += GROUP_THREADS) {
            const int n = e / S, c = e - n * S;
            a0s[n * 25 + c] = __ldg(q.a0 + (int64_t)min(b0 + n, B - 1) * q.a0_sb + c);
        }
        group_sync(g);
        if (gt < 3 * TH) {                 // bias gradients of layers 1..3: row sums over the 16 trajectories
            const int which = gt >> 6, m = gt & 63;
            const float* src = (which == 0 ? D1s : (which == 1 ? B2s : B3s)) + m * 17;
            float acc = 0.0f;
#pragma unroll
            for (int n = 0; n < TN; n++) acc += src[n];
            sl[(which == 0 ? ob1 : (which == 1 ? ob2 : ob3)) + m] = acc;
        } else if (gt < 3 * TH + X) {
            const int c = gt - 3 * TH;
            float acc = 0.0f;
#pragma unroll
            for (int n = 0; n < TN; n++) acc += dks[c * 17 + n];
            sl[ob4 + c] = acc;
        }
        // unfold layer 1: W1 = [Wa | Wb | Wc] acting on [a0; s - a0; s]
        for (int e = gt; e < TH * S; e += GROUP_THREADS) {
            const int m = e / S, c = e - m * S;
            float P = 0.0f;
#pragma unroll
            for (int n = 0; n < TN; n++) P = fmaf(D1s[m * 17 + n], a0s[n * 25 + c], P);
            const float G = Gs[m * 25 + (c < X ? c : TX + (c - X))];      // tile column of s[c]: x in 0..15, held inputs from 16
            sl[oW1 + m * K1 + c] = P;
            sl[oW1 + m * K1 + S + c] = G - P;
            sl[oW1 + m * K1 + 2 * S + c] = G;
        }
        if (q.d_a0) {
            for (int e = gt; e < TN * S; e += GROUP_THREADS) {
                const int n = e / S, c = e - n * S, b = b0 + n;
                if (b >= B) continue;
                float acc = 0.0f;
                for (int m = 0; m < TH; m++)
                    acc = fmaf(__ldg(q.W1 + m * K1 + c) - __ldg(q.W1 + m * K1 + S + c), D1s[m * 17 + n], acc);
                q.d_a0[(int64_t)b * q.d_a0_sb + c] = acc;
            }
        }
    }
    // ---- teardown --------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) tmem_dealloc(tmem, TM_COLS);
}

// d_theta[i] = sum over the group slabs, fixed order (deterministic)
__global__ void psn_tc_grad_reduce_kernel(const float* __restrict__ slab, int n_slabs, int n_theta, int stride, float* __restrict__ d_theta) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_theta; i += gridDim.x * blockDim.x) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int s = 0;
        for (; s + 3 < n_slabs; s += 4) {
            a0 += slab[(size_t)s * stride + i];
            a1 += slab[(size_t)(s + 1) * stride + i];
            a2 += slab[(size_t)(s + 2) * stride + i];
            a3 += slab[(size_t)(s + 3) * stride + i];
        }
        for (; s < n_slabs; s++) a0 += slab[(size_t)s * stride + i];
        d_theta[i] = (a0 + a1) + (a2 + a3);
    }
}

}  // namespace

bool psn_tc_supports(const psnode_problem* p);

// the tape-based reverse sweep covers what the scripts' ODE training step needs: parameter gradients, d_x0, d_a0
bool psn_tc_bwd_supports(const psnode_problem* p, const psnode_adjoint* a) {
    if (p->kind != PSNODE_ODE || !psn_tc_supports(p) || !p->tape) return false;
    if (p->tape_floats < psn_tc_tape_floats(p->B, p->T, p->method)) return false;
    if (a->d_z.p || a->d_v.p || a->d_zjump || a->d_vjump || a->d_xteach.p || a->d_iteach.p) return false;
    return true;
}

int64_t psn_tc_backward_workspace(const psnode_problem* p, const psnode_adjoint*) {
    const int64_t n_theta = psnode_mlp_param_count(&p->de);
    return 256 + (int64_t)psn_tc_ngroups(p->B) * (pad4((int)n_theta) + G_AREA) * 4;
}

int psn_tc_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < psn_tc_backward_workspace(p, a)) return PSNODE_EWORKSPACE;
    const int64_t n_theta = psnode_mlp_param_count(&p->de);
    if (a->n_theta != n_theta) return PSNODE_EINVAL;
    TcBwdParams q;
    q.B = p->B; q.T = p->T; q.X = p->X; q.Z = p->Z; q.S = p->X + p->Z;
    q.n_theta = (int)n_theta;
    q.t = p->t; q.z = p->z; q.gx = a->gx;
    q.fx = psn_make_fuse(a->fuse_x, p->x_sol);
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.W1 = p->de.W[0]; q.W2 = p->de.W[1]; q.W3 = p->de.W[2]; q.W4 = p->de.W[3];
    q.tape = p->tape;
    q.err = static_cast<int*>(ws);
    q.slab = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
    q.d_x0 = a->d_x0; q.d_x0_sb = a->d_x0_sb;
    q.d_a0 = a->d_a0; q.d_a0_sb = a->d_a0_sb;
    const int ngroups = psn_tc_ngroups(p->B);
    q.groups = psn_tc_groups_per_cta(p->B);
    const int grid = (ngroups + q.groups - 1) / q.groups;
    PSN_CUDA(cudaMemsetAsync(ws, 0, 256 + (size_t)ngroups * (pad4((int)n_theta) + G_AREA) * 4, stream));
    const int smem = (int)sizeof(BwdCtaSmem) + 128;
    auto launch = [&](auto kern, const char* name) -> int {
        PSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, 2 * GROUP_THREADS, smem, stream>>>(q);
        psn_count_launch(name);
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    };
    int st;
    switch (p->method) {
        case PSNODE_EULER: st = a->fuse_x.target.p ? launch(psn_tc_bwd_kernel<PSNODE_EULER, true>, "psn_tc_bwd_kernel<euler,fused-loss>")
                                                   : launch(psn_tc_bwd_kernel<PSNODE_EULER, false>, "psn_tc_bwd_kernel<euler>"); break;
        case PSNODE_MIDPOINT: st = a->fuse_x.target.p ? launch(psn_tc_bwd_kernel<PSNODE_MIDPOINT, true>, "psn_tc_bwd_kernel<midpoint,fused-loss>")
                                                      : launch(psn_tc_bwd_kernel<PSNODE_MIDPOINT, false>, "psn_tc_bwd_kernel<midpoint>"); break;
        default: st = a->fuse_x.target.p ? launch(psn_tc_bwd_kernel<PSNODE_RK4, true>, "psn_tc_bwd_kernel<rk4,fused-loss>")
                                         : launch(psn_tc_bwd_kernel<PSNODE_RK4, false>, "psn_tc_bwd_kernel<rk4>"); break;
    }
    if (st != PSNODE_OK) return st;
    psn_tc_grad_reduce_kernel<<<32, 256, 0, stream>>>(q.slab, ngroups, (int)n_theta, pad4((int)n_theta) + G_AREA, a->d_theta);
    psn_count_launch("psn_tc_grad_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
