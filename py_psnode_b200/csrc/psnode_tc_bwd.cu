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
//     adds every PSN_DW_FLUSH steps (<= 16 accumulations per chain) instead of once at the end.
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
constexpr int TM_COLS = 512;
constexpr int GROUP_THREADS = 128;
constexpr int PSN_DW_FLUSH = 4;        // steps between two flushes of the TMEM weight-gradient accumulators
constexpr int G_AREA = TH * TK1;       // floats of the dW1f (folded layer 1) accumulator kept behind each group's slab

struct TcBwdParams {
    int B, T, Z, S, groups, n_theta;
    psnode_series t, z, gx;
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
    unsigned char w4t_hi[K16_TILE], w4t_lo[K16_TILE];     // A[k][m] = W4[m][k], K = 16
    BwdGroupSmem g[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }
__device__ __forceinline__ void st_f32x2(unsigned char* base, int off, float a, float b) { *reinterpret_cast<float2*>(base + off) = make_float2(a, b); }

template <int METHOD>
__global__ void __launch_bounds__(2 * GROUP_THREADS, 1) psn_tc_bwd_kernel(const __grid_constant__ TcBwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    BwdCtaSmem& sm = *reinterpret_cast<BwdCtaSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    const int g = tid >> 7, gt = tid & 127;
    const int warp = gt >> 5, lane = gt & 31;     // warp within the group == TMEM sub-partition
    BwdGroupSmem& gs = sm.g[g];
    const int B = q.B, T = q.T, Z = q.Z, S = q.S, K1 = 3 * q.S;
    const int gid = blockIdx.x * q.groups + g;
    const int b0 = gid * TN;
    const bool live = g < q.groups && b0 < B;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&sm.g[0].bar, 4); mbar_init(&sm.g[1].bar, 4); fence_mbar_init(); }
    if ((tid >> 5) == 0) tmem_alloc(&sm.tmem_base, TM_COLS);
    for (int e = tid; e < TX * TH; e += 2 * GROUP_THREADS) {       // e = m * 64 + k, m < 16
        const int m = e >> 6, k = e & 63;
        float hi, lo;
        split_tf32(__ldg(q.W4 + e), hi, lo);
        const int o = tile_byte(k, m, LBO_W, SBO_K16);
        st_f32(sm.w4t_hi, o, hi); st_f32(sm.w4t_lo, o, lo);
    }
    for (int e = gt; e < (int)(offsetof(BwdGroupSmem, bar) / 4); e += GROUP_THREADS) reinterpret_cast<float*>(&gs)[e] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    const uint32_t tm_dw = tmem + TM_UPPER + (uint32_t)(g * TM_DWGROUP);     // this group's weight-gradient accumulators
    // resident transposed weights -> TMEM lanes 0..15 (group 0 writes; read through the tensor core only):
    //   A[k][m] = W3[m][k], W2[m][k];  A[r][m] = (Wb+Wc)[m][r & 15]  (x columns of the folded layer 1, replicated 4 times)
    if (g == 0) {
        const int r0 = 16 * warp + (lane >> 2), cc0 = 2 * (lane & 3);
        for (int half = 0; half < 2; half++) {
            for (int cb = 0; cb < 4; cb++) {
                float w3[8], w2[8], w1[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = r0 + ((i >> 1) & 1) * 8, col = 16 * cb + cc0 + (i & 1) + (i >> 2) * 8;   // A[row][col]
                    float hi, lo;
                    split_tf32(__ldg(q.W3 + col * TH + row), hi, lo); w3[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W2 + col * TH + row), hi, lo); w2[i] = half ? lo : hi;
                    split_tf32(__ldg(q.W1 + col * K1 + S + (row & 15)) + __ldg(q.W1 + col * K1 + 2 * S + (row & 15)), hi, lo);
                    w1[i] = half ? lo : hi;
                }
                tmem_st_16x256b_x2(tmem + lane_base + TM_W3T + 64 * half + 16 * cb, w3);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W2T + 64 * half + 16 * cb, w2);
                tmem_st_16x256b_x2(tmem + lane_base + TM_W1T + 64 * half + 16 * cb, w1);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // fragment maps (as in the forward kernel): element i <-> (row m0 + 8*((i>>1)&1), trajectory c0 + (i&1) + 8*(i>>2))
    const int m0 = 16 * warp + (lane >> 2), c0 = 2 * (lane & 3);
    auto frag_row = [&](int i) { return m0 + ((i >> 1) & 1) * 8; };
    auto frag_col = [&](int i) { return c0 + (i & 1) + (i >> 2) * 8; };
    int off_act[8];      // delta tile [n][m]
#pragma unroll
    for (int i = 0; i < 8; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    int off2[4];         // [m][n] tiles: pair p = elements (ia, ia + 1), ia = 2*(p&1) + 4*(p>>1)
#pragma unroll
    for (int p = 0; p < 4; p++) off2[p] = tile_byte(m0 + 8 * (p & 1), c0 + 8 * (p >> 1), LBO_W, SBO_K16);
    // the two state elements this thread owns in the stage algebra: states sm0, sm0 + 8 of trajectory column sn
    const int sm0 = lane >> 2, sn = c0 + (warp & 1) + 8 * (warp >> 1);
    const int off_dk[2] = {(int)tile_byte(sn, sm0, LBO, SBO_ACT), (int)tile_byte(sn, sm0 + 8, LBO, SBO_ACT)};
    const int off_sq[2] = {(int)tile_byte(sm0, sn, LBO_W, SBO_K16), (int)tile_byte(sm0 + 8, sn, LBO_W, SBO_K16)};   // [state][n]

    // descriptors
    const uint32_t idesc16 = make_idesc_tf32(TH, TN), idesc24 = make_idesc_tf32(TH, TK1), idesc64 = make_idesc_tf32(TH, TH);
    const uint64_t d_dT_hi = make_desc(smem_u32(gs.dT_hi), LBO, SBO_ACT), d_dT_lo = make_desc(smem_u32(gs.dT_lo), LBO, SBO_ACT);
    const uint64_t d_w4t_hi = make_desc(smem_u32(sm.w4t_hi), LBO_W, SBO_K16), d_w4t_lo = make_desc(smem_u32(sm.w4t_lo), LBO_W, SBO_K16);
    const uint64_t d_dA_hi0 = make_desc(smem_u32(gs.dA_hi[0]), LBO_W, SBO_K16), d_dA_lo0 = make_desc(smem_u32(gs.dA_lo[0]), LBO_W, SBO_K16);
    const uint64_t d_dA_hi1 = make_desc(smem_u32(gs.dA_hi[1]), LBO_W, SBO_K16), d_dA_lo1 = make_desc(smem_u32(gs.dA_lo[1]), LBO_W, SBO_K16);
    const uint64_t d_aB_hi0 = make_desc(smem_u32(gs.aB_hi[0]), LBO_W, SBO_K16), d_aB_lo0 = make_desc(smem_u32(gs.aB_lo[0]), LBO_W, SBO_K16);
    const uint64_t d_aB_hi1 = make_desc(smem_u32(gs.aB_hi[1]), LBO_W, SBO_K16), d_aB_lo1 = make_desc(smem_u32(gs.aB_lo[1]), LBO_W, SBO_K16);
    const uint64_t d_dkB_hi = make_desc(smem_u32(gs.dkB_hi), LBO_W, SBO_K16), d_dkB_lo = make_desc(smem_u32(gs.dkB_lo), LBO_W, SBO_K16);
    const uint32_t acc_base = tmem + TM_ACC + (uint32_t)(g * 4) * TN;
    const uint32_t my_acc = acc_base + (uint32_t)warp * TN;
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
    uint32_t phase = 0;

    // ---- helpers ---------------------------------------------------------------------------------------
    // data MMA with K = 64: this warp's two K-steps for the three 3xTF32 terms (small terms first), then commit
    auto issue_data = [&](uint32_t w_tm) {          // w_tm: TMEM column of the resident operand (hi; lo at +64)
        if (elect_one()) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int term = 0; term < 3; term++) {
                const uint32_t wa = term == 0 ? w_tm + 64 : w_tm;
                const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
                for (int kk = 0; kk < 2; kk++) {
                    const int ks = 2 * warp + kk;
                    mma_tf32_ts(my_acc, tmem + wa + 8 * ks, bd + KSTEP_B * ks, idesc16, accumulate);
                    accumulate = 1;
                }
            }
            mma_commit(&gs.bar);
        }
        __syncwarp();
    };
    // g3 = W4^T dk (K = 16): six MMAs, entry e = 2*term + kstep, warp w takes e = w and e = w + 4
    auto issue_m1 = [&]() {
        if (elect_one()) {
            tc_fence_after();
            uint32_t accumulate = 0;
            for (int e = warp; e < 6; e += 4) {
                const int term = e >> 1, ks = e & 1;
                const uint64_t ad = term == 0 ? d_w4t_lo : d_w4t_hi;
                const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
                mma_tf32(my_acc, ad + KSTEP_W * ks, bd + KSTEP_B * ks, idesc16, accumulate);
                accumulate = 1;
            }
            mma_commit(&gs.bar);
        }
        __syncwarp();
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
    auto collect = [&](float (&d)[8]) {
        wait_mma();
        float t0[8], t1[8], t2[8], t3[8];
        tmem_ld_16x256b_x2(acc_base + lane_base + 0 * TN, t0);
        tmem_ld_16x256b_x2(acc_base + lane_base + 1 * TN, t1);
        tmem_ld_16x256b_x2(acc_base + lane_base + 2 * TN, t2);
        tmem_ld_16x256b_x2(acc_base + lane_base + 3 * TN, t3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
    };
    // dL/dy elements (state sm0 / sm0 + 8, trajectory sn) of this thread; every 16-row block holds the same 16 x 16 tile
    auto collect_gy = [&](float (&kv)[2]) {
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
    auto publish = [&]() {
        fence_async_smem();
        tc_fence_before();
        group_sync(g);
    };
    // fragment -> [row][n] tile (hi / lo), 8-byte stores of the two adjacent trajectories
    auto store_pairs = [&](unsigned char* hi_t, unsigned char* lo_t, const float (&v)[8]) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int ia = 2 * (p & 1) + 4 * (p >> 1);
            float h0, l0, h1, l1;
            split_tf32_fast(v[ia], h0, l0);
            split_tf32_fast(v[ia + 1], h1, l1);
            st_f32x2(hi_t, off2[p], h0, h1);
            st_f32x2(lo_t, off2[p], l0, l1);
        }
    };
    // delta = g * elu'(a); delta -> dT tile [n][m] and dA tile [m][n]; running sums for the bias gradient
    auto make_delta = [&](const float (&gsum)[8], const float (&act)[8], float (&bsum)[8], unsigned char* dA_hi, unsigned char* dA_lo) {
        float d[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            d[i] = gsum[i] * psn_elu_grad_from_out(act[i]);
            bsum[i] += d[i];
            float hi, lo;
            split_tf32_fast(d[i], hi, lo);
            st_f32(gs.dT_hi, off_act[i], hi);
            st_f32(gs.dT_lo, off_act[i], lo);
        }
        store_pairs(dA_hi, dA_lo, d);
    };
    auto ld_frag = [&](const float* src, float (&v)[8]) {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(src + gt * 8));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(src + gt * 8) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    };
    // held inputs of the step that ends at grid point j (warp 1, lane = trajectory), as in the forward kernel
    auto load_held = [&](int j, float (&u)[TU]) {
        const int bb = min(b0 + (lane & 15), B - 1);
        const int k = q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1;
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, j - 1, bb, c);
        }
    };

    float* sl = q.slab + (int64_t)gid * (q.n_theta + G_AREA);
    float* garea = sl + q.n_theta;        // [64][24] running dW1f
    const int oW1 = 0, ob1 = TH * K1, oW2 = ob1 + TH, ob2 = oW2 + TH * TH, oW3 = ob2 + TH, ob3 = oW3 + TH * TH, oW4 = ob3 + TH,
              ob4 = oW4 + TX * TH;

    // drain the tensor pipe and add the TMEM weight-gradient accumulators into the slab (round-to-nearest fp32 adds; every
    // element is owned by one thread, the slab was zeroed by the launcher)
    auto flush_dw = [&]() {
        if (elect_one()) { tc_fence_after(); mma_commit(&gs.bar); }
        __syncwarp();
        wait_mma();
        float v[8];
        auto add2 = [&](float* dst, float x, float y) {
            float2 o = *reinterpret_cast<float2*>(dst);
            o.x += x; o.y += y;
            *reinterpret_cast<float2*>(dst) = o;
        };
        for (int cb = 0; cb < 4; cb++) {
            tmem_ld_16x256b_x2(tm_dw + lane_base + TM_DW2 + 16 * cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i += 2) add2(sl + oW2 + frag_row(i) * TH + 16 * cb + frag_col(i), v[i], v[i + 1]);
            tmem_ld_16x256b_x2(tm_dw + lane_base + TM_DW3 + 16 * cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i += 2) add2(sl + oW3 + frag_row(i) * TH + 16 * cb + frag_col(i), v[i], v[i + 1]);
        }
        tmem_ld_16x256b_x2(tm_dw + lane_base + TM_DW4T, v);          // rows = hidden k, columns = output m
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; i++) sl[oW4 + frag_col(i) * TH + frag_row(i)] += v[i];
        for (int cb = 0; cb < 2; cb++) {
            tmem_ld_16x256b_x2(tm_dw + lane_base + TM_DW1F + 16 * cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const int c = 16 * cb + frag_col(i);
                if (c < TK1) add2(garea + frag_row(i) * TK1 + c, v[i], v[i + 1]);
            }
        }
    };

    if (live) {
        bool fresh = true;
        const int bown = b0 + sn, bbown = min(bown, B - 1);
        const bool valid = bown < B;
        float lam[2], D1[8], dB2[8], dB3[8], dB4[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 8; i++) { D1[i] = 0.f; dB2[i] = 0.f; dB3[i] = 0.f; }
#pragma unroll
        for (int r = 0; r < 2; r++) lam[r] = (valid && q.gx.p) ? ldser(q.gx, T - 1, bown, sm0 + 8 * r) : 0.0f;

        const float c13 = (float)(1.0 / 3.0);
        const float* tp = q.tape + ((int64_t)gid * (T - 1) * NST + (int64_t)(T - 1) * NST - 1) * PSN_TAPE_STAGE;   // last record
        float a1[8], a2[8], a3[8], yv[2] = {0.f, 0.f};
        float dtn = 0.0f, un[TU];
#pragma unroll
        for (int c = 0; c < TU; c++) un[c] = 0.0f;
        if (T > 1) {
            ld_frag(tp, a1); ld_frag(tp + PSN_TAPE_FRAG, a2); ld_frag(tp + 2 * PSN_TAPE_FRAG, a3);
            const float2 y2 = __ldcs(reinterpret_cast<const float2*>(tp + 3 * PSN_TAPE_FRAG + gt * 2));
            yv[0] = y2.x; yv[1] = y2.y;
            dtn = __fsub_rn(ldser(q.t, T - 1, bbown, 0), ldser(q.t, T - 2, bbown, 0));
            if (warp == 1) load_held(T - 1, un);
        }

        for (int j = T - 1; j >= 1; j--) {
            const float dt = dtn;
            float u[TU];
#pragma unroll
            for (int c = 0; c < TU; c++) u[c] = un[c];
            float gxn[2];
#pragma unroll
            for (int r = 0; r < 2; r++) gxn[r] = (valid && q.gx.p) ? ldser(q.gx, j - 1, bown, sm0 + 8 * r) : 0.0f;
            if (j > 1) {
                dtn = __fsub_rn(ldser(q.t, j - 1, bbown, 0), ldser(q.t, j - 2, bbown, 0));
                if (warp == 1) load_held(j - 1, un);
            }
            // x_j = x_{j-1} + dt * sum_s bw[s] k_s : dL/dk_s starts at lam * dt * bw[s]; dxs collects dL/dx_{j-1}
            float dxs[2], d1[2] = {0.f, 0.f}, d2[2] = {0.f, 0.f}, d3[2] = {0.f, 0.f}, dcur[2];
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const float ld = lam[r] * dt;
                dxs[r] = lam[r];
                if (METHOD == PSNODE_RK4) { d1[r] = ld * 0.125f; d2[r] = ld * 0.375f; d3[r] = ld * 0.375f; dcur[r] = ld * 0.125f; }
                else dcur[r] = ld;
            }
#pragma unroll 1
            for (int e = NST - 1; e >= 0; e--) {
                const bool has_next = !(j == 1 && e == 0);
                const float* tpn = tp - PSN_TAPE_STAGE;
                float gsum[8];
                // ---- P0: dk tiles, a3 -> aB[0]; g3 = W4^T dk ; dW4^T += a3 dk^T ----
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    float hi, lo;
                    split_tf32_fast(dcur[r], hi, lo);
                    st_f32(gs.dT_hi, off_dk[r], hi); st_f32(gs.dT_lo, off_dk[r], lo);
                    st_f32(gs.dkB_hi, off_sq[r], hi); st_f32(gs.dkB_lo, off_sq[r], lo);
                    dB4[r] += dcur[r];
                }
                store_pairs(gs.aB_hi[0], gs.aB_lo[0], a3);
                publish();
                issue_m1();
                if (warp == 0) issue_dw(d_aB_hi0, d_aB_lo0, d_dkB_hi, d_dkB_lo, tm_dw + TM_DW4T, idesc16, fresh);
                // ---- P1: d3 ; a2 -> aB[1]; g2 = W3^T d3 ; dW3 += d3 a2^T ----
                collect(gsum);
                make_delta(gsum, a3, dB3, gs.dA_hi[1], gs.dA_lo[1]);
                store_pairs(gs.aB_hi[1], gs.aB_lo[1], a2);
                if (has_next) ld_frag(tpn + 2 * PSN_TAPE_FRAG, a3);
                publish();
                issue_data(TM_W3T);
                if (warp == 1) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW3, idesc64, fresh);
                // ---- P2: d2 ; a1 -> aB[0]; g1 = W2^T d2 ; dW2 += d2 a1^T ----
                collect(gsum);
                make_delta(gsum, a2, dB2, gs.dA_hi[0], gs.dA_lo[0]);
                store_pairs(gs.aB_hi[0], gs.aB_lo[0], a1);
                if (has_next) ld_frag(tpn + PSN_TAPE_FRAG, a2);
                publish();
                issue_data(TM_W2T);
                if (warp == 2) issue_dw(d_dA_hi0, d_dA_lo0, d_aB_hi0, d_aB_lo0, tm_dw + TM_DW2, idesc64, fresh);
                // ---- P3: d1 ; [y; u] -> aB[1] rows 0..23; gy = (Wb+Wc)_x^T d1 ; dW1f += d1 [y;u]^T ----
                collect(gsum);
                make_delta(gsum, a1, D1, gs.dA_hi[1], gs.dA_lo[1]);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    float hi, lo;
                    split_tf32_fast(yv[r], hi, lo);
                    st_f32(gs.aB_hi[1], off_sq[r], hi); st_f32(gs.aB_lo[1], off_sq[r], lo);
                }
                if (warp == 1 && lane < TN) {
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
                    const float2 y2 = __ldcs(reinterpret_cast<const float2*>(tpn + 3 * PSN_TAPE_FRAG + gt * 2));
                    yv[0] = y2.x; yv[1] = y2.y;
                }
                publish();
                issue_data(TM_W1T);
                if (warp == 3) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW1F, idesc24, fresh);
                // ---- P4: dL/dy of this stage -> Runge-Kutta adjoint algebra (my_fixed_grid.py:15-59 reversed) ----
                float gy[2];
                collect_gy(gy);
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const float gq = gy[r];
                    dxs[r] += gq;
                    if (METHOD == PSNODE_RK4) {
                        const float tg = dt * gq;
                        if (e == 3) { d3[r] += tg; d2[r] -= tg; d1[r] += tg; dcur[r] = d3[r]; }
                        else if (e == 2) { d2[r] += tg; d1[r] -= tg * c13; dcur[r] = d2[r]; }
                        else if (e == 1) { d1[r] += tg * c13; dcur[r] = d1[r]; }
                    } else if (METHOD == PSNODE_MIDPOINT) {
                        if (e == 1) dcur[r] = (0.5f * dt) * gq;
                    }
                }
                tp = tpn;
                fresh = false;
            }
#pragma unroll
            for (int r = 0; r < 2; r++) lam[r] = dxs[r] + gxn[r];
            if (((T - j) % PSN_DW_FLUSH) == 0 || j == 1) { flush_dw(); fresh = true; }
        }
        if (q.d_x0 && valid) {
#pragma unroll
            for (int r = 0; r < 2; r++) q.d_x0[(int64_t)bown * q.d_x0_sb + sm0 + 8 * r] = lam[r];
        }

        // ---- the weight gradients are in the slab already (last flush); bias gradients and layer-1 unfolding follow ----
        group_sync(g);
        float* scr = reinterpret_cast<float*>(gs.dA_hi[0]);          // 32 KB of dead tiles: scratch
        float* D1s = scr;                 // [64][17]
        float* Gs = D1s + TH * 17;        // [64][25]
        float* a0s = Gs + TH * 25;        // [16][25]
        float* red = a0s + TN * 25;       // [4][16]
        for (int e = gt; e < TH * TK1; e += GROUP_THREADS) Gs[(e / TK1) * 25 + (e % TK1)] = 0.0f;
        group_sync(g);
#pragma unroll
        for (int cb = 0; cb < 2; cb++)                               // read back this thread's own flushed elements
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int c = 16 * cb + frag_col(i);
                if (c < TK1) Gs[frag_row(i) * 25 + c] = garea[frag_row(i) * TK1 + c];
            }
        // bias gradients: sum over the 4 trajectory columns of the fragment, then over the 4 lanes that share a row
        auto row_sums = [&](const float (&v)[8], int off) {
            float s0 = (v[0] + v[1]) + (v[4] + v[5]), s1 = (v[2] + v[3]) + (v[6] + v[7]);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            if ((lane & 3) == 0) { sl[off + m0] = s0; sl[off + m0 + 8] = s1; }
        };
        row_sums(D1, ob1);
        row_sums(dB2, ob2);
        row_sums(dB3, ob3);
        {
            float s0 = dB4[0], s1 = dB4[1];
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            if ((lane & 3) == 0) { red[warp * 16 + sm0] = s0; red[warp * 16 + sm0 + 8] = s1; }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) D1s[frag_row(i) * 17 + frag_col(i)] = D1[i];
        for (int e = gt; e < TN * S; e += GROUP_THREADS) {
            const int n = e / S, c = e - n * S;
            a0s[n * 25 + c] = __ldg(q.a0 + (int64_t)min(b0 + n, B - 1) * q.a0_sb + c);
        }
        group_sync(g);
        if (gt < TX) sl[ob4 + gt] = (red[gt] + red[16 + gt]) + (red[32 + gt] + red[48 + gt]);
        // unfold layer 1: W1 = [Wa | Wb | Wc] acting on [a0; s - a0; s]
        for (int e = gt; e < TH * S; e += GROUP_THREADS) {
            const int m = e / S, c = e - m * S;
            float P = 0.0f;
#pragma unroll
            for (int n = 0; n < TN; n++) P = fmaf(D1s[m * 17 + n], a0s[n * 25 + c], P);
            const float G = Gs[m * 25 + c];
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
    return 256 + (int64_t)psn_tc_ngroups(p->B) * (n_theta + G_AREA) * 4;
}

int psn_tc_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < psn_tc_backward_workspace(p, a)) return PSNODE_EWORKSPACE;
    const int64_t n_theta = psnode_mlp_param_count(&p->de);
    if (a->n_theta != n_theta) return PSNODE_EINVAL;
    TcBwdParams q;
    q.B = p->B; q.T = p->T; q.Z = p->Z; q.S = p->X + p->Z;
    q.n_theta = (int)n_theta;
    q.t = p->t; q.z = p->z; q.gx = a->gx;
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
    PSN_CUDA(cudaMemsetAsync(ws, 0, 256 + (size_t)ngroups * (n_theta + G_AREA) * 4, stream));
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
        case PSNODE_EULER: st = launch(psn_tc_bwd_kernel<PSNODE_EULER>, "psn_tc_bwd_kernel<euler>"); break;
        case PSNODE_MIDPOINT: st = launch(psn_tc_bwd_kernel<PSNODE_MIDPOINT>, "psn_tc_bwd_kernel<midpoint>"); break;
        default: st = launch(psn_tc_bwd_kernel<PSNODE_RK4>, "psn_tc_bwd_kernel<rk4>"); break;
    }
    if (st != PSNODE_OK) return st;
    psn_tc_grad_reduce_kernel<<<32, 256, 0, stream>>>(q.slab, ngroups, (int)n_theta, (int)n_theta + G_AREA, a->d_theta);
    psn_count_launch("psn_tc_grad_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
