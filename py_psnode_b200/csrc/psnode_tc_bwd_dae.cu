// psnode_tc_bwd_dae.cu -- tensor-core reverse sweep (discrete adjoint) for the reference's H = 64 DAE nets (BASELINE configs[2]).
//
// Exact reverse mode of FixedGridODESolver.integrate_DAE (neural_dae/my_solvers.py:82-131) with DE_Func + AE_Func of
// neural_01_DAE_01_no_encode.py:61-83 -- what the reference obtains from autograd at loss.backward() (:423).  Same machinery
// as psnode_tc_bwd.cu (activation tape written by the forward kernel, 3xTF32 data path g = W^T delta, weight gradients as
// MMA chains with K = the 16 trajectories accumulated in TMEM, truncation-safe periodic flush), extended by the algebraic net:
//
//   point j : mu_j += gi[j];  AE chain with dk := mu_j at (x_j, z[j], v[j])         -> lam_j += dL/dx (x part of the AE input)
//   step  j : NST DE stages (as the ODE sweep) + du_i = sum over stages of (Wb+Wc)_i^T d1  (i is zero-order held, :104-119)
//             no event : mu_{j-1} = du_i + gi[j-1]
//             event k  : the step used i_0 = ae(x_{j-1}, z_jump[k], v_jump[k]) (:108-110): AE chain with dk := du_i at the
//                        recorded event evaluation -> lam_{j-1} += dL/dx;  mu_{j-1} = gi[j-1]
//   The AE chain is the SAME four-phase chain as a DE stage (shapes 42->64->64->64->I vs 69->64->64->64->16 after folding
//   layer 1 onto the [x | z v i] tile), so one stage routine serves both nets.
//
// Resources.  Two weight sets do not fit the TMEM plan of the ODE sweep with two groups per CTA, so this kernel runs ONE
// 16-trajectory group (8 warps, 256 threads, up to 255 registers) per CTA:
//   TMEM lanes 0..15 : 4 partial accumulators (64) | W4^T (32) | A4^T (32) | W3^T | W2^T | (Wb+Wc)_x^T (128 each)  -> TS MMAs
//   TMEM lanes 16..31: dW accumulators of the DE net (176) | of the AE net (176) | accumulators of the du_i chain (64)
//   shared memory    : A3^T, A2^T, (A1_x)^T and (Wb+Wc)_i^T as A operands of SS MMAs (128 KB), the group's tiles (44 KB)
#include <cstddef>
#include <type_traits>
#include "psnode_internal.cuh"
#include "psnode_tc.cuh"
#include "psnode_tc_tape.cuh"

namespace {
using namespace psn_tc;

constexpr int TN = PSN_TC_TN;
constexpr int TH = 64, TX = 16, TU = 8;
constexpr int TK1 = TX + TU;
constexpr int LBO = 144;
constexpr int SBO_ACT = (TH / 4) * LBO;
constexpr int ACT_TILE = (TN / 8) * SBO_ACT;
constexpr int LBO_W = 128;
constexpr int SBO_W64 = (TH / 4) * LBO_W, W64_TILE = (TH / 8) * SBO_W64;     // 64 rows x K = 64   (16 KB)
constexpr int SBO_K16 = (TN / 4) * LBO_W, K16_TILE = (TH / 8) * SBO_K16;     // 64 rows x K = 16   ( 4 KB)
constexpr int DKB_TILE = (TX / 8) * SBO_K16;
// TMEM, lanes 0..15
constexpr int TM_ACC = 0, TM_W4T = 64, TM_A4T = 96, TM_W3T = 128, TM_W2T = 256, TM_W1T = 384;
// TMEM, lanes 16..31
constexpr uint32_t TM_UPPER = 16u << 16;
constexpr int TM_DW2 = 0, TM_DW3 = 64, TM_DW4T = 128, TM_DW1F = 144, TM_DWNET = 176;   // per net
constexpr int TM_ACC_M5 = 352;
constexpr int TM_COLS = 512;
constexpr int GROUP_THREADS = 256;
constexpr int PSN_DW_CHAIN = 16;       // accumulations into a TMEM weight-gradient accumulator between two flushes (16 / NST steps)
constexpr int G_AREA = TH * TK1;
// slab = [DE parameters | pad to 4 | AE parameters | pad to 4 | two folded-layer-1 areas]: 16-byte aligned for odd X / I
__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int slab_floats(int n_de, int n_ae) { return pad4(n_de) + pad4(n_ae) + 2 * G_AREA; }

struct DaeBwdParams {
    int B, T, X, Z, V, I, S, E, n_theta, n_theta_de;       // X <= 16 state variables (tile rows / columns X..15 are zero padding)
    psnode_series t, z, v, gx, gi;
    PsnFuse fx, fi;                             // fused masked-MSE upstream gradients (replace gx / gi when their target is set)
    const float* x_sol; int64_t xs_st, xs_sb;
    const float* i_sol; int64_t is_st, is_sb;
    const float* a0; int64_t a0_sb;
    const int32_t* event_idx;
    const float* z_jump; int64_t zj_sb, zj_se;
    const float* v_jump; int64_t vj_sb, vj_se;
    const float* W1; const float* W2; const float* W3; const float* W4;
    const float* A1; const float* A2; const float* A3; const float* A4;
    const float* tape;
    float* slab;
    float* d_x0; int64_t d_x0_sb;
    float* d_a0; int64_t d_a0_sb;
    int* err;
};

struct __align__(128) DaeBwdSmem {
    unsigned char a3t_hi[W64_TILE], a3t_lo[W64_TILE];     // A[k][m] = A3[m][k]
    unsigned char a2t_hi[W64_TILE], a2t_lo[W64_TILE];
    unsigned char a1t_hi[W64_TILE], a1t_lo[W64_TILE];     // A[r][m] = A1[m][S + (r & 15)]      (x columns of the AE input, replicated)
    unsigned char wit_hi[W64_TILE], wit_lo[W64_TILE];     // A[r][m] = (Wb+Wc)[m][X+Z+V + (r & 15)], zero for (r & 15) >= I
    unsigned char dT_hi[ACT_TILE], dT_lo[ACT_TILE];
    unsigned char dA_hi[2][K16_TILE], dA_lo[2][K16_TILE];
    unsigned char aB_hi[2][K16_TILE], aB_lo[2][K16_TILE];
    unsigned char dkB_hi[DKB_TILE], dkB_lo[DKB_TILE];
    uint64_t bar;
    uint32_t tmem_base;
};

__device__ __forceinline__ float ldser(const psnode_series& s, int j, int b, int c) {
    return __ldg(s.p + (int64_t)j * s.st + (int64_t)b * s.sb + c);
}
__device__ __forceinline__ void group_sync() { asm volatile("bar.sync 1, %0;" ::"r"(GROUP_THREADS) : "memory"); }
__device__ __forceinline__ void st_f32(unsigned char* base, int off, float v) { *reinterpret_cast<float*>(base + off) = v; }
__device__ __forceinline__ void st_f32x2(unsigned char* base, int off, float a, float b) { *reinterpret_cast<float2*>(base + off) = make_float2(a, b); }

using DE_NET = std::integral_constant<bool, false>;
using AE_NET = std::integral_constant<bool, true>;

// FUSED: upstream gradients formed from the stored trajectories (psnode_adjoint.fuse_x / fuse_i); separate instantiation so the
// plain gx / gi kernel keeps its register allocation
template <int METHOD, bool FUSED>
__global__ void __launch_bounds__(GROUP_THREADS, 1) psn_tc_bwd_dae_kernel(const __grid_constant__ DaeBwdParams q) {
    constexpr int NST = METHOD == PSNODE_EULER ? 1 : (METHOD == PSNODE_MIDPOINT ? 2 : 4);
    extern __shared__ unsigned char smem_raw[];
    DaeBwdSmem& gs = *reinterpret_cast<DaeBwdSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x;
    const int wk = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int gt = tid;
    const int wq = wk & 3, h = wk >> 2;
    const bool issuer = h == 0;
    const int B = q.B, T = q.T, X = q.X, Z = q.Z, V = q.V, I = q.I, S = q.S, K1 = 3 * q.S;
    const int ZV = Z + V, KA = S + X + ZV;
    const int gid = blockIdx.x;
    const int b0 = gid * TN;

    // ---- one-time setup -------------------------------------------------------------------------------
    if (tid == 0) { mbar_init(&gs.bar, 4); fence_mbar_init(); }
    if (wk == 0) tmem_alloc(&gs.tmem_base, TM_COLS);
    for (int e = tid; e < TH * TH; e += GROUP_THREADS) {       // shared-memory A operands (transposed / replicated)
        const int r = e >> 6, m = e & 63;                      // tile row r, K index m
        float hi, lo;
        const int o = tile_byte(r, m, LBO_W, SBO_W64);
        split_tf32(__ldg(q.A3 + m * TH + r), hi, lo); st_f32(gs.a3t_hi, o, hi); st_f32(gs.a3t_lo, o, lo);
        split_tf32(__ldg(q.A2 + m * TH + r), hi, lo); st_f32(gs.a2t_hi, o, hi); st_f32(gs.a2t_lo, o, lo);
        hi = 0.0f; lo = 0.0f;
        if ((r & 15) < X) split_tf32(__ldg(q.A1 + m * KA + S + (r & 15)), hi, lo);
        st_f32(gs.a1t_hi, o, hi); st_f32(gs.a1t_lo, o, lo);
        hi = 0.0f; lo = 0.0f;
        if ((r & 15) < I) {
            const int c = X + ZV + (r & 15);            // index of i[r & 15] in s = cat(x, z, v, i)
            split_tf32(__ldg(q.W1 + m * K1 + S + c) + __ldg(q.W1 + m * K1 + 2 * S + c), hi, lo);
        }
        st_f32(gs.wit_hi, o, hi); st_f32(gs.wit_lo, o, lo);
    }
    for (int e = gt; e < (int)((offsetof(DaeBwdSmem, bar) - offsetof(DaeBwdSmem, dT_hi)) / 4); e += GROUP_THREADS)
        reinterpret_cast<float*>(gs.dT_hi)[e] = 0.0f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = gs.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    // resident operands -> TMEM lanes 0..15 (warps 0..3 write):  A[k][m] = W3[m][k], W2[m][k];  A[r][m] = (Wb+Wc)[m][r & 15];
    // A[k][m] = W4[m][k] (m < 16);  A[k][m] = A4[m][k] (m < I, else 0)
    if (issuer) {
        const int r0 = 16 * wq + (lane >> 2), cc0 = 2 * (lane & 3);
        for (int half = 0; half < 2; half++) {
            for (int cb = 0; cb < 4; cb++) {
                float w3[8], w2[8], w1[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int row = r0 + ((i >> 1) & 1) * 8, col = 16 * cb + cc0 + (i & 1) + (i >> 2) * 8;
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
            float w4[8], a4[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int row = r0 + ((i >> 1) & 1) * 8, col = cc0 + (i & 1) + (i >> 2) * 8;
                float hi, lo;
                hi = 0.0f; lo = 0.0f;
                if (col < X) split_tf32(__ldg(q.W4 + col * TH + row), hi, lo);
                w4[i] = half ? lo : hi;
                hi = 0.0f; lo = 0.0f;
                if (col < I) split_tf32(__ldg(q.A4 + col * TH + row), hi, lo);
                a4[i] = half ? lo : hi;
            }
            tmem_st_16x256b_x2(tmem + lane_base + TM_W4T + 16 * half, w4);
            tmem_st_16x256b_x2(tmem + lane_base + TM_A4T + 16 * half, a4);
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
    int off_act[4];
#pragma unroll
    for (int i = 0; i < 4; i++) off_act[i] = tile_byte(frag_col(i), frag_row(i), LBO, SBO_ACT);
    const int off2[2] = {(int)tile_byte(m0, 8 * h + c0, LBO_W, SBO_K16), (int)tile_byte(m0 + 8, 8 * h + c0, LBO_W, SBO_K16)};
    // the element this thread owns in the stage algebra: row srow (state; algebraic variable if srow < I) of trajectory sn
    const int srow = (lane >> 2) + 8 * h, sn = c0 + (wq & 1) + 8 * (wq >> 1);
    const int off_dk = (int)tile_byte(sn, srow, LBO, SBO_ACT);
    const int off_sq = (int)tile_byte(srow, sn, LBO_W, SBO_K16);
    const int ftape = (32 * wq + lane) * 8 + 4 * h;
    const int ytape = 3 * PSN_TAPE_FRAG + (32 * wq + lane) * 2 + h;
    const bool own_i = srow < I;

    const uint32_t idesc16 = make_idesc_tf32(TH, TN), idesc24 = make_idesc_tf32(TH, TK1), idesc64 = make_idesc_tf32(TH, TH);
    // three base descriptors, one per (LBO, SBO) pair; every other tile sits at a constant byte offset from its base, so its
    // descriptor is base + (offset >> 4): ptxas keeps 3 descriptors in uniform registers instead of rebuilding 20
    const uint64_t d_dT_hi = make_desc(smem_u32(gs.dT_hi), LBO, SBO_ACT);
    const uint64_t d_w64 = make_desc(smem_u32(gs.a3t_hi), LBO_W, SBO_W64);
    const uint64_t d_k16 = make_desc(smem_u32(gs.dA_hi[0]), LBO_W, SBO_K16);
    auto at = [](uint64_t base, size_t from, size_t to) { return base + (uint64_t)((to - from) >> 4); };
    constexpr size_t o_w = offsetof(DaeBwdSmem, a3t_hi), o_dA = offsetof(DaeBwdSmem, dA_hi);
    const uint64_t d_dT_lo = at(d_dT_hi, offsetof(DaeBwdSmem, dT_hi), offsetof(DaeBwdSmem, dT_lo));
    const uint64_t d_a3t_hi = d_w64, d_a3t_lo = at(d_w64, o_w, offsetof(DaeBwdSmem, a3t_lo));
    const uint64_t d_a2t_hi = at(d_w64, o_w, offsetof(DaeBwdSmem, a2t_hi)), d_a2t_lo = at(d_w64, o_w, offsetof(DaeBwdSmem, a2t_lo));
    const uint64_t d_a1t_hi = at(d_w64, o_w, offsetof(DaeBwdSmem, a1t_hi)), d_a1t_lo = at(d_w64, o_w, offsetof(DaeBwdSmem, a1t_lo));
    const uint64_t d_wit_hi = at(d_w64, o_w, offsetof(DaeBwdSmem, wit_hi)), d_wit_lo = at(d_w64, o_w, offsetof(DaeBwdSmem, wit_lo));
    const uint64_t d_dA_hi0 = d_k16, d_dA_hi1 = at(d_k16, 0, K16_TILE);
    const uint64_t d_dA_lo0 = at(d_k16, o_dA, offsetof(DaeBwdSmem, dA_lo)), d_dA_lo1 = at(d_k16, o_dA, offsetof(DaeBwdSmem, dA_lo) + K16_TILE);
    const uint64_t d_aB_hi0 = at(d_k16, o_dA, offsetof(DaeBwdSmem, aB_hi)), d_aB_hi1 = at(d_k16, o_dA, offsetof(DaeBwdSmem, aB_hi) + K16_TILE);
    const uint64_t d_aB_lo0 = at(d_k16, o_dA, offsetof(DaeBwdSmem, aB_lo)), d_aB_lo1 = at(d_k16, o_dA, offsetof(DaeBwdSmem, aB_lo) + K16_TILE);
    const uint64_t d_dkB_hi = at(d_k16, o_dA, offsetof(DaeBwdSmem, dkB_hi)), d_dkB_lo = at(d_k16, o_dA, offsetof(DaeBwdSmem, dkB_lo));
    const uint32_t acc_base = tmem + TM_ACC;
    const uint32_t my_acc = acc_base + (uint32_t)wq * TN;
    const uint32_t acc_m5 = tmem + TM_UPPER + TM_ACC_M5;
    constexpr uint64_t KSTEP_B = (uint64_t)((2 * LBO) >> 4), KSTEP_W = (uint64_t)((2 * LBO_W) >> 4);
    uint32_t phase = 0;

    // ---- MMA helpers -----------------------------------------------------------------------------------
    // K = 64 data MMA, A resident in TMEM (column w_tm, lo at +64): issuing warp wq takes K-steps 2wq, 2wq+1 of the 3 terms
    auto mma_ts64 = [&](uint32_t w_tm) {
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
    };
    // same with the A operand in shared memory, into accumulator set `acc` (+ wq * 16)
    auto mma_ss64 = [&](uint64_t a_hi, uint64_t a_lo, uint32_t acc) {
        uint32_t accumulate = 0;
        for (int term = 0; term < 3; term++) {
            const uint64_t ad = term == 0 ? a_lo : a_hi;
            const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
            for (int kk = 0; kk < 2; kk++) {
                const int ks = 2 * wq + kk;
                mma_tf32(acc + (uint32_t)wq * TN, ad + KSTEP_W * ks, bd + KSTEP_B * ks, idesc16, accumulate);
                accumulate = 1;
            }
        }
    };
    // g3 = W4^T dk (K = 16, A in TMEM column w_tm, lo at +16): six MMAs, entry e = 2*term + kstep, warp wq takes e = wq, wq + 4
    auto mma_m1 = [&](uint32_t w_tm) {
        uint32_t accumulate = 0;
        for (int e = wq; e < 6; e += 4) {
            const int term = e >> 1, ks = e & 1;
            const uint32_t wa = w_tm + (term == 0 ? 16 : 0);
            const uint64_t bd = term == 1 ? d_dT_lo : d_dT_hi;
            mma_tf32_ts(my_acc, tmem + wa + 8 * ks, bd + KSTEP_B * ks, idesc16, accumulate);
            accumulate = 1;
        }
    };
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
    auto collect = [&](float (&d)[4]) {
        wait_mma();
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc_base + lane_base + 8 * h;
        tmem_ld_16x256b_x1(a + 0 * TN, t0);
        tmem_ld_16x256b_x1(a + 1 * TN, t1);
        tmem_ld_16x256b_x1(a + 2 * TN, t2);
        tmem_ld_16x256b_x1(a + 3 * TN, t3);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = (t0[i] + t1[i]) + (t2[i] + t3[i]);
    };
    // element (row srow, trajectory sn) of a replicated 16 x 16 result tile (every 16-row block holds the same tile)
    auto own_element = [&](uint32_t acc) {
        float t0[4], t1[4], t2[4], t3[4];
        const uint32_t a = acc + lane_base + 8 * (wq >> 1);
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
        group_sync();
    };
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
    auto event_of_step = [&](int j) { return q.event_idx ? __ldg(q.event_idx + (j - 1)) : -1; };
    // z / v part of a layer-1 input tile (warp 4, lane = trajectory): grid point jp, or event k
    auto load_zv = [&](int jp, int k, float (&u)[TU]) {
        const int bb = min(b0 + (lane & 15), B - 1);
#pragma unroll
        for (int c = 0; c < TU; c++) {
            u[c] = 0.0f;
            if (c < Z) u[c] = k >= 0 ? __ldg(q.z_jump + (int64_t)bb * q.zj_sb + (int64_t)k * q.zj_se + c) : ldser(q.z, jp, bb, c);
            else if (c < ZV) u[c] = k >= 0 ? __ldg(q.v_jump + (int64_t)bb * q.vj_sb + (int64_t)k * q.vj_se + (c - Z)) : ldser(q.v, jp, bb, c - Z);
        }
    };

    // ---- slab layout: the reference's parameter order (DE net, then AE net), then the two folded-layer-1 areas ------------
    float* sl = q.slab + (int64_t)gid * slab_floats(q.n_theta_de, q.n_theta - q.n_theta_de);
    float* garea_de = sl + pad4(q.n_theta_de) + pad4(q.n_theta - q.n_theta_de);
    float* garea_ae = garea_de + G_AREA;
    const int oW1 = 0, ob1 = TH * K1, oW2 = ob1 + TH, ob2 = oW2 + TH * TH, oW3 = ob2 + TH, ob3 = oW3 + TH * TH, oW4 = ob3 + TH,
              ob4 = oW4 + X * TH;
    const int oA1 = pad4(q.n_theta_de), oab1 = oA1 + TH * KA, oA2 = oab1 + TH, oab2 = oA2 + TH * TH, oA3 = oab2 + TH, oab3 = oA3 + TH * TH,
              oA4 = oab3 + TH, oab4 = oA4 + I * TH;

    // drain the tensor pipe and add one net's TMEM weight-gradient accumulators into the slab (round-to-nearest adds)
    auto flush_net = [&](uint32_t tm_dw, int o2, int o3, int o4, int n4, float* garea) {
        // the slab values are fetched first, all at once (independent L2 round trips), then combined with the accumulators
        const int r0 = m0, cc = c0;
        float2 p2[4][2], p3[4][2], pg[2][2];
        float p4[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {          // dW2 / dW3: 64 columns = 8 blocks of 8, this warp takes blocks h, h+2, h+4, h+6
            const int cb = h + 2 * q4;
#pragma unroll
            for (int r = 0; r < 2; r++) {
                p2[q4][r] = *reinterpret_cast<const float2*>(sl + o2 + (r0 + 8 * r) * TH + 8 * cb + cc);
                p3[q4][r] = *reinterpret_cast<const float2*>(sl + o3 + (r0 + 8 * r) * TH + 8 * cb + cc);
            }
        }
        const int mc = 8 * h + cc;                // layer 4, transposed: rows = hidden k, columns = output m (n4 real)
        const bool ok0 = mc < n4, ok1 = mc + 1 < n4;
        p4[0] = ok0 ? sl[o4 + mc * TH + r0] : 0.0f; p4[1] = ok1 ? sl[o4 + (mc + 1) * TH + r0] : 0.0f;
        p4[2] = ok0 ? sl[o4 + mc * TH + r0 + 8] : 0.0f; p4[3] = ok1 ? sl[o4 + (mc + 1) * TH + r0 + 8] : 0.0f;
#pragma unroll
        for (int q2 = 0; q2 < 2; q2++) {          // folded layer 1: 24 columns = 3 blocks of 8: blocks h and h + 2 (< 3)
            const int cb = h + 2 * q2;
#pragma unroll
            for (int r = 0; r < 2; r++)
                pg[q2][r] = cb < 3 ? *reinterpret_cast<const float2*>(garea + (r0 + 8 * r) * TK1 + 8 * cb + cc) : make_float2(0.f, 0.f);
        }
        float v[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; q4++) {
            const int cb = h + 2 * q4;
            tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW2 + 8 * cb, v);
            tmem_ld_wait();
            *reinterpret_cast<float2*>(sl + o2 + r0 * TH + 8 * cb + cc) = make_float2(p2[q4][0].x + v[0], p2[q4][0].y + v[1]);
            *reinterpret_cast<float2*>(sl + o2 + (r0 + 8) * TH + 8 * cb + cc) = make_float2(p2[q4][1].x + v[2], p2[q4][1].y + v[3]);
            tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW3 + 8 * cb, v);
            tmem_ld_wait();
            *reinterpret_cast<float2*>(sl + o3 + r0 * TH + 8 * cb + cc) = make_float2(p3[q4][0].x + v[0], p3[q4][0].y + v[1]);
            *reinterpret_cast<float2*>(sl + o3 + (r0 + 8) * TH + 8 * cb + cc) = make_float2(p3[q4][1].x + v[2], p3[q4][1].y + v[3]);
        }
        tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW4T + 8 * h, v);
        tmem_ld_wait();
        if (ok0) { sl[o4 + mc * TH + r0] = p4[0] + v[0]; sl[o4 + mc * TH + r0 + 8] = p4[2] + v[2]; }
        if (ok1) { sl[o4 + (mc + 1) * TH + r0] = p4[1] + v[1]; sl[o4 + (mc + 1) * TH + r0 + 8] = p4[3] + v[3]; }
#pragma unroll
        for (int q2 = 0; q2 < 2; q2++) {
            const int cb = h + 2 * q2;
            if (cb < 3) {                          // warp-uniform
                tmem_ld_16x256b_x1(tm_dw + lane_base + TM_DW1F + 8 * cb, v);
                tmem_ld_wait();
                *reinterpret_cast<float2*>(garea + r0 * TK1 + 8 * cb + cc) = make_float2(pg[q2][0].x + v[0], pg[q2][0].y + v[1]);
                *reinterpret_cast<float2*>(garea + (r0 + 8) * TK1 + 8 * cb + cc) = make_float2(pg[q2][1].x + v[2], pg[q2][1].y + v[3]);
            }
        }
    };
    auto flush_dw = [&](bool de_dirty, bool ae_dirty) {      // a net's accumulators hold data only after one of its chains ran
        if (issuer) {
            if (elect_one()) { tc_fence_after(); mma_commit(&gs.bar); }
            __syncwarp();
        }
        wait_mma();
        if (de_dirty) flush_net(tmem + TM_UPPER, oW2, oW3, oW4, X, garea_de);
        if (ae_dirty) flush_net(tmem + TM_UPPER + TM_DWNET, oA2, oA3, oA4, I, garea_ae);
    };

    if (b0 < B) {
        const int bown = b0 + sn, bbown = min(bown, B - 1);
        const bool valid = bown < B;
        const bool own_x_ok = srow < X;
        // register accumulators: sums of delta_1 per (neuron, trajectory) and bias gradients of layers 2..4, per net
        float D1[4], dB2[4], dB3[4], dB4 = 0.0f, D1a[4], dB2a[4], dB3a[4], dB4a = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) { D1[i] = 0.f; dB2[i] = 0.f; dB3[i] = 0.f; D1a[i] = 0.f; dB2a[i] = 0.f; dB3a[i] = 0.f; }
        bool fresh_de = true, fresh_ae = true;
        float a1[4], a2[4], a3[4];                 // fragments of the record that is processed next
        const float* tbase = q.tape + (int64_t)gid * psn_dae_group_recs(T, NST, q.E) * PSN_TAPE_STAGE;
        auto rec_ptr = [&](int64_t idx) { return tbase + idx * PSN_TAPE_STAGE; };
        auto load_rec = [&](const float* r) { ld_frag(r, a1); ld_frag(r + PSN_TAPE_FRAG, a2); ld_frag(r + 2 * PSN_TAPE_FRAG, a3); };

        // One chain = the four dependent phases of a net evaluation in reverse.  NET = DE_NET (a Runge-Kutta stage: data MMAs
        // from TMEM, plus the du_i chain) or AE_NET (data MMAs from shared memory).  a1..a3 hold this evaluation's activations;
        // rec_next (or nullptr) is prefetched into them phase by phase.  Returns dL/d(x part of the input) of element (srow, sn).
        auto chain = [&](auto NET, float dkv, float xin, const float (&u)[TU], const float* rec_next, bool fresh,
                         float (&bD1)[4], float (&bB2)[4], float (&bB3)[4], float& bB4, float& dui_out) {
            constexpr bool AE = decltype(NET)::value;
            const uint32_t tm_dw = tmem + TM_UPPER + (AE ? TM_DWNET : 0);
            float gsum[4];
            // ---- P0: dk tiles, a3 -> aB[0]; g3 = W4^T dk ; dW4^T += a3 dk^T ----
            {
                float hi, lo;
                split_tf32_fast(dkv, hi, lo);
                st_f32(gs.dT_hi, off_dk, hi); st_f32(gs.dT_lo, off_dk, lo);
                st_f32(gs.dkB_hi, off_sq, hi); st_f32(gs.dkB_lo, off_sq, lo);
                bB4 += dkv;
            }
            store_pairs(gs.aB_hi[0], gs.aB_lo[0], a3);
            publish();
            if (issuer) {
                if (elect_one()) { tc_fence_after(); mma_m1(AE ? TM_A4T : TM_W4T); mma_commit(&gs.bar); }
                __syncwarp();
            }
            if (wk == 0) issue_dw(d_aB_hi0, d_aB_lo0, d_dkB_hi, d_dkB_lo, tm_dw + TM_DW4T, idesc16, fresh);
            // ---- P1 ----
            collect(gsum);
            make_delta(gsum, a3, bB3, gs.dA_hi[1], gs.dA_lo[1]);
            store_pairs(gs.aB_hi[1], gs.aB_lo[1], a2);
            if (rec_next) ld_frag(rec_next + 2 * PSN_TAPE_FRAG, a3);
            publish();
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    if (AE) mma_ss64(d_a3t_hi, d_a3t_lo, acc_base); else mma_ts64(TM_W3T);
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
            if (wk == 1) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW3, idesc64, fresh);
            // ---- P2 ----
            collect(gsum);
            make_delta(gsum, a2, bB2, gs.dA_hi[0], gs.dA_lo[0]);
            store_pairs(gs.aB_hi[0], gs.aB_lo[0], a1);
            if (rec_next) ld_frag(rec_next + PSN_TAPE_FRAG, a2);
            publish();
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    if (AE) mma_ss64(d_a2t_hi, d_a2t_lo, acc_base); else mma_ts64(TM_W2T);
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
            if (wk == 2) issue_dw(d_dA_hi0, d_dA_lo0, d_aB_hi0, d_aB_lo0, tm_dw + TM_DW2, idesc64, fresh);
            // ---- P3: layer-1 input tile [x part | u part] -> aB[1] rows 0..23 ----
            collect(gsum);
            make_delta(gsum, a1, bD1, gs.dA_hi[1], gs.dA_lo[1]);
            {
                float hi, lo;
                split_tf32_fast(xin, hi, lo);
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
            if (rec_next) ld_frag(rec_next, a1);
            publish();
            if (issuer) {
                if (elect_one()) {
                    tc_fence_after();
                    if (AE) mma_ss64(d_a1t_hi, d_a1t_lo, acc_base);
                    else { mma_ts64(TM_W1T); mma_ss64(d_wit_hi, d_wit_lo, acc_m5); }     // dL/dx part and dL/di part
                    mma_commit(&gs.bar);
                }
                __syncwarp();
            }
            if (wk == 3) issue_dw(d_dA_hi1, d_dA_lo1, d_aB_hi1, d_aB_lo1, tm_dw + TM_DW1F, idesc24, fresh);
            // ---- P4 ----
            wait_mma();
            const float gq = own_element(acc_base);
            if (!AE) dui_out = own_element(acc_m5);
            return gq;
        };

        // ---- reverse sweep ---------------------------------------------------------------------------------
        auto own_x = [&](int j) { return own_x_ok ? __ldg(q.x_sol + (int64_t)j * q.xs_st + (int64_t)bbown * q.xs_sb + srow) : 0.0f; };
        // fused masked-MSE upstream gradients: nothing is kept in registers between calls (every operand is re-read from the
        // kernel parameters), so the MMA descriptors stay in uniform registers (tools/sass_r2ur_check.py)
        auto own_gx = [&](int j) -> float {
            if (!(valid && own_x_ok)) return 0.0f;
            if (FUSED && q.fx.term.target.p) return psn_fuse_grad(q.fx, psn_fuse_scale(q.fx), j, bown, srow);
            return q.gx.p ? ldser(q.gx, j, bown, srow) : 0.0f;
        };
        auto own_gi = [&](int j) -> float {
            if (!(valid && own_i)) return 0.0f;
            if (FUSED && q.fi.term.target.p) return psn_fuse_grad(q.fi, psn_fuse_scale(q.fi), j, bown, srow);
            return q.gi.p ? ldser(q.gi, j, bown, srow) : 0.0f;
        };
        float lam = own_gx(T - 1), mu = own_gi(T - 1);
        const float c13 = (float)(1.0 / 3.0);
        float dui_dummy = 0.0f;
        // held inputs [z0 v0 i0] of step j (warp 4, trajectory = lane): jumped z / v and the re-evaluated i_0 on an event step
        auto load_held = [&](int j, int k, float (&uh)[TU]) {
            load_zv(j - 1, k, uh);
            const int n = lane & 15, bb = min(b0 + n, B - 1);
#pragma unroll
            for (int c = 0; c < TU; c++) {
                if (c >= ZV && c < ZV + I) {
                    const int ci = c - ZV;
                    if (k >= 0) {   // recorded by forward thread (w, h, lane') that owns (row ci, trajectory n)
                        const int wf = (n & 1) + 2 * (n >> 3), lf = 4 * (ci & 7) + ((n & 7) >> 1), hf = ci >> 3;
                        uh[c] = __ldcs(rec_ptr(psn_dae_rec_event(k, T, NST)) + 3 * PSN_TAPE_FRAG + (32 * wf + lf) * 2 + hf);
                    } else {
                        uh[c] = __ldg(q.i_sol + (int64_t)(j - 1) * q.is_st + (int64_t)bb * q.is_sb + ci);
                    }
                }
            }
        };
        // everything a step needs from global memory is fetched one step ahead (registers *_n)
        float un_ae[TU], un_de[TU], tan = 0.0f, tbn = 0.0f;
#pragma unroll
        for (int c = 0; c < TU; c++) { un_ae[c] = 0.0f; un_de[c] = 0.0f; }
        if (wk == 4) load_zv(T - 1, -1, un_ae);
        if (T > 1) {
            tan = ldser(q.t, T - 1, bbown, 0); tbn = ldser(q.t, T - 2, bbown, 0);
            if (wk == 4) load_held(T - 1, event_of_step(T - 1), un_de);
        }
        load_rec(rec_ptr(psn_dae_rec_point(T - 1, T, NST)));
        for (int j = T - 1; j >= 1; j--) {
            const int k = event_of_step(j);
            const float dt = __fsub_rn(tan, tbn);
            float u_ae[TU], u_de[TU];
#pragma unroll
            for (int c = 0; c < TU; c++) { u_ae[c] = un_ae[c]; u_de[c] = un_de[c]; }
            const float gxn = own_gx(j - 1), gin = own_gi(j - 1);
            if (wk == 4) load_zv(j - 1, -1, un_ae);                       // point j-1
            if (j > 1) {
                tan = tbn; tbn = ldser(q.t, j - 2, bbown, 0);
                if (wk == 4) load_held(j - 1, event_of_step(j - 1), un_de);
            }
            // ---- point j: i_j = ae(x_j, z[j], v[j]) ----
            lam += chain(AE_NET{}, own_i ? mu : 0.0f, own_x(j), u_ae, rec_ptr(psn_dae_rec_stage(j, NST - 1, NST)), fresh_ae,
                         D1a, dB2a, dB3a, dB4a, dui_dummy);
            fresh_ae = false;
            // ---- step j ----
            float dxs = lam, d1 = 0.f, d2 = 0.f, d3 = 0.f, dcur, dui = 0.0f;
            {
                const float ld = lam * dt;
                if (METHOD == PSNODE_RK4) { d1 = ld * 0.125f; d2 = ld * 0.375f; d3 = ld * 0.375f; dcur = ld * 0.125f; }
                else dcur = ld;
            }
            const int64_t after_step = k >= 0 ? psn_dae_rec_event(k, T, NST) : psn_dae_rec_point(j - 1, T, NST);
#pragma unroll 1
            for (int e = NST - 1; e >= 0; e--) {
                const float* rec = rec_ptr(psn_dae_rec_stage(j, e, NST));
                const float* rnext = e > 0 ? rec - PSN_TAPE_STAGE : rec_ptr(after_step);
                const float yv = __ldcs(rec + ytape);
                float du_e = 0.0f;
                const float gq = chain(DE_NET{}, dcur, yv, u_de, rnext, fresh_de, D1, dB2, dB3, dB4, du_e);
                fresh_de = false;
                dui += du_e;
                dxs += gq;
                if (METHOD == PSNODE_RK4) {
                    const float tg = dt * gq;
                    if (e == 3) { d3 += tg; d2 -= tg; d1 += tg; dcur = d3; }
                    else if (e == 2) { d2 += tg; d1 -= tg * c13; dcur = d2; }
                    else if (e == 1) { d1 += tg * c13; dcur = d1; }
                } else if (METHOD == PSNODE_MIDPOINT) {
                    if (e == 1) dcur = (0.5f * dt) * gq;
                }
            }
            lam = dxs + gxn;
            if (k >= 0) {           // back through i_0 = ae(x_{j-1}, z_jump[k], v_jump[k])  (my_solvers.py:108-110); the z / v part of
                                    // u_de is exactly that input (its i columns only reach unused columns of the AE's G area)
                lam += chain(AE_NET{}, own_i ? dui : 0.0f, own_x(j - 1), u_de, rec_ptr(psn_dae_rec_point(j - 1, T, NST)), fresh_ae,
                             D1a, dB2a, dB3a, dB4a, dui_dummy);
                mu = gin;
            } else {
                mu = dui + gin;
            }
            if (((T - j) % (PSN_DW_CHAIN / NST)) == 0) { flush_dw(!fresh_de, !fresh_ae); fresh_de = true; fresh_ae = true; }
        }
        // ---- point 0: i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95) ----
        lam += chain(AE_NET{}, own_i ? mu : 0.0f, own_x(0), un_ae, nullptr, fresh_ae, D1a, dB2a, dB3a, dB4a, dui_dummy);
        flush_dw(!fresh_de, true);
        if (q.d_x0 && valid && own_x_ok) q.d_x0[(int64_t)bown * q.d_x0_sb + srow] = lam;

        // ---- bias gradients and layer-1 unfolding, one net after the other through the same scratch ----------------
        float* scr = reinterpret_cast<float*>(gs.dA_hi[0]);          // 32 KB of dead tiles
        float* D1s = scr;                 // [64][17]
        float* B2s = D1s + TH * 17;
        float* B3s = B2s + TH * 17;
        float* Gs = B3s + TH * 17;        // [64][25]
        float* a0s = Gs + TH * 25;        // [16][25]
        float* dks = a0s + TN * 25;       // [16][17]
        for (int net = 0; net < 2; net++) {
            group_sync();
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int o = frag_row(i) * 17 + frag_col(i);
                D1s[o] = net ? D1a[i] : D1[i]; B2s[o] = net ? dB2a[i] : dB2[i]; B3s[o] = net ? dB3a[i] : dB3[i];
            }
            dks[srow * 17 + sn] = net ? dB4a : dB4;
            const float* garea = net ? garea_ae : garea_de;
            for (int e = gt; e < TH * TK1; e += GROUP_THREADS) Gs[(e / TK1) * 25 + (e % TK1)] = garea[e];
            if (net == 0)
                for (int e = gt; e < TN * S; e += GROUP_THREADS) {
                    const int n = e / S, c = e - n * S;
                    a0s[n * 25 + c] = __ldg(q.a0 + (int64_t)min(b0 + n, B - 1) * q.a0_sb + c);
                }
            group_sync();
            const int ob_1 = net ? oab1 : ob1, ob_2 = net ? oab2 : ob2, ob_3 = net ? oab3 : ob3, ob_4 = net ? oab4 : ob4;
            const int n4 = net ? I : X;
            if (gt < 3 * TH) {
                const int which = gt >> 6, m = gt & 63;
                const float* src = (which == 0 ? D1s : (which == 1 ? B2s : B3s)) + m * 17;
                float acc = 0.0f;
#pragma unroll
                for (int n = 0; n < TN; n++) acc += src[n];
                sl[(which == 0 ? ob_1 : (which == 1 ? ob_2 : ob_3)) + m] = acc;
            } else if (gt < 3 * TH + n4) {
                const int c = gt - 3 * TH;
                float acc = 0.0f;
#pragma unroll
                for (int n = 0; n < TN; n++) acc += dks[c * 17 + n];
                sl[ob_4 + c] = acc;
            }
            if (net == 0) {     // DE layer 1: W1 = [Wa | Wb | Wc] acting on [a0; s - a0; s]
                for (int e = gt; e < TH * S; e += GROUP_THREADS) {
                    const int m = e / S, c = e - m * S;
                    float P = 0.0f;
#pragma unroll
                    for (int n = 0; n < TN; n++) P = fmaf(D1s[m * 17 + n], a0s[n * 25 + c], P);
                    const float G = Gs[m * 25 + (c < X ? c : TX + (c - X))];      // tile column of s[c]
                    sl[oW1 + m * K1 + c] = P;
                    sl[oW1 + m * K1 + S + c] = G - P;
                    sl[oW1 + m * K1 + 2 * S + c] = G;
                }
            } else {            // AE layer 1: A1 = [Aa | Ax Az Av] acting on [a0; x; z; v]
                for (int e = gt; e < TH * KA; e += GROUP_THREADS) {
                    const int m = e / KA, c = e - m * KA;
                    float val;
                    if (c < S) {
                        val = 0.0f;
#pragma unroll
                        for (int n = 0; n < TN; n++) val = fmaf(D1s[m * 17 + n], a0s[n * 25 + c], val);
                    } else {
                        const int sc = c - S;                                     // index into cat(x, z, v)
                        val = Gs[m * 25 + (sc < X ? sc : TX + (sc - X))];
                    }
                    sl[oA1 + m * KA + c] = val;
                }
            }
            if (q.d_a0) {       // d_a0 = (Wa - Wb)^T D1 (DE)  +  Aa^T D1a (AE)
                for (int e = gt; e < TN * S; e += GROUP_THREADS) {
                    const int n = e / S, c = e - n * S, b = b0 + n;
                    if (b >= B) continue;
                    float acc = 0.0f;
                    if (net == 0) {
                        for (int m = 0; m < TH; m++)
                            acc = fmaf(__ldg(q.W1 + m * K1 + c) - __ldg(q.W1 + m * K1 + S + c), D1s[m * 17 + n], acc);
                        q.d_a0[(int64_t)b * q.d_a0_sb + c] = acc;
                    } else {
                        for (int m = 0; m < TH; m++) acc = fmaf(__ldg(q.A1 + m * KA + c), D1s[m * 17 + n], acc);
                        q.d_a0[(int64_t)b * q.d_a0_sb + c] += acc;
                    }
                }
            }
        }
    }
    // ---- teardown --------------------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (wk == 0) tmem_dealloc(tmem, TM_COLS);
}

__global__ void psn_tc_dae_grad_reduce_kernel(const float* __restrict__ slab, int n_slabs, int n_theta, int n_de, int stride, float* __restrict__ d_theta) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_theta; o += gridDim.x * blockDim.x) {
        const int i = o < n_de ? o : o - n_de + pad4(n_de);      // slab position of parameter o
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int s = 0;
        for (; s + 3 < n_slabs; s += 4) {
            a0 += slab[(size_t)s * stride + i];
            a1 += slab[(size_t)(s + 1) * stride + i];
            a2 += slab[(size_t)(s + 2) * stride + i];
            a3 += slab[(size_t)(s + 3) * stride + i];
        }
        for (; s < n_slabs; s++) a0 += slab[(size_t)s * stride + i];
        d_theta[o] = (a0 + a1) + (a2 + a3);
    }
}

}  // namespace

bool psn_tc_supports(const psnode_problem* p);

// what the scripts' DAE training step needs: parameter gradients of both nets, d_x0 (x_init), d_a0
bool psn_tc_dae_bwd_supports(const psnode_problem* p, const psnode_adjoint* a) {
    if (p->kind != PSNODE_DAE || !psn_tc_supports(p) || !p->tape) return false;
    if (p->tape_floats < psn_tc_dae_tape_floats(p->B, p->T, p->method, p->event_idx ? p->E : 0)) return false;
    if (a->d_z.p || a->d_v.p || a->d_zjump || a->d_vjump || a->d_xteach.p || a->d_iteach.p) return false;
    return true;
}

int64_t psn_tc_dae_backward_workspace(const psnode_problem* p, const psnode_adjoint*) {
    return 256 + (int64_t)psn_tc_ngroups(p->B) * slab_floats((int)psnode_mlp_param_count(&p->de), (int)psnode_mlp_param_count(&p->ae)) * 4;
}

int psn_tc_dae_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    if (ws == nullptr || ws_bytes < psn_tc_dae_backward_workspace(p, a)) return PSNODE_EWORKSPACE;
    const int64_t n_de = psnode_mlp_param_count(&p->de), n_theta = n_de + psnode_mlp_param_count(&p->ae);
    if (a->n_theta != n_theta) return PSNODE_EINVAL;
    DaeBwdParams q;
    q.B = p->B; q.T = p->T; q.X = p->X; q.Z = p->Z; q.V = p->V; q.I = p->I; q.S = p->X + p->Z + p->V + p->I;
    q.E = p->event_idx ? p->E : 0;
    q.n_theta = (int)n_theta; q.n_theta_de = (int)n_de;
    q.t = p->t; q.z = p->z; q.v = p->v; q.gx = a->gx; q.gi = a->gi;
    q.fx = psn_make_fuse(a->fuse_x, p->x_sol); q.fi = psn_make_fuse(a->fuse_i, p->i_sol);
    q.x_sol = p->x_sol.p; q.xs_st = p->x_sol.st; q.xs_sb = p->x_sol.sb;
    q.i_sol = p->i_sol.p; q.is_st = p->i_sol.st; q.is_sb = p->i_sol.sb;
    q.a0 = p->a0; q.a0_sb = p->a0_sb;
    q.event_idx = p->event_idx;
    q.z_jump = p->z_jump; q.zj_sb = p->zj_sb; q.zj_se = p->zj_se;
    q.v_jump = p->v_jump; q.vj_sb = p->vj_sb; q.vj_se = p->vj_se;
    q.W1 = p->de.W[0]; q.W2 = p->de.W[1]; q.W3 = p->de.W[2]; q.W4 = p->de.W[3];
    q.A1 = p->ae.W[0]; q.A2 = p->ae.W[1]; q.A3 = p->ae.W[2]; q.A4 = p->ae.W[3];
    q.tape = p->tape;
    q.err = static_cast<int*>(ws);
    q.slab = reinterpret_cast<float*>(static_cast<char*>(ws) + 256);
    q.d_x0 = a->d_x0; q.d_x0_sb = a->d_x0_sb;
    q.d_a0 = a->d_a0; q.d_a0_sb = a->d_a0_sb;
    const int ngroups = psn_tc_ngroups(p->B);
    PSN_CUDA(cudaMemsetAsync(ws, 0, 256 + (size_t)ngroups * slab_floats((int)n_de, (int)(n_theta - n_de)) * 4, stream));
    const int smem = (int)sizeof(DaeBwdSmem) + 128;
    auto launch = [&](auto kern, const char* name) -> int {
        PSN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<ngroups, GROUP_THREADS, smem, stream>>>(q);
        psn_count_launch(name);
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    };
    int st;
    const bool fused = a->fuse_x.target.p || a->fuse_i.target.p;
    switch (p->method) {
        case PSNODE_EULER: st = fused ? launch(psn_tc_bwd_dae_kernel<PSNODE_EULER, true>, "psn_tc_bwd_dae_kernel<euler,fused-loss>")
                                      : launch(psn_tc_bwd_dae_kernel<PSNODE_EULER, false>, "psn_tc_bwd_dae_kernel<euler>"); break;
        case PSNODE_MIDPOINT: st = fused ? launch(psn_tc_bwd_dae_kernel<PSNODE_MIDPOINT, true>, "psn_tc_bwd_dae_kernel<midpoint,fused-loss>")
                                         : launch(psn_tc_bwd_dae_kernel<PSNODE_MIDPOINT, false>, "psn_tc_bwd_dae_kernel<midpoint>"); break;
        default: st = fused ? launch(psn_tc_bwd_dae_kernel<PSNODE_RK4, true>, "psn_tc_bwd_dae_kernel<rk4,fused-loss>")
                            : launch(psn_tc_bwd_dae_kernel<PSNODE_RK4, false>, "psn_tc_bwd_dae_kernel<rk4>"); break;
    }
    if (st != PSNODE_OK) return st;
    psn_tc_dae_grad_reduce_kernel<<<32, 256, 0, stream>>>(q.slab, ngroups, (int)n_theta, (int)n_de, slab_floats((int)n_de, (int)(n_theta - n_de)), a->d_theta);
    psn_count_launch("psn_tc_dae_grad_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
