// psnode_lg.cu -- "layer GEMM" path (impl = layer) for the latent `*_02_direct_encode` nets whose weights do not fit on one SM:
// DAE_02 / ODE_02 with X = Z (= V = I) = hidden = 128 or 256 (BASELINE configs[4]: DAE_Model, H = 256, 8192 trajectories per GPU;
// neural_01_DAE_02_direct_encode.py:70-100, :122, :137-147; integrate_DAE, neural_dae/my_solvers.py:82-131).
//
// At these shapes ONE layer of the stage MLP over the whole batch shard is a 1 GFLOP dense GEMM (8192 x 256 x 256), large
// enough to fill the chip by itself, while the five in-loop weight matrices (2.5 MB as tf32 hi + lo) exceed any SM's shared
// memory + TMEM.  So the time loop is a stream of per-layer tcgen05 GEMM launches with fused epilogues instead of one
// persistent kernel; activations (8 MB per tensor) stay L2-resident between launches.  Per step of the DAE:
//     [event: h = ELU(A1x x + preAE_jump[k]);  i0 = A2 h + ab2]                       (my_solvers.py:108-110)
//     G  = F_i i0                                                                     (held across the stages)
//     per stage:  a1 = ELU(F_x y + preDE[row] + G);   k = W2 a1 + b2 -> Runge-Kutta algebra -> next y / x_j
//     h  = ELU(A1x x_j + preAE[j]);   i_j = A2 h + ab2                                (:121, one explicit evaluation)
// with the folded layer 1 of SURVEY 8d: F = W_b + W_c, and the state-independent halves hoisted over the whole series by the
// same kernel:  preDE[r] = [F_z | F_v] [z[r]; v[r]] + c_de,  preAE[r] = [A1z | A1v] [z[r]; v[r]] + c_ae.
//
// GEMM kernel: D[feature m][trajectory n] = A . B^T, CTA tile 128 x 128, K in chunks of 32 (one SWIZZLE_128B slab).
//   A = prepared weights, tf32 hi and lo planes written once per call (lg_prep_kernel), streamed by TMA (UTMALDG);
//   B = fp32 activations / series rows, streamed by TMA through 3-D tensor maps over the strided views, split into hi / lo
//       in shared memory by the threads (3xTF32: A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation in TMEM, 2 K-partials);
//   3-stage mbarrier pipeline, single-thread MMA issue, epilogue from TMEM with one output feature per lane so that every
//   global access of a warp is a full 128-byte line of the (trajectory, feature) row-major tensors.
// Bound per launch: L2 -> SM operand traffic (384 KB per CTA) against 96 M128.N128.K8 MMAs.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

// hi / lo split of the streamed operand.  The tensor core TRUNCATES a tf32 operand to 10 mantissa bits; with PSN_LG_ROUND_LO the lo part is
// rounded to nearest here first (2 integer instructions per element), which halves its error and removes the truncation's sign dependence
__device__ __forceinline__ float4 lg_split4(const float4 v, float4& lo) {
    float4 hi = split4_hi(v, lo);
#if PSN_LG_ROUND_LO
    lo.x = __uint_as_float((__float_as_uint(lo.x) + 0x1000u) & 0xFFFFE000u); lo.y = __uint_as_float((__float_as_uint(lo.y) + 0x1000u) & 0xFFFFE000u);
    lo.z = __uint_as_float((__float_as_uint(lo.z) + 0x1000u) & 0xFFFFE000u); lo.w = __uint_as_float((__float_as_uint(lo.w) + 0x1000u) & 0xFFFFE000u);
#endif
    return hi;
}

constexpr int TM = 128, TN = 128;
constexpr int NST = 3;
constexpr int SLAB = TN * 128;          // 16 KB: 128 rows x 32 fp32
constexpr int LG_THREADS = 320;         // 8 split / epilogue warps + TMA producer warp + MMA issuer warp
#ifndef PSN_LG_NPART
#define PSN_LG_NPART 2
#endif
#ifndef PSN_LG_CLASS_SPLIT
#define PSN_LG_CLASS_SPLIT 1            // partial accumulators by term class (small cross terms | big term), see the issuer loop; 0: by K range
#endif
#ifndef PSN_LG_ROUND_LO
#define PSN_LG_ROUND_LO 0
#endif
// Partial accumulators (the tensor core's fp32 accumulation truncates: chains must stay short).  Measured on a 2000-step cfg5-width run
// against the float64 oracle (tests/test_gpu_layer.py::test_layer_dae_full_length_drift; the reference's own fp32 run: 5.6e-6 off):
//   2 partials by K range 1.65e-5 | 2 by term class 1.37e-5 (same speed, the default) | 4 by K range 9.8e-6 (-7 % speed) |
//   4 by class x K half 8.0e-6 (-7.5 %: -DPSN_LG_NPART=4) | rounding the streamed operand's lo part to nearest first: no effect
constexpr int NPART = PSN_LG_NPART;
constexpr int GEN_W = 8;                // widest raw input of a generated B source

enum { LG_PLAIN = 0, LG_HIDDEN = 1, LG_RK = 2, LG_DELTA = 3, LG_BRK = 4 };      // LG_BRK: LgParams::stage = the EPI_B* kind
// epilogue kinds (template parameter of the GEMM kernel)
enum { EPI_PLAIN = 0, EPI_HIDDEN, EPI_EULER, EPI_MID0, EPI_MID1, EPI_RK0, EPI_RK1, EPI_RK2, EPI_RK3, EPI_DELTA,
       EPI_BRK3, EPI_BRK2, EPI_BRK1, EPI_BMID1, EPI_BSUM, EPI_COUNT };
// Reverse pass (psn_lg_backward):
// EPI_DELTA: out = D * ELU'(act), act = the recomputed post-ELU activation (add1 operand); out2 (+)= out; acc += out.
// EPI_B*:    D = dy_e = F_x^T delta1_e of stage e; the Runge-Kutta adjoint (transpose of my_fixed_grid.py:15-59) turns it into the slope
//            adjoint of the stage before (out = dk_{e-1}, acc += dk_{e-1}; dy_e kept in k1 / k2 / k3), and after stage 0 into the
//            state adjoint of the previous grid point: out = x0 + k1 + k2 + k3 + D + add1  (x0 = gx_j, k* = the kept dy, add1 = dL/dx_sol[j-1]).

struct __align__(1024) LgSmem {
    unsigned char a_hi[NST][SLAB], a_lo[NST][SLAB], b_hi[NST][SLAB], b_lo[NST][SLAB];
    uint64_t full[NST], split[NST], done[NST];
    uint32_t tmem_base;
    float raw[2][TN][GEN_W];            // encoder fusion: the raw input rows of this trajectory tile (B-generator mode)
};

__device__ long long g_lg_dbg[64];        // clock64 stamps of one CTA (PSNODE_LG_DBG=<cta index + 1>): phase breakdown on the device
// PSNODE_LG_TRACE=1: per launch, over all CTAs: earliest entry / earliest and latest return of griddepcontrol.wait / latest exit (%globaltimer, ns)
constexpr int LG_TRACE_N = 4096;
__device__ unsigned long long g_lg_tmin[LG_TRACE_N][2], g_lg_tmax[LG_TRACE_N][2];
__device__ __forceinline__ unsigned long long lg_gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct LgParams {
    int dbg, rotate;
    int N, R, nbt;                      // trajectories per row, rows (1 for a layer launch), n-tiles per row
    int nsrc, kchunks;                  // B sources (K segments) and 32-wide chunks per source
    int mode;
    // event handling (neural_base.py:52-65, 180-196): ev[ev_j] = index of the event that fires when leaving grid point ev_j, or -1
    const int32_t* ev; int ev_j; int skip_unless_event; int skip_if_event;
    int b_r0;                           // added to the row coordinate of the B operand (slot of a ring of activation buffers)
    int a_c0;                           // added to the K coordinate of the A operand (second half of a [P | Q] plane used on its own)
    int out2_acc;                       // EPI_DELTA: out2 += out instead of out2 = out
    float* out2_jump; int64_t out2_jump_sr;   // when an event fires at ev_j, out2 = out2_jump + k * out2_jump_sr (row of the event)
    const float* add1; int64_t add1_sr, add1_ld;             // [r][n][m]  (hoisted layer-1 half / per-trajectory constant)
    const float* add1_jump; int64_t add1_jump_sr;            // event rows of add1 (selected when ev[ev_j] >= 0)
    const float* add2; int64_t add2_ld;                      // [n][m]
    const float* bias;                                       // [m]
    float* out; int64_t out_sr, out_ld;                      // [r][n][m]
    float* out2; int64_t out2_ld;                            // second copy (trajectory row / i_sol row)
    // Runge-Kutta epilogue
    int method, stage;
    float* x0; float* k1; float* k2; float* k3; int64_t st_ld;
    float* acc;                                              // reverse pass: acc[n][m] += out (row stride st_ld)
    const float* t_cur; const float* t_prev; int64_t t_sb;   // t[j], t[j-1] rows (element n at n * t_sb)
    // B-generator (encoder fusion): bit s of `gen` set = source s of the B operand is not loaded but computed in
    // shared memory as the hidden layer of an input encoder, b[n][k] = ELU(sum_c gen_W[k][c] * raw[r][n][c] + gen_b[k]), from the raw
    // (rows, trajectories, gen_w <= 8) series: the encoded (T, B, H) latent series never exists in HBM
    int gen;
    const float* gen_raw[2]; int64_t gen_sr[2], gen_sb[2]; int gen_w[2];
    const float* gen_W[2]; const float* gen_b[2];
    int m_live;                                              // > 0: output features m >= m_live are padding (not stored)
    int trace;                                               // >= 0: slot of the launch trace (debugging aid)
    int* err;
};

template <int EPI>
__global__ void __launch_bounds__(LG_THREADS, 1) psn_lg_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                                                                   const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1,
                                                                   const __grid_constant__ LgParams q) {
    // programmatic dependent launch: everything above the first read of the previous launch's output may overlap its tail
    if (q.trace >= 0 && threadIdx.x == 0) atomicMin(&g_lg_tmin[q.trace][0], lg_gtime());
    if (threadIdx.x == 0) {               // the descriptors are kernel parameters: fetch them while the previous launch is still running
        prefetch_tmap(&map_a_hi); prefetch_tmap(&map_a_lo);
        if (!(q.gen & 1)) prefetch_tmap(&map_b0);
        if (q.nsrc > 1 && !(q.gen & 2)) prefetch_tmap(&map_b1);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (q.trace >= 0 && threadIdx.x == 0) { const unsigned long long tt = lg_gtime(); atomicMin(&g_lg_tmin[q.trace][1], tt); atomicMax(&g_lg_tmax[q.trace][1], tt); }
    int evk = -1;
    if (q.ev) evk = __ldg(q.ev + q.ev_j);
    if (q.skip_unless_event && evk < 0) return;
    if (q.skip_if_event && evk >= 0) return;
    extern __shared__ unsigned char smem_raw[];
    LgSmem& sm = *reinterpret_cast<LgSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = cw & 3, hh = cw >> 2;
    const int r = blockIdx.x / q.nbt, b0 = (blockIdx.x - r * q.nbt) * TN;
    const int mblk = blockIdx.y;
    const int nchunk = q.nsrc * q.kchunks;

    const bool dbg = q.dbg != 0 && blockIdx.x == q.dbg - 1 && blockIdx.y == 0 && tid == 0;
    int dbg_n = 0;
    auto stamp = [&]() { if (dbg && dbg_n < 32) g_lg_dbg[dbg_n++] = clock64(); };
    stamp();
    // Warp roles: warps 0..7 split the B slabs and run the epilogue, warp 8 is the TMA producer, warp 9 the MMA issuer.  Each role
    // only ever waits on the mbarrier of the role before it (full -> split -> done -> refill), so the three run concurrently
    // up to the depth of the ring.  (Version 1 had thread 0 do all three in sequence with a CTA barrier per chunk: 1750 cycles
    // per chunk against 768 of tensor-pipe work.)
    // K chunks are visited in an order rotated by the CTA index (fixed, deterministic order per trajectory tile).
    const int rot = q.rotate ? (int)(blockIdx.x % (unsigned)nchunk) : 0;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.split[s], 8); mbar_init(&sm.done[s], 1); }
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, NPART * TN);
    if (q.gen) {
        for (int e = tid; e < 2 * TN * GEN_W; e += LG_THREADS) {
            const int src = e / (TN * GEN_W), n = (e / GEN_W) % TN, cc = e % GEN_W;
            float v = 0.0f;
            if (((q.gen >> src) & 1) && cc < q.gen_w[src] && b0 + n < q.N)
                v = __ldg(q.gen_raw[src] + (int64_t)r * q.gen_sr[src] + (int64_t)(b0 + n) * q.gen_sb[src] + cc);
            sm.raw[src][n][cc] = v;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;
    stamp();

    if (cw == 8) {
        // ---- TMA producer ----
        if (elect_one()) {
            for (int c = 0; c < nchunk; c++) {
                const int s = c % NST;
                if (c >= NST && !mbar_wait(&sm.done[s], (uint32_t)(((c - NST) / NST) & 1))) { atomicExch(q.err, 12); __trap(); }
                int ck = c + rot; if (ck >= nchunk) ck -= nchunk;
                const int src = ck / q.kchunks, kc = ck - src * q.kchunks;
                const bool generated = (q.gen >> src) & 1;
                mbar_expect_tx(&sm.full[s], (generated ? 2 : 3) * SLAB);
                tma_load_3d(sm.a_hi[s], &map_a_hi, ck * 32 + q.a_c0, mblk * TM, 0, &sm.full[s]);
                tma_load_3d(sm.a_lo[s], &map_a_lo, ck * 32 + q.a_c0, mblk * TM, 0, &sm.full[s]);
                if (!generated) tma_load_3d(sm.b_hi[s], src == 0 ? &map_b0 : &map_b1, kc * 32, b0, r + q.b_r0, &sm.full[s]);
            }
        }
        __syncwarp();
    } else if (cw == 9) {
        // ---- MMA issuer ----
        const uint32_t idesc = make_idesc_tf32(TM, TN);
        for (int c = 0; c < nchunk; c++) {
            const int s = c % NST;
            if (!mbar_wait(&sm.split[s], (uint32_t)((c / NST) & 1))) { atomicExch(q.err, 14); __trap(); }
            if (elect_one()) {
                tc_fence_after();
                const uint64_t da_hi = make_desc_sw128(smem_u32(sm.a_hi[s])), da_lo = make_desc_sw128(smem_u32(sm.a_lo[s]));
                const uint64_t db_hi = make_desc_sw128(smem_u32(sm.b_hi[s])), db_lo = make_desc_sw128(smem_u32(sm.b_lo[s]));
#if PSN_LG_CLASS_SPLIT
                // partial accumulators by TERM CLASS: the two small cross terms (A_lo.B_hi, A_hi.B_lo: 2^-11 of the result) go to their own
                // accumulator, the big term A_hi.B_hi to another (x K-halves when NPART = 4): the truncating fp32 accumulation of the tensor
                // core then only sees chains of nchunk * 4 / (NPART / 2) same-magnitude MMAs, and the small terms lose nothing to a large sum
                constexpr int KH = NPART / 2;
                const int kh = (c * KH) / nchunk;
                const bool first = c == (kh * nchunk + KH - 1) / KH;               // first chunk of this K-half
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint64_t ad = term == 0 ? da_lo : da_hi;
                    const uint64_t bd = term == 1 ? db_lo : db_hi;
                    const uint32_t acc = tmem + (uint32_t)(((term == 2 ? 1 : 0) * KH + kh) * TN);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
                        mma_tf32(acc, ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, (first && kk == 0 && term != 1) ? 0u : 1u);
                }
#else
                const int part = (c * NPART) / nchunk;
                const bool first = c == (part * nchunk + NPART - 1) / NPART;       // first chunk of this partial
                uint32_t accumulate = first ? 0u : 1u;
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint64_t ad = term == 0 ? da_lo : da_hi;
                    const uint64_t bd = term == 1 ? db_lo : db_hi;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        mma_tf32(tmem + (uint32_t)(part * TN), ad + (uint64_t)(2 * kk), bd + (uint64_t)(2 * kk), idesc, accumulate);
                        accumulate = 1;
                    }
                }
#endif
                mma_commit(&sm.done[s]);
            }
            __syncwarp();
        }
    } else {
        // ---- splitters: raw fp32 B slab -> tf32 hi (in place) + lo (elementwise: the swizzled layout is preserved) ----
        for (int c = 0; c < nchunk; c++) {
            const int s = c % NST;
            if (!mbar_wait(&sm.full[s], (uint32_t)((c / NST) & 1))) { atomicExch(q.err, 11); __trap(); }
            stamp();
            float4* h4 = reinterpret_cast<float4*>(sm.b_hi[s]);
            float4* l4 = reinterpret_cast<float4*>(sm.b_lo[s]);
            int gsrc = -1, gkc = 0;
            if (q.gen) {
                int ck = c + rot; if (ck >= nchunk) ck -= nchunk;
                const int src = ck / q.kchunks;
                if ((q.gen >> src) & 1) { gsrc = src; gkc = ck - src * q.kchunks; }
            }
            if (gsrc >= 0) {
                // the slab is written as TMA would have delivered it: 128-byte row n, 16-byte unit u at position u ^ (n & 7)
                const float* gW = q.gen_W[gsrc];
                const float* gb = q.gen_b[gsrc];
                const int gw = q.gen_w[gsrc];
#pragma unroll 1
                for (int e = 0; e < SLAB / 16 / 256; e++) {
                    const int idx = tid + e * 256;
                    const int n = idx >> 3, k0 = gkc * 32 + (((idx & 7) ^ (n & 7)) << 2);
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        float pre = __ldg(gb + k0 + u);
                        for (int cc = 0; cc < gw; cc++) pre = fmaf(__ldg(gW + (k0 + u) * gw + cc), sm.raw[gsrc][n][cc], pre);
                        v[u] = psn_elu(pre);
                    }
                    float4 lo;
                    const float4 hi = lg_split4(make_float4(v[0], v[1], v[2], v[3]), lo);
                    h4[idx] = hi;
                    l4[idx] = lo;
                }
            } else {
#pragma unroll
                for (int e = 0; e < SLAB / 16 / 256; e++) {
                    const int idx = tid + e * 256;
                    float4 lo;
                    const float4 hi = lg_split4(h4[idx], lo);
                    h4[idx] = hi;
                    l4[idx] = lo;
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.split[s]);
        }
    }
    if (cw >= 8) {          // producer / issuer are done; they only join the final barrier
        tc_fence_before();
        __syncthreads();
        return;
    }
    stamp();

    // ---- epilogue: lane = output feature, columns = trajectories ---------------------------------------------------------
    // Operands of the fused epilogue (hoisted layer-1 half, G, or the Runge-Kutta state) are fetched 16 columns ahead of the
    // arithmetic, so their L2 latency is paid once per tile.  The epilogue kind is a template parameter and the batch loop is
    // rolled: the first version (run-time mode switch inside fully unrolled loops) was 16 000 instructions of straight-line
    // code executed once per CTA and spent 38 000 of its 54 000 cycles waiting for instruction fetches (`no_inst` stalls).
    {
        constexpr bool rk = EPI >= EPI_EULER && EPI <= EPI_RK3;
        constexpr bool brk = EPI >= EPI_BRK3 && EPI <= EPI_BSUM;
        const int m = mblk * TM + 32 * wq + lane;
        const float* add1 = q.add1;
        int64_t add1_row = (int64_t)r * q.add1_sr;
        if (evk >= 0 && q.add1_jump) { add1 = q.add1_jump; add1_row = (int64_t)evk * q.add1_jump_sr; }
        const bool m_ok = q.m_live <= 0 || m < q.m_live;
        const float bias = (q.bias && m_ok) ? __ldg(q.bias + m) : 0.0f;
        const float c13 = (float)(1.0 / 3.0);
        struct Buf { float p0[16], p1[16], p2[16], p3[16], dt[16]; };
        // per-thread base pointers (element (column c of this thread's 64, feature m) at base + c * ld): no 64-bit index
        // arithmetic and no per-element bounds test on full tiles
        const int ncol0 = b0 + 64 * hh;
        const bool full_tile = b0 + TN <= q.N;
        const float* pa1 = add1 ? add1 + add1_row + (int64_t)ncol0 * q.add1_ld + m : nullptr;
        const float* pa2 = q.add2 ? q.add2 + (int64_t)ncol0 * q.add2_ld + m : nullptr;
        float* pout = q.out + (int64_t)r * q.out_sr + (int64_t)ncol0 * q.out_ld + m;
        float* pout2 = q.out2 ? q.out2 + (int64_t)ncol0 * q.out2_ld + m : nullptr;
        if (evk >= 0 && q.out2_jump) pout2 = q.out2_jump + (int64_t)evk * q.out2_jump_sr + (int64_t)ncol0 * q.out2_ld + m;
        float* px0 = (rk || brk) ? q.x0 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk1 = (rk || brk) ? q.k1 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk2 = (rk || brk) ? q.k2 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk3 = (rk || brk) ? q.k3 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pacc = q.acc ? q.acc + (int64_t)ncol0 * q.st_ld + m : nullptr;
        const float* ptc = (rk || (brk && EPI != EPI_BSUM)) ? q.t_cur + (int64_t)ncol0 * q.t_sb : nullptr;
        const float* ptp = (rk || (brk && EPI != EPI_BSUM)) ? q.t_prev + (int64_t)ncol0 * q.t_sb : nullptr;
        const int a1ld = (int)q.add1_ld, a2ld = (int)q.add2_ld, old = (int)q.out_ld, o2ld = (int)q.out2_ld, sld = (int)q.st_ld, tsb = (int)q.t_sb;
        const int nlive = full_tile ? 64 : max(0, min(64, q.N - ncol0));      // live columns of this thread
        auto fetch = [&](int bt, Buf& e) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int c = full_tile ? 16 * bt + i : min(16 * bt + i, max(nlive - 1, 0));
                if constexpr (rk) {
                    e.p0[i] = px0[c * sld];
                    if constexpr (EPI >= EPI_RK1) e.p1[i] = pk1[c * sld];
                    if constexpr (EPI >= EPI_RK2) e.p2[i] = pk2[c * sld];
                    if constexpr (EPI >= EPI_RK3) e.p3[i] = pk3[c * sld];
                    e.dt[i] = __fsub_rn(__ldg(ptc + c * tsb), __ldg(ptp + c * tsb));
                } else if constexpr (brk) {
                    e.p0[i] = px0[c * sld];
                    if constexpr (EPI == EPI_BRK2 || EPI == EPI_BRK1 || EPI == EPI_BSUM) e.p1[i] = pk1[c * sld];
                    if constexpr (EPI == EPI_BRK1 || EPI == EPI_BSUM) e.p2[i] = pk2[c * sld];
                    if constexpr (EPI == EPI_BSUM) {
                        e.p3[i] = pk3[c * sld];
                        e.dt[i] = pa1 ? __ldg(pa1 + c * a1ld) : 0.0f;               // upstream gradient row
                    } else {
                        e.dt[i] = __fsub_rn(__ldg(ptc + c * tsb), __ldg(ptp + c * tsb));
                        e.p3[i] = pacc ? pacc[c * sld] : 0.0f;                      // read-modify-write operands are fetched ahead too
                    }
                } else if constexpr (EPI == EPI_DELTA) {
                    e.p0[i] = __ldg(pa1 + c * a1ld);
                    e.p1[i] = (pout2 && q.out2_acc) ? pout2[c * o2ld] : 0.0f;
                    e.p2[i] = pacc ? pacc[c * sld] : 0.0f;
                } else {
                    e.p0[i] = pa1 ? __ldg(pa1 + c * a1ld) : 0.0f;
                    e.p1[i] = pa2 ? __ldg(pa2 + c * a2ld) : 0.0f;
                }
            }
        };
        auto finish = [&](int bt, const Buf& e) {
            const int n0 = 64 * hh + 16 * bt;
            float t0[16], t1[16];
            tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)n0, t0);
            tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)(TN + n0), t1);
            if constexpr (NPART == 4) {          // two register arrays only: fold the partials pairwise as they arrive
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) t0[i] += t1[i];
                tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)(2 * TN + n0), t1);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) t0[i] += t1[i];
                tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)(3 * TN + n0), t1);
                tmem_ld_wait();
            } else {
                tmem_ld_wait();
            }
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int c = 16 * bt + i;
                if (!full_tile && c >= nlive) continue;
                float v = (t0[i] + t1[i]) + bias;
                if constexpr (EPI == EPI_DELTA) {
                    v = (t0[i] + t1[i]) * psn_elu_grad_from_out(e.p0[i]);
                    pout[c * old] = v;
                    if (pout2) pout2[c * o2ld] = e.p1[i] + v;
                    if (pacc) pacc[c * sld] = e.p2[i] + v;
                } else if constexpr (brk) {
                    const float D = t0[i] + t1[i], gx = e.p0[i];
                    if constexpr (EPI == EPI_BSUM) {
                        pout[c * old] = ((((gx + e.p1[i]) + e.p2[i]) + e.p3[i]) + D) + e.dt[i];
                    } else {
                        const float dt = e.dt[i];
                        float dk;
                        if constexpr (EPI == EPI_BRK3) { pk1[c * sld] = D; dk = dt * fmaf(0.375f, gx, D); }
                        else if constexpr (EPI == EPI_BRK2) { pk2[c * sld] = D; dk = dt * (fmaf(0.375f, gx, D) - e.p1[i]); }
                        else if constexpr (EPI == EPI_BRK1) { pk3[c * sld] = D; dk = dt * (fmaf(0.125f, gx, e.p1[i]) + c13 * (D - e.p2[i])); }
                        else { pk1[c * sld] = D; dk = (0.5f * dt) * D; }
                        pout[c * old] = dk;
                        if (pacc) pacc[c * sld] = e.p3[i] + dk;
                    }
                } else if constexpr (!rk) {
                    v = (v + e.p0[i]) + e.p1[i];
                    if constexpr (EPI == EPI_HIDDEN) v = psn_elu(v);
                    if (m_ok) pout[c * old] = v;
                    if (pout2) pout2[c * o2ld] = v;
                } else {
                    // reference operation order (neural_dae/my_fixed_grid.py:15-59)
                    const float dt = e.dt[i], x0 = e.p0[i], kk = v;
                    float xn;
                    constexpr bool last = EPI == EPI_EULER || EPI == EPI_MID1 || EPI == EPI_RK3;
                    if constexpr (EPI == EPI_EULER || EPI == EPI_MID1) xn = __fadd_rn(x0, __fmul_rn(dt, kk));
                    else if constexpr (EPI == EPI_MID0) xn = __fadd_rn(x0, __fmul_rn(kk, __fmul_rn(0.5f, dt)));
                    else if constexpr (EPI == EPI_RK0) { pk1[c * sld] = kk; xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(dt, kk), c13)); }
                    else if constexpr (EPI == EPI_RK1) { pk2[c * sld] = kk; xn = __fadd_rn(x0, __fmul_rn(dt, __fsub_rn(kk, __fmul_rn(e.p1[i], c13)))); }
                    else if constexpr (EPI == EPI_RK2) { pk3[c * sld] = kk; xn = __fadd_rn(x0, __fmul_rn(dt, __fadd_rn(__fsub_rn(e.p1[i], e.p2[i]), kk))); }
                    else {
                        const float ksum = __fadd_rn(__fadd_rn(e.p1[i], __fmul_rn(3.0f, __fadd_rn(e.p2[i], e.p3[i]))), kk);
                        xn = __fadd_rn(x0, __fmul_rn(__fmul_rn(ksum, dt), 0.125f));
                    }
                    pout[c * old] = xn;                                          // next stage input / x_j
                    if constexpr (last) {
                        px0[c * sld] = xn;
                        if (pout2) pout2[c * o2ld] = xn;                         // trajectory row j
                    }
                }
            }
        };
        // A warp whose 64 columns all lie beyond a ragged batch (B = 5: warps 4..7 of the only tile) has nothing to fetch or store; its
        // operand addresses would be out of bounds (compute-sanitizer memcheck), so it only keeps the barriers company.
        const bool warp_live = full_tile || nlive > 0;
        Buf cur, nxt;
        if (warp_live) fetch(0, cur);       // issued BEFORE the wait for the last MMAs: the first operand batch's L2 / HBM latency hides under them
        if (!mbar_wait(&sm.done[(nchunk - 1) % NST], (uint32_t)(((nchunk - 1) / NST) & 1))) { atomicExch(q.err, 13); __trap(); }
        tc_fence_after();
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the next layer's grid may be set up while this epilogue runs
        stamp();
#pragma unroll 1
        for (int bt = 0; bt < 4 && warp_live; bt++) {
            if (bt + 1 < 4) fetch(bt + 1, nxt);
            finish(bt, cur);
            cur = nxt;
        }
    }
    stamp();
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, NPART * TN);
    stamp();
    if (dbg) g_lg_dbg[63] = dbg_n;
    if (q.trace >= 0 && threadIdx.x == 0) atomicMax(&g_lg_tmax[q.trace][0], lg_gtime());
}

// ---- one-time preparation per call ------------------------------------------------------------------------------------------
// c[b][m] = bias[m] + sum_{k < S} (W[m][k] - (sub >= 0 ? W[m][sub + k] : 0)) a0[b][k]; block = 8 trajectories, thread = m (blockDim = H)
__global__ void psn_lg_const_kernel(const float* __restrict__ W, int ldw, int sub, const float* __restrict__ bias, const float* __restrict__ a0,
                                    int64_t a0_sb, int S, int B, int H, float* __restrict__ c) {
    extern __shared__ float a[];            // [8][S]
    const int m = threadIdx.x, b0 = blockIdx.x * 8;
    for (int e = m; e < 8 * S; e += blockDim.x) {
        const int n = e / S, k = e - n * S;
        a[e] = __ldg(a0 + (int64_t)min(b0 + n, B - 1) * a0_sb + k);
    }
    __syncthreads();
    float acc[8];
    const float bv = __ldg(bias + m);
#pragma unroll
    for (int n = 0; n < 8; n++) acc[n] = bv;
    for (int k = 0; k < S; k++) {
        float w = __ldg(W + (int64_t)m * ldw + k);
        if (sub >= 0) w -= __ldg(W + (int64_t)m * ldw + sub + k);
#pragma unroll
        for (int n = 0; n < 8; n++) acc[n] = fmaf(w, a[n * S + k], acc[n]);
    }
    for (int n = 0; n < 8; n++)
        if (b0 + n < B) c[(int64_t)(b0 + n) * H + m] = acc[n];
}
// initial state: x0 = ycur = x_sol[0] = x_init (DAE) / x[0] (ODE)
__global__ void psn_lg_init_kernel(const float* __restrict__ src, int64_t src_sb, int B, int H, float* __restrict__ x0, float* __restrict__ ycur,
                                   float* __restrict__ xsol0, int64_t xsol_sb) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, m = idx - b * H;
    const float v = __ldg(src + (int64_t)b * src_sb + m);
    x0[idx] = v;
    ycur[idx] = v;
    xsol0[(int64_t)b * xsol_sb + m] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn lg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// 3-D map (feature, row-in-batch, outer row) over p[outer * s_outer + row * s_row + feature], box {32, 128, 1}, SWIZZLE_128B
bool lg_make_map(CUtensorMap* map, const float* p, int width, int64_t rows, int64_t s_row, int64_t outer, int64_t s_outer) {
    if (!lg_encode_fn() || (reinterpret_cast<uintptr_t>(p) & 15) || (s_row & 3) || (outer > 1 && (s_outer & 3))) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)(outer > 0 ? outer : 1)};
    const cuuint64_t gstr[2] = {(cuuint64_t)s_row * 4, (cuuint64_t)(outer > 1 ? s_outer : s_row * rows) * 4};
    const cuuint32_t box[3] = {32, 128, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return lg_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- weight-gradient products of the layer path -----------------------------------------------------------------------------
// D[m][k] = sum_{slot, n} P[slot][n][m] * Q[slot][n][k]      (dW = sum over trajectories / stages / steps of delta (x) activation)
// Both operands are consumed in their natural (trajectory, feature) row-major layout: the reduction index n is the ROW index, i.e.
// they are MN-major UMMA operands.  A TMA box {32 features, 32 rows} with SWIZZLE_128B is exactly the canonical MN-major atom stack
// (8 K-rows x 128 B per atom): a K = 8 step advances the start by 1024 B (one atom), feature blocks of 32 are 4096 B apart (LBO).
// CTA tile 128 x 128, split-K over (slot, row-chunk) ranges; 3xTF32; the accumulator alternates between two TMEM buffers and is
// drained into registers (round-to-nearest) every 2 chunks (24 accumulations); per-split slabs are reduced in a fixed order.
constexpr int WG_THREADS = 320;
struct __align__(1024) WgSmem {
    unsigned char p_hi[NST][SLAB], p_lo[NST][SLAB], q_hi[NST][SLAB], q_lo[NST][SLAB];
    uint64_t full[NST], split[NST], done[NST], drained[2];
    uint32_t tmem_base;
};
struct WgParams {
    int nslots, N, nsplit;          // rows per slot N; split-K ways
    int mblks, kblks;
    float* slabs;                   // [split][tile][128][128]
    int* err;
    const int32_t* ev; int ev_j;    // skip (early exit) unless an event fires at ev_j (ev == NULL: never skip)
};
// MN-major tf32 operands exist in ONE shared-memory layout only: 128-byte swizzle with 32-byte atomicity (UMMA layout type 1,
// SWIZZLE_128B_BASE32B; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  The swizzle pattern repeats every 4 rows of 128 B, so the
// stride between K groups (SBO) is 512 B; with the plain 128B swizzle the tensor core returns zeros.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, int lbo, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) psn_lg_wgrad_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_q,
                                                                    const __grid_constant__ WgParams q) {
    if (q.ev && __ldg(q.ev + q.ev_j) < 0) return;
    extern __shared__ unsigned char smem_raw[];
    WgSmem& sm = *reinterpret_cast<WgSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = cw & 3, hh = (cw >> 2) & 1;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int mblk = tile / q.kblks, kblk = tile - mblk * q.kblks;
    const int cps = (q.N + 31) / 32;                       // row chunks per slot
    const int64_t total = (int64_t)q.nslots * cps;
    const int64_t c0 = total * split / q.nsplit, c1 = total * (split + 1) / q.nsplit;
    const int n = (int)(c1 - c0);

    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.split[s], 8); mbar_init(&sm.done[s], 1); }
        mbar_init(&sm.drained[0], 8); mbar_init(&sm.drained[1], 8);
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 2 * TN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;

    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; i++) acc[i] = 0.0f;

    if (cw == 8) {
        if (elect_one()) {
            for (int c = 0; c < n; c++) {
                const int s = c % NST;
                if (c >= NST && !mbar_wait(&sm.done[s], (uint32_t)(((c - NST) / NST) & 1))) { atomicExch(q.err, 21); __trap(); }
                const int64_t cc = c0 + c;
                const int slot = (int)(cc / cps), row0 = (int)(cc - (int64_t)slot * cps) * 32;
                mbar_expect_tx(&sm.full[s], 2 * SLAB);
#pragma unroll
                for (int fb = 0; fb < 4; fb++) {
                    tma_load_3d(sm.p_hi[s] + fb * 4096, &map_p, mblk * TM + fb * 32, row0, slot, &sm.full[s]);
                    tma_load_3d(sm.q_hi[s] + fb * 4096, &map_q, kblk * TN + fb * 32, row0, slot, &sm.full[s]);
                }
            }
        }
        __syncwarp();
    } else if (cw == 9) {
        const uint32_t idesc = make_idesc_tf32_mn(TM, TN);
        for (int c = 0; c < n; c++) {
            const int s = c % NST;
            const int pair = c >> 1;
            if ((c & 1) == 0 && pair >= 2 && !mbar_wait(&sm.drained[pair & 1], (uint32_t)(((pair - 2) >> 1) & 1))) { atomicExch(q.err, 22); __trap(); }
            if (!mbar_wait(&sm.split[s], (uint32_t)((c / NST) & 1))) { atomicExch(q.err, 23); __trap(); }
            if (elect_one()) {
                tc_fence_after();
                // LBO = 4096 B (next block of 32 features = one TMA box), SBO = 512 B (next group of 4 reduction rows), K = 8 step = 1024 B
                const uint64_t dp_hi = make_desc_mn(smem_u32(sm.p_hi[s]), 4096, 512), dp_lo = make_desc_mn(smem_u32(sm.p_lo[s]), 4096, 512);
                const uint64_t dq_hi = make_desc_mn(smem_u32(sm.q_hi[s]), 4096, 512), dq_lo = make_desc_mn(smem_u32(sm.q_lo[s]), 4096, 512);
                constexpr uint64_t kadv = 1024 >> 4;
                uint32_t accumulate = (c & 1) ? 1u : 0u;
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint64_t ad = term == 0 ? dp_lo : dp_hi;
                    const uint64_t bd = term == 1 ? dq_lo : dq_hi;
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        mma_tf32(tmem + (uint32_t)((pair & 1) * TN), ad + kadv * kk, bd + kadv * kk, idesc, accumulate);
                        accumulate = 1;
                    }
                }
                mma_commit(&sm.done[s]);
            }
            __syncwarp();
        }
    } else {
        auto drain = [&](int pair, int last_chunk) {
            if (!mbar_wait(&sm.done[last_chunk % NST], (uint32_t)((last_chunk / NST) & 1))) { atomicExch(q.err, 24); __trap(); }
            tc_fence_after();
#pragma unroll
            for (int b4 = 0; b4 < 4; b4++) {
                float v[16];
                tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)((pair & 1) * TN + 64 * hh + 16 * b4), v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) acc[16 * b4 + i] += v[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.drained[pair & 1]);
        };
        for (int c = 0; c < n; c++) {
            const int s = c % NST;
            if (!mbar_wait(&sm.full[s], (uint32_t)((c / NST) & 1))) { atomicExch(q.err, 25); __trap(); }
            float4* ph = reinterpret_cast<float4*>(sm.p_hi[s]); float4* pl = reinterpret_cast<float4*>(sm.p_lo[s]);
            float4* qh = reinterpret_cast<float4*>(sm.q_hi[s]); float4* ql = reinterpret_cast<float4*>(sm.q_lo[s]);
#pragma unroll
            for (int e = 0; e < SLAB / 16 / 256; e++) {
                const int idx = tid + e * 256;
                float4 lo;
                float4 hi = split4_hi(ph[idx], lo);
                ph[idx] = hi; pl[idx] = lo;
                hi = split4_hi(qh[idx], lo);
                qh[idx] = hi; ql[idx] = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.split[s]);
            // drain the previous pair once this pair's second chunk has been handed to the issuer
            if ((c & 1) == 1 && c >= 3) drain((c >> 1) - 1, c - 2);
        }
        if (n > 0) {        // pairs the loop has not drained (it drains pair p - 1 when it hands the second chunk of pair p over)
            const int last_pair = (n - 1) >> 1;
            const int cmax = ((n - 1) & 1) ? n - 1 : n - 2;                 // last odd chunk index
            const int upto = cmax >= 3 ? (cmax >> 1) - 1 : -1;             // last pair drained in the loop
            for (int pp = upto + 1; pp <= last_pair; pp++) drain(pp, min(2 * pp + 1, n - 1));
        }
        float* slab = q.slabs + ((int64_t)split * gridDim.x + tile) * TM * TN + (int64_t)(32 * wq + lane) * TN + 64 * hh;
#pragma unroll
        for (int i = 0; i < 64; i += 4) *reinterpret_cast<float4*>(slab + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 2 * TN);
}

// out[m][k] (+)= sum_s slabs[s][tile(m,k)][m%128][k%128]   (out row stride ld, fixed summation order)
__global__ void psn_lg_wgrad_reduce_kernel(const float* __restrict__ slabs, int nsplit, int ntiles, int kblks, int M, int K, float* __restrict__ out,
                                           int64_t ld, int accumulate, const int32_t* ev, int ev_j) {
    if (ev && __ldg(ev + ev_j) < 0) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * K) return;
    const int m = idx / K, k = idx - m * K;
    const int tile = (m / TM) * kblks + k / TN;
    const float* src = slabs + (int64_t)tile * TM * TN + (m % TM) * TN + (k % TN);
    float s = 0.0f;
    for (int sp = 0; sp < nsplit; sp++) s += src[(int64_t)sp * ntiles * TM * TN];
    float* dst = out + (int64_t)m * ld + k;
    *dst = accumulate ? *dst + s : s;
}

bool lg_make_map32(CUtensorMap* map, const float* p, int width, int64_t rows, int64_t s_row, int64_t outer, int64_t s_outer) {
    if (!lg_encode_fn() || (reinterpret_cast<uintptr_t>(p) & 15) || (s_row & 3) || (outer > 1 && (s_outer & 3))) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)width, (cuuint64_t)rows, (cuuint64_t)(outer > 0 ? outer : 1)};
    const cuuint64_t gstr[2] = {(cuuint64_t)s_row * 4, (cuuint64_t)(outer > 1 ? s_outer : s_row * rows) * 4};
    const cuuint32_t box[3] = {32, 32, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return lg_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int WG_NSPLIT_MAX = 37;
int64_t lg_wgrad_slab_floats(int M, int K) { return (int64_t)WG_NSPLIT_MAX * (M / TM) * (K / TN) * TM * TN; }

// out[M x K] (+)= sum over slots and rows of P^T Q.  P: (nslots, N, M) with strides (p_ss, p_sn, 1); Q: (nslots, N, K) likewise.
int lg_wgrad(const float* P, int64_t p_sn, int64_t p_ss, int M, const float* Q, int64_t q_sn, int64_t q_ss, int K, int nslots, int N, float* out,
             int64_t out_ld, int accumulate, float* slabs, int* err, cudaStream_t stream, const int32_t* ev = nullptr, int ev_j = 0, int max_ctas = 148) {
    CUtensorMap mp, mq;
    if (!lg_make_map32(&mp, P, M, N, p_sn, nslots, p_ss) || !lg_make_map32(&mq, Q, K, N, q_sn, nslots, q_ss))
        return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer weight gradients)");
    WgParams q;
    q.nslots = nslots; q.N = N;
    q.mblks = M / TM; q.kblks = K / TN;
    const int ntiles = q.mblks * q.kblks;
    const int64_t total = (int64_t)nslots * ((N + 31) / 32);
    int nsplit = max_ctas / ntiles;
    if (nsplit > WG_NSPLIT_MAX) nsplit = WG_NSPLIT_MAX;
    if (nsplit > total) nsplit = (int)total;
    if (nsplit < 1) nsplit = 1;
    q.nsplit = nsplit;
    q.slabs = slabs; q.err = err;
    q.ev = ev; q.ev_j = ev_j;
    const int smem = (int)sizeof(WgSmem) + 1024;
    static bool attr = false;
    if (!attr) { PSN_CUDA(cudaFuncSetAttribute(psn_lg_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
    psn_lg_wgrad_kernel<<<dim3((unsigned)ntiles, (unsigned)nsplit), WG_THREADS, smem, stream>>>(mp, mq, q);
    psn_count_launch("psn_lg_wgrad_kernel");
    psn_lg_wgrad_reduce_kernel<<<(M * K + 255) / 256, 256, 0, stream>>>(slabs, nsplit, ntiles, q.kblks, M, K, out, out_ld, accumulate, ev, ev_j);
    psn_count_launch("psn_lg_wgrad_reduce_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

using LgKernFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const LgParams);
const LgKernFn* lg_kernels() {
    static const LgKernFn kerns[EPI_COUNT] = {psn_lg_gemm_kernel<EPI_PLAIN>, psn_lg_gemm_kernel<EPI_HIDDEN>, psn_lg_gemm_kernel<EPI_EULER>,
                                              psn_lg_gemm_kernel<EPI_MID0>, psn_lg_gemm_kernel<EPI_MID1>, psn_lg_gemm_kernel<EPI_RK0>,
                                              psn_lg_gemm_kernel<EPI_RK1>, psn_lg_gemm_kernel<EPI_RK2>, psn_lg_gemm_kernel<EPI_RK3>,
                                              psn_lg_gemm_kernel<EPI_DELTA>, psn_lg_gemm_kernel<EPI_BRK3>, psn_lg_gemm_kernel<EPI_BRK2>,
                                              psn_lg_gemm_kernel<EPI_BRK1>, psn_lg_gemm_kernel<EPI_BMID1>, psn_lg_gemm_kernel<EPI_BSUM>};
    return kerns;
}
int lg_gemm_smem() { return (int)sizeof(LgSmem) + 1024; }
int lg_prepare_kernels() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < EPI_COUNT; i++) PSN_CUDA(cudaFuncSetAttribute(lg_kernels()[i], cudaFuncAttributeMaxDynamicSharedMemorySize, lg_gemm_smem()));
        attr_set = true;
    }
    return PSNODE_OK;
}
int lg_epi_of(const LgParams& q) {
    if (q.mode == LG_PLAIN) return (int)EPI_PLAIN;
    if (q.mode == LG_HIDDEN) return (int)EPI_HIDDEN;
    if (q.mode == LG_DELTA) return (int)EPI_DELTA;
    if (q.mode == LG_BRK) return q.stage;
    if (q.method == PSNODE_EULER) return (int)EPI_EULER;
    if (q.method == PSNODE_MIDPOINT) return q.stage == 0 ? (int)EPI_MID0 : (int)EPI_MID1;
    return (int)EPI_RK0 + q.stage;
}
bool lg_use_pdl() {
    static const bool v = std::getenv("PSNODE_LG_PDL") ? std::atoi(std::getenv("PSNODE_LG_PDL")) != 0 : true;
    return v;
}
// one GEMM launch: A planes (hi, lo maps), one or two B sources
int g_lg_trace_next = -1;                   // host side of PSNODE_LG_TRACE: next free trace slot (-1: tracing off)
const char* g_lg_trace_name[LG_TRACE_N];
int lg_launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b0, const CUtensorMap& b1, const LgParams& q_in, int mblks,
                   cudaStream_t stream, const char* name) {
    LgParams q = q_in;
    q.trace = -1;
    if (g_lg_trace_next >= 0 && g_lg_trace_next < LG_TRACE_N) { g_lg_trace_name[g_lg_trace_next] = name; q.trace = g_lg_trace_next++; }
    dim3 grid((unsigned)(q.R * q.nbt), (unsigned)mblks);
    const LgKernFn kern = lg_kernels()[lg_epi_of(q)];
    if (lg_use_pdl()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(LG_THREADS); cfg.dynamicSmemBytes = (size_t)lg_gemm_smem(); cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kern, a_hi, a_lo, b0, b1, q);
    } else {
        kern<<<grid, LG_THREADS, lg_gemm_smem(), stream>>>(a_hi, a_lo, b0, b1, q);
    }
    psn_count_launch(name);
    return PSNODE_OK;
}

int64_t al(int64_t floats) { return (floats + 63) & ~(int64_t)63; }

// prepared weight planes: the forward's seven (A = weights as they multiply activations) and the reverse pass's transposes
enum { M_DE1X = 0, M_DE1ZV, M_DE1I, M_DE2, M_AE1X, M_AE1ZV, M_AE2, M_C_DE, M_C_AE,
       M_T_DE2, M_T_DE1X, M_T_DE1I, M_T_AE2, M_T_AE1X, M_T_DZ, M_T_DV, M_T_DA0,
       M_XD1, M_XD2, M_ID1, M_ID2, NMAT };                                      // decoders of the encoded entry (M_?D2: rows padded to 128)
constexpr int NMAT_BWD = M_T_DA0 + 1;
constexpr int NMAT_FWD = M_C_AE + 1;

// BH-sized work buffers of the reverse pass (one contiguous array, one tensor map, addressed by index)
enum { BUF_GX = 0, BUF_DY3, BUF_DY2, BUF_DY1, BUF_GI0, BUF_HEV, BUF_DHEV, BUF_G, BUF_K1, BUF_K2, BUF_K3, BUF_DCDE, BUF_DCAE, BUF_ACCDK, BUF_ACCGI,
       BUF_UPX, BUF_UPI, BUF_SINGLES };

enum { LG_FWD = 0, LG_BWD = 1, LG_ENC = 2 };
struct LgLayout {
    int H, KZV, S, nmat, bwd, enc;
    // encoded entry: time-chunk scratch (rows of B x H floats)
    int rc;
    int64_t encb_de, encb_ae, pj_de, pj_ae, xs, is, dtmp;
    int mrows[NMAT], mcols[NMAT];
    int64_t err, wts_hi[NMAT], wts_lo[NMAT], c_de, c_ae, pre_de, pre_ae, x0, k1, k2, k3, ycur, a1, hbuf, icur, G, total;
    int64_t rows_de, rows_ae;
    // reverse pass
    int ring, nst, nbufs;
    int64_t dpj_de, dpj_ae, bufs, slabs;
    // `ring` = steps per weight-gradient batch, `depth` = physical slots (2 * ring when the batch GEMMs overlap the next batch's sweep)
    int depth;
    int64_t slabs2;
    int ring_y(int s, int e) const { return BUF_SINGLES + (0 * depth + s) * nst + e; }
    int ring_a1(int s, int e) const { return BUF_SINGLES + (1 * depth + s) * nst + e; }
    int ring_dk(int s, int e) const { return BUF_SINGLES + (2 * depth + s) * nst + e; }
    int ring_d1(int s, int e) const { return BUF_SINGLES + (3 * depth + s) * nst + e; }
    int ring_one(int kind, int s) const { return BUF_SINGLES + 4 * depth * nst + kind * depth + s; }      // kind: 0 i0, 1 dsum, 2 gi, 3 h, 4 dh
};
enum { R_I0 = 0, R_DSUM, R_GI, R_H, R_DH };

int lg_chunk_rows() {
    static const int v = std::getenv("PSNODE_LG_CHUNK") ? std::atoi(std::getenv("PSNODE_LG_CHUNK")) : 64;
    return v < 1 ? 1 : v;
}
// SMs a layer launch of this problem leaves idle (its grid is one CTA per 128 x 128 output tile)
int lg_idle_sms(const psnode_problem* p) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int grid = ((p->B + TN - 1) / TN) * (p->X / TM);
    return grid >= sms ? 0 : sms - grid;
}
// When a layer launch leaves a large part of the chip idle (H = 128: 64 CTAs at B = 8192; small batches), the ring-batch weight-gradient
// GEMMs of the reverse sweep run on a side stream on those SMs, capped at their number, one batch behind the sweep (measured at
// B = 8192: H = 128, 84 idle SMs: 404 -> 337 us per step; H = 256, 20 idle SMs: 569 -> 636 us, the side stream becomes the critical
// path -- so only with >= 48 idle SMs).  PSNODE_LG_OVERLAP=0 switches it off.  Not with events: their gated per-step weight-gradient
// launches accumulate into the same blocks and would race with the side stream.
bool lg_overlap_wgrad(const psnode_problem* p) {
    static const bool on = !(std::getenv("PSNODE_LG_OVERLAP") && std::atoi(std::getenv("PSNODE_LG_OVERLAP")) == 0);
    return on && !(p->event_idx && p->E > 0) && lg_idle_sms(p) >= 48;
}
int lg_ring_depth() {
    static const int v = std::getenv("PSNODE_LG_RING") ? std::atoi(std::getenv("PSNODE_LG_RING")) : 8;
    return v < 1 ? 1 : (v > 64 ? 64 : v);
}

LgLayout lg_layout(const psnode_problem* p, int mode, int chunk_rows = 0) {
    LgLayout L;
    std::memset(&L, 0, sizeof(L));
    const bool bwd = mode == LG_BWD, enc = mode == LG_ENC;
    const bool dae = p->kind == PSNODE_DAE;
    const int H = p->X, E = p->event_idx ? p->E : 0;
    L.H = H;
    L.KZV = p->Z + p->V;
    L.S = p->X + p->Z + p->V + p->I;
    L.bwd = bwd ? 1 : 0;
    L.enc = enc ? 1 : 0;
    L.nmat = bwd ? NMAT_BWD : NMAT_FWD;
    const int k2 = dae ? 2 * H : H;
    const int rows[NMAT] = {H, H, H, H, H, H, H, H, H, H, H, H, H, H, H, H, L.S, H, 128, H, 128};
    const int cols[NMAT] = {H, L.KZV, H, H, H, L.KZV, H, L.S, L.S, H, H, H, H, H, k2, k2, k2, H, H, H, H};
    int64_t o = 64;
    L.err = 0;
    for (int i = 0; i < NMAT; i++) {
        L.mrows[i] = rows[i]; L.mcols[i] = cols[i];
        const bool dec = enc && (i == M_XD1 || i == M_XD2 || (dae && (i == M_ID1 || i == M_ID2)));
        if (i >= L.nmat && !dec) continue;
        const bool used = dec || dae || (i == M_DE1X || i == M_DE1ZV || i == M_DE2 || i == M_C_DE || i == M_T_DE2 || i == M_T_DE1X || i == M_T_DZ || i == M_T_DA0);
        if (!used) continue;
        L.wts_hi[i] = o; o += al((int64_t)rows[i] * cols[i]);
        L.wts_lo[i] = o; o += al((int64_t)rows[i] * cols[i]);
    }
    const int64_t BH = (int64_t)p->B * H;
    L.c_de = o; o += al(BH);
    L.c_ae = o; o += al(BH);
    // (the reverse pass overwrites pre[r] with d pre[r] in place)
    // hoisted layer-1 halves live in a time-chunk scratch (rc grid rows) in every mode: projections, steps and -- in the reverse pass --
    // the chunk's share of the hoisted gradients run chunk by chunk, so the workspace is O(chunk), not O(T)
    L.rc = chunk_rows > 0 ? chunk_rows : lg_chunk_rows();
    if (L.rc > p->T) L.rc = p->T;
    L.rows_de = L.rc;
    L.rows_ae = dae ? L.rc + 1 : 0;
    L.pj_de = o; o += al((int64_t)(E > 0 ? E : 1) * BH);
    L.pj_ae = o; o += al((int64_t)(E > 0 ? E : 1) * BH);
    if (enc) {
        L.encb_de = o; o += al(H);
        L.encb_ae = o; o += al(H);
        L.xs = o; o += al((int64_t)(L.rc + 1) * BH);
        L.is = o; o += al(dae ? (int64_t)(L.rc + 1) * BH : 64);
        L.dtmp = o; o += al((int64_t)(L.rc + 1) * BH);
    }
    L.pre_de = o; o += al(L.rows_de * BH);
    L.pre_ae = o; o += al(L.rows_ae * BH);
    L.x0 = o; o += al(BH); L.k1 = o; o += al(BH); L.k2 = o; o += al(BH); L.k3 = o; o += al(BH);
    L.ycur = o; o += al(BH); L.a1 = o; o += al(BH); L.hbuf = o; o += al(BH); L.icur = o; o += al(BH); L.G = o; o += al(BH);
    if (bwd) {
        L.ring = lg_ring_depth();
        if (p->T - 1 < L.ring) L.ring = p->T > 1 ? p->T - 1 : 1;
        L.nst = psw_nstages(p->method);
        L.depth = lg_overlap_wgrad(p) ? 2 * L.ring : L.ring;
        L.nbufs = BUF_SINGLES + L.depth * (4 * L.nst + 5);
        L.dpj_de = o; o += al((int64_t)(E > 0 ? E : 1) * BH);
        L.dpj_ae = o; o += al((int64_t)(E > 0 ? E : 1) * BH);
        L.bufs = o; o += (int64_t)L.nbufs * al(BH);
        L.slabs = o; o += al(lg_wgrad_slab_floats(H, L.S));
        L.slabs2 = o; o += al(lg_wgrad_slab_floats(H, H));
    }
    L.total = o;
    return L;
}

bool view_ok(const float* p, int64_t s0, int64_t s1) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (s0 & 3) == 0 && (s1 & 3) == 0; }

// dst[m * ldd + k] (m < M, k < K) = split_tf32(A[m][k]),  A[m][k] = W[m * ldw + col0 + k] (+ sgn * W[m * ldw + col1 + k] if col1 >= 0), or with
// `transpose` the same expression with m and k exchanged on the right-hand side (A = the transposed block)
__global__ void psn_lg_prep2_kernel(const float* __restrict__ W, int ldw, int col0, int col1, float sgn, int transpose, int M, int K, int ldd,
                                    float* __restrict__ hi, float* __restrict__ lo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * K) return;
    const int m = idx / K, k = idx - m * K;
    const int r = transpose ? k : m, c = transpose ? m : k;
    float w = __ldg(W + (int64_t)r * ldw + col0 + c);
    if (col1 >= 0) w = fmaf(sgn, __ldg(W + (int64_t)r * ldw + col1 + c), w);
    float h, l;
    split_tf32(w, h, l);
    hi[(int64_t)m * ldd + k] = h;
    lo[(int64_t)m * ldd + k] = l;
}

// encoder fusion: out[m][dcol + k] = split_tf32(sum_h F[m][h] * E2[h][k]),  F[m][h] = W[m * ldw + col0 + h] (+ W[m * ldw + col1 + h]); and the
// matching bias  pb[m] (+)= sum_h F[m][h] * e2[h]   (the encoder's second Linear folded into the held-input half of layer 1)
__global__ void psn_lg_fold_enc_kernel(const float* __restrict__ W, int ldw, int col0, int col1, const float* __restrict__ E2, const float* __restrict__ e2,
                                       int H, int ldd, int dcol, float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ pb, int pb_acc) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * H) return;
    const int m = idx / H, k = idx - m * H;
    float acc = 0.0f, accb = 0.0f;
    for (int h = 0; h < H; h++) {
        float f = __ldg(W + (int64_t)m * ldw + col0 + h);
        if (col1 >= 0) f += __ldg(W + (int64_t)m * ldw + col1 + h);
        acc = fmaf(f, __ldg(E2 + (int64_t)h * H + k), acc);
        if (k == 0) accb = fmaf(f, __ldg(e2 + h), accb);
    }
    float hh, ll;
    split_tf32(acc, hh, ll);
    hi[(int64_t)m * ldd + dcol + k] = hh;
    lo[(int64_t)m * ldd + dcol + k] = ll;
    if (k == 0) pb[m] = pb_acc ? pb[m] + accb : accb;
}

// Everything the forward pass and the reverse pass share: prepared planes, tensor maps, hoisted projections, launch helpers.
struct LgCtx {
    const psnode_problem* p;
    LgLayout L;
    float* w;
    int* err;
    cudaStream_t stream;
    bool dae;
    int H, B, T, E, S, nbt, nstages;
    int64_t BH;
    CUtensorMap mw_hi[NMAT], mw_lo[NMAT], m_z, m_v, m_zj, m_vj, m_y, m_a1, m_h, m_i;
    int dbg_cta, rotate;

    LgParams base() const {
        LgParams q;
        std::memset(&q, 0, sizeof(q));
        q.N = B; q.R = 1; q.nbt = nbt; q.nsrc = 1; q.kchunks = H / 32; q.mode = LG_PLAIN;
        q.add1_ld = H; q.add2_ld = H; q.out_ld = H; q.st_ld = H;
        q.err = err;
        q.dbg = dbg_cta; q.rotate = rotate;
        return q;
    }
    int launch(int which, const CUtensorMap& b0, const CUtensorMap& b1, const LgParams& q, const char* name) const {
        return lg_launch_gemm(mw_hi[which], mw_lo[which], b0, b1, q, L.mrows[which] / TM, stream, name);
    }
    void prep(int which, const float* W, int ldw, int col0, int col1, float sgn, int transpose, int M, int K, int dcol) const {
        const int n = M * K;
        psn_lg_prep2_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, ldw, col0, col1, sgn, transpose, M, K, L.mcols[which],
                                                                 w + L.wts_hi[which] + dcol, w + L.wts_lo[which] + dcol);
        psn_count_launch("psn_lg_prep_kernel");
    }
    // hoisted half of R series rows starting at row r_first: out[r] = [F_z | F_v] [z[r_first + r]; v[r_first + r]] + cadd
    void project(int which, const CUtensorMap& bz, const CUtensorMap& bv, int r_first, int R, const float* cadd, float* out, const char* name) const {
        if (R <= 0) return;
        LgParams q = base();
        const bool has_z = p->Z > 0;
        q.R = R; q.nsrc = (dae && has_z) ? 2 : 1; q.kchunks = H / 32;
        q.b_r0 = r_first;
        q.add1 = cadd; q.add1_sr = 0; q.add1_ld = H;
        q.out = out; q.out_sr = BH; q.out_ld = H;
        launch(which, has_z ? bz : bv, dae ? bv : bz, q, name);
    }
};

// weights -> planes, per-trajectory constants, tensor maps, hoisted layer-1 halves over the whole series
int lg_setup(LgCtx& c, const psnode_problem* p, int mode, void* ws, int64_t ws_bytes, cudaStream_t stream, const psnode_codec* codec = nullptr) {
    const bool bwd = mode == LG_BWD, enc = mode == LG_ENC;
    c.p = p;
    c.L = lg_layout(p, mode, codec ? codec->chunk_rows : 0);
    const LgLayout& L = c.L;
    if (ws == nullptr || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    float* w = c.w = static_cast<float*>(ws);
    c.err = reinterpret_cast<int*>(w + L.err);
    c.stream = stream;
    const bool dae = c.dae = p->kind == PSNODE_DAE;
    const int H = c.H = L.H, B = c.B = p->B, T = c.T = p->T, E = c.E = p->event_idx ? p->E : 0;
    const int S = c.S = L.S;
    c.BH = (int64_t)B * H;
    c.nstages = psw_nstages(p->method);
    c.nbt = (B + TN - 1) / TN;
    static const int dbg_cta = std::getenv("PSNODE_LG_DBG") ? std::atoi(std::getenv("PSNODE_LG_DBG")) : 0;
    static const int rotate = std::getenv("PSNODE_LG_ROTATE") ? std::atoi(std::getenv("PSNODE_LG_ROTATE")) : 1;
    c.dbg_cta = dbg_cta; c.rotate = rotate;
    PSN_CUDA(cudaMemsetAsync(c.err, 0, 256, stream));
    static const bool trace = std::getenv("PSNODE_LG_TRACE") && std::atoi(std::getenv("PSNODE_LG_TRACE")) != 0;
    if (trace) {
        void *pmin = nullptr, *pmax = nullptr;
        cudaGetSymbolAddress(&pmin, g_lg_tmin); cudaGetSymbolAddress(&pmax, g_lg_tmax);
        cudaMemsetAsync(pmin, 0xFF, sizeof(unsigned long long) * LG_TRACE_N * 2, stream);
        cudaMemsetAsync(pmax, 0, sizeof(unsigned long long) * LG_TRACE_N * 2, stream);
        g_lg_trace_next = 0;
    }
    // ---- weights: folded, split into tf32 hi / lo planes ----
    const float* W1 = p->de.W[0];
    const int ld1 = 3 * S;
    const int X = p->X, Z = p->Z, V = p->V;
    const float* A1 = dae ? p->ae.W[0] : nullptr;
    const int lda = S + X + Z + V;
    c.prep(M_DE1X, W1, ld1, S, 2 * S, 1.0f, 0, H, H, 0);                              // F_x = (W_b + W_c)[:, 0:X]
    if (!enc) c.prep(M_DE1ZV, W1, ld1, S + X, 2 * S + X, 1.0f, 0, H, L.KZV, 0);       // [F_z | F_v]
    c.prep(M_DE2, p->de.W[1], H, 0, -1, 0.0f, 0, H, H, 0);
    if (dae) {
        c.prep(M_DE1I, W1, ld1, S + X + Z + V, 2 * S + X + Z + V, 1.0f, 0, H, H, 0);   // F_i
        c.prep(M_AE1X, A1, lda, S, -1, 0.0f, 0, H, H, 0);
        if (!enc) c.prep(M_AE1ZV, A1, lda, S + X, -1, 0.0f, 0, H, L.KZV, 0);
        c.prep(M_AE2, p->ae.W[1], H, 0, -1, 0.0f, 0, H, H, 0);
    }
    if (enc) {
        // [F_z E2z | F_v E2v] and [A1z E2z | A1v E2v]: the held-input halves act on the encoders' HIDDEN layers; decoder planes
        auto fold = [&](int which, const float* W, int ldw, int col0, int col1, const psnode_mlp& en, int dcol, float* pb, int pb_acc) {
            psn_lg_fold_enc_kernel<<<(H * H + 255) / 256, 256, 0, stream>>>(W, ldw, col0, col1, en.W[1], en.b[1], H, L.mcols[which], dcol,
                                                                            w + L.wts_hi[which], w + L.wts_lo[which], pb, pb_acc);
            psn_count_launch("psn_lg_fold_enc_kernel");
        };
        fold(M_DE1ZV, W1, ld1, S + X, 2 * S + X, codec->z_enc, 0, w + L.encb_de, 0);
        if (dae) {
            fold(M_DE1ZV, W1, ld1, S + X + Z, 2 * S + X + Z, codec->v_enc, H, w + L.encb_de, 1);
            fold(M_AE1ZV, A1, lda, S + X, -1, codec->z_enc, 0, w + L.encb_ae, 0);
            fold(M_AE1ZV, A1, lda, S + X + Z, -1, codec->v_enc, H, w + L.encb_ae, 1);
        }
        c.prep(M_XD1, codec->x_dec.W[0], H, 0, -1, 0.0f, 0, H, H, 0);
        PSN_CUDA(cudaMemsetAsync(w + L.wts_hi[M_XD2], 0, (size_t)128 * H * 4, stream));
        PSN_CUDA(cudaMemsetAsync(w + L.wts_lo[M_XD2], 0, (size_t)128 * H * 4, stream));
        c.prep(M_XD2, codec->x_dec.W[1], H, 0, -1, 0.0f, 0, codec->XR, H, 0);
        if (dae) {
            c.prep(M_ID1, codec->i_dec.W[0], H, 0, -1, 0.0f, 0, H, H, 0);
            PSN_CUDA(cudaMemsetAsync(w + L.wts_hi[M_ID2], 0, (size_t)128 * H * 4, stream));
            PSN_CUDA(cudaMemsetAsync(w + L.wts_lo[M_ID2], 0, (size_t)128 * H * 4, stream));
            c.prep(M_ID2, codec->i_dec.W[1], H, 0, -1, 0.0f, 0, codec->IR, H, 0);
        }
    }
    // per-trajectory layer-1 constants c = (W_a - W_b) a0 + b1 (and A1a a0 + ab1): a K = S GEMM on the same kernel when all_initial can be
    // a TMA operand, else a CUDA-core kernel
    const bool a0_tma = view_ok(p->a0, p->a0_sb, 0);
    if (a0_tma) {
        c.prep(M_C_DE, W1, ld1, 0, S, -1.0f, 0, H, S, 0);
        if (dae) c.prep(M_C_AE, A1, lda, 0, -1, 0.0f, 0, H, S, 0);
    } else {
        if (dae) {
            psn_lg_const_kernel<<<(B + 7) / 8, H, 8 * S * 4, stream>>>(A1, lda, -1, p->ae.b[0], p->a0, p->a0_sb, S, B, H, w + L.c_ae);
            psn_count_launch("psn_lg_const_kernel");
        }
        psn_lg_const_kernel<<<(B + 7) / 8, H, 8 * S * 4, stream>>>(W1, ld1, S, p->de.b[0], p->a0, p->a0_sb, S, B, H, w + L.c_de);
        psn_count_launch("psn_lg_const_kernel");
    }
    if (bwd) {
        c.prep(M_T_DE2, p->de.W[1], H, 0, -1, 0.0f, 1, H, H, 0);                      // W2^T
        c.prep(M_T_DE1X, W1, ld1, S, 2 * S, 1.0f, 1, H, H, 0);                        // F_x^T
        if (Z > 0) c.prep(M_T_DZ, W1, ld1, S + X, 2 * S + X, 1.0f, 1, H, H, 0);       // [F_z^T | A1z^T]
        c.prep(M_T_DA0, W1, ld1, 0, S, -1.0f, 1, S, H, 0);                            // [(W_a - W_b)^T | A1a^T]
        if (dae) {
            c.prep(M_T_DE1I, W1, ld1, S + X + Z + V, 2 * S + X + Z + V, 1.0f, 1, H, H, 0);
            c.prep(M_T_AE2, p->ae.W[1], H, 0, -1, 0.0f, 1, H, H, 0);
            c.prep(M_T_AE1X, A1, lda, S, -1, 0.0f, 1, H, H, 0);
            if (Z > 0) c.prep(M_T_DZ, A1, lda, S + X, -1, 0.0f, 1, H, H, H);
            c.prep(M_T_DV, W1, ld1, S + X + Z, 2 * S + X + Z, 1.0f, 1, H, H, 0);      // [F_v^T | A1v^T]
            c.prep(M_T_DV, A1, lda, S + X + Z, -1, 0.0f, 1, H, H, H);
            c.prep(M_T_DA0, A1, lda, 0, -1, 0.0f, 1, S, H, H);
        }
    }
    PSN_CUDA(cudaGetLastError());

    // ---- tensor maps (once per call) ----
    bool ok = true;
    for (int i = 0; i < NMAT; i++) {
        if (L.wts_hi[i] == 0) continue;
        ok = ok && lg_make_map(&c.mw_hi[i], w + L.wts_hi[i], L.mcols[i], L.mrows[i], L.mcols[i], 1, 0) &&
             lg_make_map(&c.mw_lo[i], w + L.wts_lo[i], L.mcols[i], L.mrows[i], L.mcols[i], 1, 0);
    }
    if (!enc) {
        if (Z > 0) ok = ok && lg_make_map(&c.m_z, p->z.p, H, B, p->z.sb, T, p->z.st);
        if (dae) ok = ok && lg_make_map(&c.m_v, p->v.p, H, B, p->v.sb, T, p->v.st);
        if (E > 0) {
            if (Z > 0) ok = ok && lg_make_map(&c.m_zj, p->z_jump, H, B, p->zj_sb, E, p->zj_se);
            if (dae) ok = ok && lg_make_map(&c.m_vj, p->v_jump, H, B, p->vj_sb, E, p->vj_se);
        }
    }
    ok = ok && lg_make_map(&c.m_y, w + L.ycur, H, B, H, 1, 0) && lg_make_map(&c.m_a1, w + L.a1, H, B, H, 1, 0);
    if (dae) ok = ok && lg_make_map(&c.m_h, w + L.hbuf, H, B, H, 1, 0) && lg_make_map(&c.m_i, w + L.icur, H, B, H, 1, 0);
    if (!ok) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer path)");
    if (lg_prepare_kernels() != PSNODE_OK) return PSNODE_ECUDA;
    if (a0_tma) {
        CUtensorMap m_a0;
        if (!lg_make_map(&m_a0, p->a0, S, B, p->a0_sb, 1, 0)) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (all_initial)");
        LgParams q = c.base();
        q.kchunks = S / 32;
        q.bias = p->de.b[0];
        q.out = w + L.c_de;
        c.launch(M_C_DE, m_a0, m_a0, q, "psn_lg_gemm_kernel<c_de>");
        if (dae) {
            q.bias = p->ae.b[0];
            q.out = w + L.c_ae;
            c.launch(M_C_AE, m_a0, m_a0, q, "psn_lg_gemm_kernel<c_ae>");
        }
    }

    if (enc) return PSNODE_OK;                                   // the encoded entry generates its projections from the raw series
    // ---- hoisted layer-1 halves of the event rows (the grid rows are projected chunk by chunk by the callers) ----
    if (E > 0) {
        c.project(M_DE1ZV, c.m_zj, c.m_vj, 0, E, w + L.c_de, w + L.pj_de, "psn_lg_gemm_kernel<pre_de_jump>");
        if (dae) c.project(M_AE1ZV, c.m_zj, c.m_vj, 0, E, w + L.c_ae, w + L.pj_ae, "psn_lg_gemm_kernel<pre_ae_jump>");
    }
    return PSNODE_OK;
}

// ---- small fp32 kernels of the reverse pass (element-wise over one (B, H) buffer) --------------------------------------------
// start of the sweep: gx = dL/dx_sol[T-1], gi = dL/di_sol[T-1], acc_gi = gi
__global__ void psn_lg_bwd_init_kernel(int B, int H, const float* __restrict__ gx_up, int64_t gx_sb, const float* __restrict__ gi_up, int64_t gi_sb,
                                       float* __restrict__ gx, float* __restrict__ gi, float* __restrict__ acc_gi) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, m = idx - b * H;
    gx[idx] = __ldg(gx_up + (int64_t)b * gx_sb + m);
    if (gi_up) {
        const float g = __ldg(gi_up + (int64_t)b * gi_sb + m);
        gi[idx] = g;
        acc_gi[idx] = g;
    }
}
// head of step j: stage-0 input y0 = x_sol[j-1], held i0 = i_sol[j-1], slope adjoint of the last stage dk = gx * dt * c
__global__ void psn_lg_bwd_head_kernel(int B, int H, const float* __restrict__ xprev, int64_t x_sb, const float* __restrict__ iprev, int64_t i_sb,
                                       float* __restrict__ y0, float* __restrict__ i0, const float* __restrict__ gx, const float* __restrict__ tcur,
                                       const float* __restrict__ tprev, int64_t t_sb, float c, float* __restrict__ dk, float* __restrict__ acc_dk) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, m = idx - b * H;
    y0[idx] = __ldg(xprev + (int64_t)b * x_sb + m);
    if (iprev) i0[idx] = __ldg(iprev + (int64_t)b * i_sb + m);
    const float dt = __fsub_rn(__ldg(tcur + (int64_t)b * t_sb), __ldg(tprev + (int64_t)b * t_sb));
    const float v = gx[idx] * (dt * c);
    dk[idx] = v;
    acc_dk[idx] += v;
}
// tail of step j: d pre_de row (the row of the event when one fired at j-1, and then zeros in the grid row), and the adjoint of
// i_{j-1}: dL/di_sol[j-1] (+ F_i^T dsum when i_{j-1} itself was the held i0, i.e. no event re-evaluated it)
__global__ void psn_lg_bwd_tail_kernel(int B, int H, const int32_t* __restrict__ ev, int ev_j, const float* __restrict__ dsum, float* __restrict__ dpre_row,
                                       float* __restrict__ dpre_jump, int64_t jump_sr, const float* __restrict__ gi0, const float* __restrict__ gi_up,
                                       int64_t gi_sb, float* __restrict__ gi_next, float* __restrict__ acc_gi) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, m = idx - b * H;
    const int evk = ev ? __ldg(ev + ev_j) : -1;
    const float d = dsum[idx];
    if (evk < 0) dpre_row[idx] = d;
    else { dpre_row[idx] = 0.0f; dpre_jump[(int64_t)evk * jump_sr + idx] = d; }
    if (gi_next) {
        const float up = __ldg(gi_up + (int64_t)b * gi_sb + m), g0 = gi0[idx];
        gi_next[idx] = evk < 0 ? up + g0 : up;
        acc_gi[idx] += up + g0;
    }
}
// one grid row of the fused masked-MSE upstream gradient (psnode_adjoint.fuse_x / fuse_i): up[b][m] = dL/dsol[j][b][m]
__global__ void psn_lg_fuse_row_kernel(int B, int H, int j, PsnFuse fx, float* __restrict__ upx, PsnFuse fi, float* __restrict__ upi) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    const int b = idx / H, m = idx - b * H;
    if (upx) upx[idx] = psn_fuse_grad(fx, psn_fuse_scale(fx), j, b, m);
    if (upi) upi[idx] = psn_fuse_grad(fi, psn_fuse_scale(fi), j, b, m);
}
// out[m] = sum_b src[b][m]: block = 32 features, 32 row lanes
__global__ void __launch_bounds__(1024) psn_lg_colsum_kernel(const float* __restrict__ src, int B, int H, float* __restrict__ out) {
    __shared__ float part[32][33];
    const int c = threadIdx.x & 31, r = threadIdx.x >> 5, m = blockIdx.x * 32 + c;
    float s = 0.0f;
    for (int b = r; b < B; b += 32) s += src[(int64_t)b * H + m];
    part[r][c] = s;
    __syncthreads();
    if (r == 0) {
        float t = 0.0f;
        for (int i = 0; i < 32; i++) t += part[i][c];
        out[m] = t;
    }
}
// dW1 = [dW_a | dW_b | dW_c] with dW_a = dc (x) a0 (already in place), dW_c = dF (already in place), dW_b = dF - dW_a
__global__ void psn_lg_unfold_w1_kernel(float* __restrict__ dW1, int H, int S) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * S) return;
    const int m = idx / S, k = idx - m * S;
    float* row = dW1 + (int64_t)m * 3 * S;
    row[S + k] = row[2 * S + k] - row[k];
}
__global__ void psn_lg_copy_rows_kernel(int B, int H, const float* __restrict__ src, float* __restrict__ dst, int64_t dst_sb) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * H) return;
    dst[(int64_t)(idx / H) * dst_sb + (idx % H)] = src[idx];
}
__global__ void psn_lg_zero_rows_kernel(int R, int B, int H, float* __restrict__ dst, int64_t dst_sr, int64_t dst_sb) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)R * B * H) return;
    const int64_t r = idx / ((int64_t)B * H), rem = idx - r * (int64_t)B * H;
    dst[r * dst_sr + (rem / H) * dst_sb + (rem % H)] = 0.0f;
}

}  // namespace

bool psn_lg_supports(const psnode_problem* p) {
    if (p->teacher_x || p->teacher_i) return false;
    const int H = p->X;
    if (H != 128 && H != 256) return false;
    const bool dae = p->kind == PSNODE_DAE;
    if (p->Z != H && !(dae && p->Z == 0)) return false;            // Z = 0: the z_dim == 0 variant of DAE_02 (neural_01_DAE_02_direct_encode.py:73, :90)
    if (dae && (p->V != H || p->I != H)) return false;
    const int S = p->X + p->Z + p->V + p->I;
    if (p->de.n_layers != 2 || p->de.in_dim[0] != 3 * S || p->de.out_dim[0] != H || p->de.out_dim[1] != H) return false;
    if (dae && (p->ae.n_layers != 2 || p->ae.in_dim[0] != S + p->X + p->Z + p->V || p->ae.out_dim[0] != H || p->ae.out_dim[1] != H)) return false;
    if (p->Z > 0 && !view_ok(p->z.p, p->z.st, p->z.sb)) return false;
    if (dae && !view_ok(p->v.p, p->v.st, p->v.sb)) return false;
    if (p->event_idx) {
        if (p->Z > 0 && !view_ok(p->z_jump, p->zj_sb, p->zj_se)) return false;
        if (dae && !view_ok(p->v_jump, p->vj_sb, p->vj_se)) return false;
    }
    if ((p->x_sol.sb & 3) || (p->x_sol.st & 3)) return false;
    return lg_encode_fn() != nullptr;
}

int64_t psn_lg_forward_workspace(const psnode_problem* p) { return lg_layout(p, LG_FWD).total * 4; }

namespace {
// one grid step of the forward pass (shared by psn_lg_forward and the encoded entry)
struct LgStepIO {
    const float* pre_de_row;        // hoisted DE half of grid row j-1
    const float* pre_de_jump;       // its event rows (row stride B * H)
    const float* pre_ae_row;        // hoisted AE half of grid row j
    const float* pre_ae_jump;
    float* xrow; int64_t xld;       // where x_j goes (row stride xld)
    float* irow; int64_t ild;       // where i_j goes
};
// algebraic evaluation i = ae(x, z, v) on the current state (ycur holds x at step boundaries)
void lg_ae_eval(const LgCtx& c, const float* pre_row, const float* pre_jump, float* irow, int64_t ild, bool event_only, int ev_j) {
    const psnode_problem* p = c.p;
    float* w = c.w;
    LgParams q = c.base();
    q.mode = LG_HIDDEN;
    q.add1 = pre_row; q.add1_sr = 0;
    if (event_only) {
        q.ev = p->event_idx; q.ev_j = ev_j; q.skip_unless_event = 1;
        q.add1_jump = pre_jump; q.add1_jump_sr = c.BH;
    }
    q.out = w + c.L.hbuf;
    c.launch(M_AE1X, c.m_y, c.m_y, q, event_only ? "psn_lg_gemm_kernel<ae1,event>" : "psn_lg_gemm_kernel<ae1>");
    LgParams q2 = c.base();
    q2.bias = p->ae.b[1];
    if (event_only) { q2.ev = p->event_idx; q2.ev_j = ev_j; q2.skip_unless_event = 1; }
    q2.out = w + c.L.icur;
    if (irow) { q2.out2 = irow; q2.out2_ld = ild; }
    c.launch(M_AE2, c.m_h, c.m_h, q2, event_only ? "psn_lg_gemm_kernel<ae2,event>" : "psn_lg_gemm_kernel<ae2>");
}
void lg_step(const LgCtx& c, int j, const LgStepIO& io) {
    const psnode_problem* p = c.p;
    float* w = c.w;
    const LgLayout& L = c.L;
    if (c.dae) {
        if (c.E > 0) lg_ae_eval(c, nullptr, io.pre_ae_jump, nullptr, 0, true, j - 1);   // event: i_0 re-evaluated with the jumped inputs (:108-110)
        LgParams qg = c.base();                                                          // G = F_i i0
        qg.out = w + L.G;
        c.launch(M_DE1I, c.m_i, c.m_i, qg, "psn_lg_gemm_kernel<de1_i>");
    }
    for (int e = 0; e < c.nstages; e++) {
        LgParams q1 = c.base();
        q1.mode = LG_HIDDEN;
        q1.add1 = io.pre_de_row;
        if (c.E > 0) { q1.ev = p->event_idx; q1.ev_j = j - 1; q1.add1_jump = io.pre_de_jump; q1.add1_jump_sr = c.BH; }
        if (c.dae) q1.add2 = w + L.G;
        q1.out = w + L.a1;
        c.launch(M_DE1X, c.m_y, c.m_y, q1, "psn_lg_gemm_kernel<de1>");
        LgParams q2 = c.base();
        q2.mode = LG_RK;
        q2.bias = p->de.b[1];
        q2.method = p->method; q2.stage = e;
        q2.x0 = w + L.x0; q2.k1 = w + L.k1; q2.k2 = w + L.k2; q2.k3 = w + L.k3;
        q2.t_cur = p->t.p + (int64_t)j * p->t.st; q2.t_prev = p->t.p + (int64_t)(j - 1) * p->t.st; q2.t_sb = p->t.sb;
        q2.out = w + L.ycur;
        q2.out2 = io.xrow; q2.out2_ld = io.xld;
        c.launch(M_DE2, c.m_a1, c.m_a1, q2, "psn_lg_gemm_kernel<de2,rk>");
    }
    if (c.dae) lg_ae_eval(c, io.pre_ae_row, nullptr, io.irow, io.ild, false, 0);         // i_j = ae(x_j, z[j], v[j])  (:121)
}
void lg_dump_trace(const LgCtx& c) {
    if (g_lg_trace_next < 0) return;         // debugging aid only: synchronises
    static unsigned long long tmin[LG_TRACE_N][2], tmax[LG_TRACE_N][2];
    cudaStreamSynchronize(c.stream);
    cudaMemcpyFromSymbol(tmin, g_lg_tmin, sizeof(tmin));
    cudaMemcpyFromSymbol(tmax, g_lg_tmax, sizeof(tmax));
    const int n = g_lg_trace_next;
    const int from = n > 120 ? n / 2 : 1, to = from + 60 < n ? from + 60 : n;
    std::fprintf(stderr, "psn_lg trace (ns): launch | first entry - prev end | first wait-return - prev end | last wait-return - prev end | duration (last exit - first wait-return)\n");
    double g0 = 0, g1 = 0, g2 = 0, d = 0; int cnt = 0;
    for (int k = from; k < to; k++) {
        const long long pe = (long long)tmax[k - 1][0];
        const long long a = (long long)tmin[k][0] - pe, b = (long long)tmin[k][1] - pe, cc = (long long)tmax[k][1] - pe, dur = (long long)tmax[k][0] - (long long)tmin[k][1];
        std::fprintf(stderr, "  %4d %-40s %8lld %8lld %8lld %8lld\n", k, g_lg_trace_name[k], a, b, cc, dur);
        g0 += a; g1 += b; g2 += cc; d += dur; cnt++;
    }
    if (cnt) std::fprintf(stderr, "  mean over %d launches: entry gap %.0f, wait gap %.0f .. %.0f, duration %.0f ns\n", cnt, g0 / cnt, g1 / cnt, g2 / cnt, d / cnt);
    g_lg_trace_next = -1;
}
void lg_dump_stamps(const LgCtx& c) {
    lg_dump_trace(c);
    if (!c.dbg_cta) return;                  // debugging aid only: synchronises
    long long h[64];
    cudaStreamSynchronize(c.stream);
    cudaMemcpyFromSymbol(h, g_lg_dbg, sizeof(h));
    std::fprintf(stderr, "psn_lg stamps (cycles since kernel start, last launch, CTA %d):", c.dbg_cta - 1);
    for (int i = 1; i < (int)h[63] && i < 32; i++) std::fprintf(stderr, " %lld", h[i] - h[0]);
    std::fprintf(stderr, "\n");
}
}  // namespace

int psn_lg_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    LgCtx c;
    const int st = lg_setup(c, p, LG_FWD, ws, ws_bytes, stream);
    if (st != PSNODE_OK) return st;
    const LgLayout& L = c.L;
    float* w = c.w;
    const bool dae = c.dae;
    const int H = c.H, B = c.B, T = c.T;
    const int64_t BH = c.BH;
    // ---- initial state ----
    {
        const float* src = dae ? p->x_init : p->x.p;
        const int64_t sb = dae ? p->x_init_sb : p->x.sb;
        psn_lg_init_kernel<<<(int)((BH + 255) / 256), 256, 0, stream>>>(src, sb, B, H, w + L.x0, w + L.ycur, p->x_sol.p, p->x_sol.sb);
        psn_count_launch("psn_lg_init_kernel");
    }
    const int RC = L.rc;
    if (dae) {                                                                 // i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95)
        c.project(M_AE1ZV, c.m_z, c.m_v, 0, 1, w + L.c_ae, w + L.pre_ae, "psn_lg_gemm_kernel<pre_ae>");
        lg_ae_eval(c, w + L.pre_ae, nullptr, p->i_sol.p, p->i_sol.sb, false, 0);
    }
    // ---- time chunks: the hoisted halves of rc grid rows at a time, then the chunk's steps j = r0 + 1 .. r1 ----
    for (int r0 = 0; r0 < T - 1; r0 += RC) {
        const int r1 = r0 + RC < T - 1 ? r0 + RC : T - 1, n = r1 - r0;
        c.project(M_DE1ZV, c.m_z, c.m_v, r0, n, w + L.c_de, w + L.pre_de, "psn_lg_gemm_kernel<pre_de>");
        if (dae) c.project(M_AE1ZV, c.m_z, c.m_v, r0 + 1, n, w + L.c_ae, w + L.pre_ae + BH, "psn_lg_gemm_kernel<pre_ae>");
        for (int j = r0 + 1; j <= r1; j++) {
            const int k = j - r0;
            LgStepIO io;
            io.pre_de_row = w + L.pre_de + (int64_t)(k - 1) * BH; io.pre_de_jump = w + L.pj_de;
            io.pre_ae_row = w + L.pre_ae + (int64_t)k * BH; io.pre_ae_jump = w + L.pj_ae;
            io.xrow = p->x_sol.p + (int64_t)j * p->x_sol.st; io.xld = p->x_sol.sb;
            io.irow = dae ? p->i_sol.p + (int64_t)j * p->i_sol.st : nullptr; io.ild = p->i_sol.sb;
            lg_step(c, j, io);
        }
    }
    PSN_CUDA(cudaGetLastError());
    lg_dump_stamps(c);
    return PSNODE_OK;
}

// ---- encoded entry (psnode_forward_encoded): encoders fused into the hoisted projections, time chunks, decoders before the store ----
bool psn_lg_encoded_supports(const psnode_problem* p, const psnode_codec* cd) {
    if (!p || !cd) return false;
    const int H = p->X;
    const bool dae = p->kind == PSNODE_DAE;
    if (H != 128 && H != 256) return false;
    if (p->teacher_x || p->teacher_i || p->Z != H || (dae && (p->V != H || p->I != H)) || (!dae && (p->V || p->I))) return false;
    const int S = p->X + p->Z + p->V + p->I;
    if (p->de.n_layers != 2 || p->de.in_dim[0] != 3 * S || p->de.out_dim[0] != H || p->de.out_dim[1] != H) return false;
    if (dae && (p->ae.n_layers != 2 || p->ae.in_dim[0] != S + p->X + p->Z + p->V || p->ae.out_dim[0] != H || p->ae.out_dim[1] != H)) return false;
    auto enc_ok = [&](const psnode_mlp& m, int raw) {
        return m.n_layers == 2 && raw >= 1 && raw <= GEN_W && m.in_dim[0] == raw && m.out_dim[0] == H && m.in_dim[1] == H && m.out_dim[1] == H &&
               m.W[0] && m.b[0] && m.W[1] && m.b[1];
    };
    auto dec_ok = [&](const psnode_mlp& m, int raw) {
        return m.n_layers == 2 && raw >= 1 && raw <= 128 && m.in_dim[0] == H && m.out_dim[0] == H && m.in_dim[1] == H && m.out_dim[1] == raw &&
               m.W[0] && m.b[0] && m.W[1] && m.b[1];
    };
    if (!enc_ok(cd->z_enc, cd->ZR) || !cd->z_raw.p || !dec_ok(cd->x_dec, cd->XR) || !cd->x_out.p) return false;
    if (dae && (!enc_ok(cd->v_enc, cd->VR) || !cd->v_raw.p || !dec_ok(cd->i_dec, cd->IR) || !cd->i_out.p)) return false;
    if (p->event_idx && (p->E < 1 || !cd->zj_raw || (dae && !cd->vj_raw))) return false;
    if (!p->t.p || !p->a0 || (dae ? !p->x_init : !p->x.p)) return false;
    if (p->method < PSNODE_EULER || p->method > PSNODE_RK4 || p->B < 1 || p->T < 1) return false;
    return lg_encode_fn() != nullptr;
}

int64_t psn_lg_encoded_workspace(const psnode_problem* p, const psnode_codec* cd) { return lg_layout(p, LG_ENC, cd->chunk_rows).total * 4; }

int psn_lg_forward_encoded(const psnode_problem* p, const psnode_codec* cd, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    LgCtx c;
    const int st = lg_setup(c, p, LG_ENC, ws, ws_bytes, stream, cd);
    if (st != PSNODE_OK) return st;
    const LgLayout& L = c.L;
    float* w = c.w;
    const bool dae = c.dae;
    const int H = c.H, B = c.B, T = c.T, E = c.E, RC = L.rc;
    const int64_t BH = c.BH;
    CUtensorMap m_xs, m_is, m_tmp;
    bool ok = lg_make_map(&m_xs, w + L.xs, H, B, H, RC + 1, BH) && lg_make_map(&m_tmp, w + L.dtmp, H, B, H, RC + 1, BH);
    if (dae) ok = ok && lg_make_map(&m_is, w + L.is, H, B, H, RC + 1, BH);
    if (!ok) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (encoded entry)");
    // hoisted halves of `rows` grid rows starting at raw row r0 (or of the E event rows): the B operand is generated from the raw series
    auto project = [&](int which, int rows, const float* zraw, int64_t z_sr, int64_t z_sb, const float* vraw, int64_t v_sr, int64_t v_sb,
                       const float* cadd, const float* bias, float* out, const char* name) {
        if (rows <= 0) return;
        LgParams q = c.base();
        q.R = rows; q.nsrc = dae ? 2 : 1;
        q.add1 = cadd; q.add1_sr = 0;
        q.bias = bias;
        q.out = out; q.out_sr = BH;
        q.gen = dae ? 3 : 1;
        q.gen_raw[0] = zraw; q.gen_sr[0] = z_sr; q.gen_sb[0] = z_sb; q.gen_w[0] = cd->ZR; q.gen_W[0] = cd->z_enc.W[0]; q.gen_b[0] = cd->z_enc.b[0];
        if (dae) {
            q.gen_raw[1] = vraw; q.gen_sr[1] = v_sr; q.gen_sb[1] = v_sb; q.gen_w[1] = cd->VR; q.gen_W[1] = cd->v_enc.W[0]; q.gen_b[1] = cd->v_enc.b[0];
        }
        c.launch(which, c.m_y, c.m_y, q, name);          // (the tensor maps of the B operand are not used in generator mode)
    };
    // decoder over `rows` latent rows of a chunk scratch, first chunk row `row0`, into grid rows j0.. of the decoded output
    auto decode = [&](int d1, int d2, const psnode_mlp& dec, const CUtensorMap& m_src, int row0, int rows, const psnode_series_out& out, int j0, int width) {
        LgParams q = c.base();
        q.mode = LG_HIDDEN;
        q.R = rows; q.b_r0 = row0;
        q.bias = dec.b[0];
        q.out = w + L.dtmp; q.out_sr = BH;
        c.launch(d1, m_src, m_src, q, "psn_lg_gemm_kernel<dec1>");
        LgParams q2 = c.base();
        q2.R = rows;
        q2.bias = dec.b[1];
        q2.m_live = width;
        q2.out = out.p + (int64_t)j0 * out.st; q2.out_sr = out.st; q2.out_ld = out.sb;
        c.launch(d2, m_tmp, m_tmp, q2, "psn_lg_gemm_kernel<dec2>");
    };
    if (E > 0) {
        project(M_DE1ZV, E, cd->zj_raw, cd->zjr_se, cd->zjr_sb, cd->vj_raw, cd->vjr_se, cd->vjr_sb, w + L.c_de, w + L.encb_de, w + L.pj_de,
                "psn_lg_gemm_kernel<pre_de_jump,enc>");
        if (dae) project(M_AE1ZV, E, cd->zj_raw, cd->zjr_se, cd->zjr_sb, cd->vj_raw, cd->vjr_se, cd->vjr_sb, w + L.c_ae, w + L.encb_ae, w + L.pj_ae,
                         "psn_lg_gemm_kernel<pre_ae_jump,enc>");
    }
    // ---- initial state: latent x_0 into chunk row 0 ----
    {
        const float* src = dae ? p->x_init : p->x.p;
        const int64_t sb = dae ? p->x_init_sb : p->x.sb;
        psn_lg_init_kernel<<<(int)((BH + 255) / 256), 256, 0, stream>>>(src, sb, B, H, w + L.x0, w + L.ycur, w + L.xs, H);
        psn_count_launch("psn_lg_init_kernel");
    }
    auto raw_row = [](const psnode_series& s, int r) { return s.p + (int64_t)r * s.st; };
    if (dae) {                                                  // i_0 = ae(x_0, z[0], v[0])
        project(M_AE1ZV, 1, raw_row(cd->z_raw, 0), cd->z_raw.st, cd->z_raw.sb, raw_row(cd->v_raw, 0), cd->v_raw.st, cd->v_raw.sb, w + L.c_ae,
                w + L.encb_ae, w + L.pre_ae, "psn_lg_gemm_kernel<pre_ae,enc>");
        lg_ae_eval(c, w + L.pre_ae, nullptr, w + L.is, H, false, 0);
    }
    if (T == 1) {
        decode(M_XD1, M_XD2, cd->x_dec, m_xs, 0, 1, cd->x_out, 0, cd->XR);
        if (dae) decode(M_ID1, M_ID2, cd->i_dec, m_is, 0, 1, cd->i_out, 0, cd->IR);
    }
    // ---- time chunks: steps j = r0 + 1 .. r1 ----
    for (int r0 = 0; r0 < T - 1; r0 += RC) {
        const int r1 = r0 + RC < T - 1 ? r0 + RC : T - 1, n = r1 - r0;
        project(M_DE1ZV, n, raw_row(cd->z_raw, r0), cd->z_raw.st, cd->z_raw.sb, dae ? raw_row(cd->v_raw, r0) : nullptr, cd->v_raw.st, cd->v_raw.sb,
                w + L.c_de, w + L.encb_de, w + L.pre_de, "psn_lg_gemm_kernel<pre_de,enc>");
        if (dae) project(M_AE1ZV, n, raw_row(cd->z_raw, r0 + 1), cd->z_raw.st, cd->z_raw.sb, raw_row(cd->v_raw, r0 + 1), cd->v_raw.st, cd->v_raw.sb,
                         w + L.c_ae, w + L.encb_ae, w + L.pre_ae + BH, "psn_lg_gemm_kernel<pre_ae,enc>");
        for (int j = r0 + 1; j <= r1; j++) {
            const int k = j - r0;                               // chunk row of grid point j (row 0 = the chunk's start state)
            LgStepIO io;
            io.pre_de_row = w + L.pre_de + (int64_t)(k - 1) * BH; io.pre_de_jump = w + L.pj_de;
            io.pre_ae_row = w + L.pre_ae + (int64_t)k * BH; io.pre_ae_jump = w + L.pj_ae;
            io.xrow = w + L.xs + (int64_t)k * BH; io.xld = H;
            io.irow = dae ? w + L.is + (int64_t)k * BH : nullptr; io.ild = H;
            lg_step(c, j, io);
        }
        // decoders: rows 1..n of the chunk (and row 0 = the initial state in the first chunk)
        const int row0 = r0 == 0 ? 0 : 1, rows = r0 == 0 ? n + 1 : n, j0 = r0 == 0 ? 0 : r0 + 1;
        decode(M_XD1, M_XD2, cd->x_dec, m_xs, row0, rows, cd->x_out, j0, cd->XR);
        if (dae) decode(M_ID1, M_ID2, cd->i_dec, m_is, row0, rows, cd->i_out, j0, cd->IR);
    }
    PSN_CUDA(cudaGetLastError());
    lg_dump_stamps(c);
    return PSNODE_OK;
}

// ---- decoder alone (used by the wide kernels' encoded entry): out[r] = D2 ELU(D1 src[r] + d1) + d2 over R contiguous (B, H) rows ----
namespace {
constexpr int DEC_RC = 64;           // latent rows per pass (the hidden layer of the decoder lives in an (rc, B, H) scratch)
struct DecLayout { int64_t err, hi1, lo1, hi2, lo2, tmp, total; };
DecLayout dec_layout(int B, int H) {
    DecLayout L;
    int64_t o = 64;
    L.err = 0;
    L.hi1 = o; o += al((int64_t)H * H); L.lo1 = o; o += al((int64_t)H * H);
    L.hi2 = o; o += al((int64_t)128 * H); L.lo2 = o; o += al((int64_t)128 * H);
    L.tmp = o; o += al((int64_t)DEC_RC * B * H);
    L.total = o;
    return L;
}
}  // namespace
int64_t psn_lg_decode_workspace(int B, int H) { return dec_layout(B, H).total * 4; }

int psn_lg_decode(const psnode_mlp* dec, int width, const float* src, int R, int B, int H, const psnode_series_out* out, void* ws, int64_t ws_bytes,
                  cudaStream_t stream) {
    const DecLayout L = dec_layout(B, H);
    if (!ws || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    if ((H != 128 && H != 256) || width < 1 || width > 128 || dec->n_layers != 2 || !lg_encode_fn()) return PSNODE_EUNSUPPORTED;
    if (lg_prepare_kernels() != PSNODE_OK) return PSNODE_ECUDA;
    float* w = static_cast<float*>(ws);
    int* err = reinterpret_cast<int*>(w);
    const int64_t BH = (int64_t)B * H;
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    PSN_CUDA(cudaMemsetAsync(w + L.hi2, 0, (size_t)128 * H * 4, stream));
    PSN_CUDA(cudaMemsetAsync(w + L.lo2, 0, (size_t)128 * H * 4, stream));
    psn_lg_prep2_kernel<<<(H * H + 255) / 256, 256, 0, stream>>>(dec->W[0], H, 0, -1, 0.0f, 0, H, H, H, w + L.hi1, w + L.lo1);
    psn_lg_prep2_kernel<<<(width * H + 255) / 256, 256, 0, stream>>>(dec->W[1], H, 0, -1, 0.0f, 0, width, H, H, w + L.hi2, w + L.lo2);
    psn_count_launch("psn_lg_prep_kernel"); psn_count_launch("psn_lg_prep_kernel");
    CUtensorMap a1h, a1l, a2h, a2l, m_src, m_tmp;
    bool ok = lg_make_map(&a1h, w + L.hi1, H, H, H, 1, 0) && lg_make_map(&a1l, w + L.lo1, H, H, H, 1, 0) && lg_make_map(&a2h, w + L.hi2, H, 128, H, 1, 0) &&
              lg_make_map(&a2l, w + L.lo2, H, 128, H, 1, 0) && lg_make_map(&m_src, src, H, B, H, R, BH) && lg_make_map(&m_tmp, w + L.tmp, H, B, H, DEC_RC, BH);
    if (!ok) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (decoder)");
    LgParams base;
    std::memset(&base, 0, sizeof(base));
    base.N = B; base.R = 1; base.nbt = (B + TN - 1) / TN; base.nsrc = 1; base.kchunks = H / 32; base.mode = LG_PLAIN;
    base.add1_ld = H; base.add2_ld = H; base.out_ld = H; base.st_ld = H;
    base.err = err; base.rotate = 1;
    for (int r0 = 0; r0 < R; r0 += DEC_RC) {
        const int n = r0 + DEC_RC < R ? DEC_RC : R - r0;
        LgParams q = base;
        q.mode = LG_HIDDEN;
        q.R = n; q.b_r0 = r0;
        q.bias = dec->b[0];
        q.out = w + L.tmp; q.out_sr = BH;
        lg_launch_gemm(a1h, a1l, m_src, m_src, q, H / TM, stream, "psn_lg_gemm_kernel<dec1>");
        LgParams q2 = base;
        q2.R = n;
        q2.bias = dec->b[1];
        q2.m_live = width;
        q2.out = out->p + (int64_t)r0 * out->st; q2.out_sr = out->st; q2.out_ld = out->sb;
        lg_launch_gemm(a2h, a2l, m_tmp, m_tmp, q2, 1, stream, "psn_lg_gemm_kernel<dec2>");
    }
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

// ---- reverse sweep (discrete adjoint) on the same per-layer GEMM kernel --------------------------------------------------------
// Exact reverse mode of the loop above (what loss.backward() replays in the reference, neural_01_DAE_02_direct_encode.py training
// loop -> my_solvers.py:82-131).  Nothing of the forward pass is kept except the trajectory: x_sol / i_sol are the checkpoints, and
// every step's stage inputs y_e and hidden activations a1_e are recomputed from x_sol[j-1] / i_sol[j-1] by the forward's own GEMM
// launches (the cfg5 activation tape would be 150 GB per shard).  Per step, walking j = T-1 .. 1:
//     AE at j:    h_j recomputed;  dh = (A2^T gi_j) * ELU'(h_j);  gx_j += A1x^T dh;  d pre_ae[j] = dh
//     DE step j:  stages recomputed;  for e = last .. 0:  delta1_e = (W2^T dk_e) * ELU'(a1_e);  dy_e = F_x^T delta1_e -> Runge-Kutta
//                 adjoint -> dk_{e-1} / gx_{j-1};  dsum = sum_e delta1_e = d pre_de[j-1] = dG;  gi0 = F_i^T dsum -> gi_{j-1}
//                 (or, when an event re-evaluated i0 at j-1, through that evaluation into gx_{j-1} and the jump rows)
// Weight-gradient products (delta (x) activation, summed over trajectories, stages and steps) run as big-K MN-major GEMMs
// (psn_lg_wgrad_kernel) once per `ring` steps on a ring of recomputed activations / deltas; input-series, jump, all_initial
// gradients and the layer-1 constants are hoisted to the end exactly as the forward hoists the corresponding products.
bool psn_lg_bwd_supports(const psnode_problem* p, const psnode_adjoint* a) {
    if (!psn_lg_supports(p)) return false;
    if (a->d_xteach.p || a->d_iteach.p) return false;
    const bool dae = p->kind == PSNODE_DAE;
    if (!a->gx.p && !a->fuse_x.target.p) return false;
    if (dae && !a->gi.p && !a->fuse_i.target.p) return false;
    if (!view_ok(p->a0, p->a0_sb, 0)) return false;
    if (!view_ok(p->x_sol.p, p->x_sol.st, p->x_sol.sb)) return false;
    if (dae && !view_ok(p->i_sol.p, p->i_sol.st, p->i_sol.sb)) return false;
    static const bool off = std::getenv("PSNODE_LG_BWD") && std::atoi(std::getenv("PSNODE_LG_BWD")) == 0;
    return !off;
}

int64_t psn_lg_backward_workspace(const psnode_problem* p, const psnode_adjoint* a) {
    (void)a;
    return lg_layout(p, LG_BWD).total * 4;
}

int psn_lg_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    LgCtx c;
    int st = lg_setup(c, p, LG_BWD, ws, ws_bytes, stream);
    if (st != PSNODE_OK) return st;
    const LgLayout& L = c.L;
    float* w = c.w;
    const bool dae = c.dae;
    const int H = c.H, B = c.B, T = c.T, E = c.E, S = c.S, NS = c.nstages;
    const int X = p->X, Z = p->Z, V = p->V;
    const int64_t BH = c.BH, BHa = al(BH);
    const int ring = L.ring, depth = L.depth;
    const bool overlap = depth > ring;
    const int nblk = (int)((BH + 255) / 256);
    auto buf = [&](int i) { return w + L.bufs + (int64_t)i * BHa; };
    float* slabs = w + L.slabs;

    // ---- d_theta layout (psnode_b200.h): de: W1 (H x 3S), b1, W2 (H x H), b2; ae: A1 (H x (S + X + Z + V)), ab1, A2, ab2 ----
    const int lda = S + X + Z + V;
    const int64_t o_W1 = 0, o_b1 = o_W1 + (int64_t)H * 3 * S, o_W2 = o_b1 + H, o_b2 = o_W2 + (int64_t)H * H;
    const int64_t o_A1 = o_b2 + H, o_ab1 = o_A1 + (int64_t)H * lda, o_A2 = o_ab1 + H, o_ab2 = o_A2 + (int64_t)H * H;
    const int64_t ntheta = dae ? o_ab2 + H : o_A1;
    if (a->n_theta < ntheta) return PSNODE_EINVAL;
    float* th = a->d_theta;
    PSN_CUDA(cudaMemsetAsync(th, 0, (size_t)ntheta * 4, stream));
    // zero: kept dy buffers (schemes with fewer stages read them as zeros), accumulators, jump-row gradients, last d pre_de row
    for (int i : {BUF_DY3, BUF_DY2, BUF_DY1, BUF_DCDE, BUF_DCAE, BUF_ACCDK, BUF_ACCGI, BUF_GI0})
        PSN_CUDA(cudaMemsetAsync(buf(i), 0, (size_t)BH * 4, stream));
    PSN_CUDA(cudaMemsetAsync(w + L.dpj_de, 0, (size_t)(E > 0 ? E : 1) * BH * 4, stream));
    PSN_CUDA(cudaMemsetAsync(w + L.dpj_ae, 0, (size_t)(E > 0 ? E : 1) * BH * 4, stream));
    // time-chunk scratch: pre_de[k] = hoisted DE half of grid row r0 + k, pre_ae[k] = hoisted AE half of grid row r0 + k; both are
    // overwritten in place by their gradients as the sweep passes, and the chunk's share of the hoisted gradients is taken before the
    // next (earlier) chunk reuses the scratch
    float* dpre_de = w + L.pre_de;
    float* dpre_ae = w + L.pre_ae;
    const int RC = L.rc;
    int r0 = 0;                                                                // first grid row of the current chunk

    // ---- tensor maps of the reverse pass ----
    CUtensorMap m_buf, m_xsol, m_isol, m_dpde, m_dpae, m_dpjde, m_dpjae;
    bool ok = lg_make_map(&m_buf, w + L.bufs, H, B, H, L.nbufs, BHa) && lg_make_map(&m_xsol, p->x_sol.p, H, B, p->x_sol.sb, T, p->x_sol.st);
    if (dae) ok = ok && lg_make_map(&m_isol, p->i_sol.p, H, B, p->i_sol.sb, T, p->i_sol.st);
    ok = ok && lg_make_map(&m_dpde, dpre_de, H, B, H, RC, BH);
    if (dae) ok = ok && lg_make_map(&m_dpae, dpre_ae, H, B, H, RC + 1, BH);
    ok = ok && lg_make_map(&m_dpjde, w + L.dpj_de, H, B, H, E > 0 ? E : 1, BH) && lg_make_map(&m_dpjae, w + L.dpj_ae, H, B, H, E > 0 ? E : 1, BH);
    if (!ok) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer path, reverse)");

    auto gemm = [&](int which, int bidx, LgParams& q, const char* name) {      // B operand = work buffer `bidx`
        q.b_r0 = bidx;
        c.launch(which, m_buf, m_buf, q, name);
    };
    auto gated = [&](LgParams& q, int ev_j) { q.ev = p->event_idx; q.ev_j = ev_j; q.skip_unless_event = 1; };
    const float* t = p->t.p;
    const float c_last = p->method == PSNODE_RK4 ? 0.125f : 1.0f;

    // AE at grid point j: gi (ring slot s) -> dh, gx, d pre_ae[j]
    auto ae_backward = [&](int j, int s) {
        LgParams q1 = c.base();                                                // h_j = ELU(A1x x_j + pre_ae[j])
        q1.mode = LG_HIDDEN;
        q1.add1 = dpre_ae + (int64_t)(j - r0) * BH;
        q1.out = buf(L.ring_one(R_H, s));
        q1.b_r0 = j;
        c.launch(M_AE1X, m_xsol, m_xsol, q1, "psn_lg_gemm_kernel<bwd:ae1>");
        LgParams q2 = c.base();                                                // dh = (A2^T gi_j) * ELU'(h_j)
        q2.mode = LG_DELTA;
        q2.add1 = buf(L.ring_one(R_H, s));
        q2.out = buf(L.ring_one(R_DH, s));
        q2.out2 = dpre_ae + (int64_t)(j - r0) * BH; q2.out2_ld = H;
        q2.acc = buf(BUF_DCAE);
        gemm(M_T_AE2, L.ring_one(R_GI, s), q2, "psn_lg_gemm_kernel<bwd:ae2^T>");
        LgParams q3 = c.base();                                                // gx += A1x^T dh
        q3.add1 = buf(BUF_GX);
        q3.out = buf(BUF_GX);
        gemm(M_T_AE1X, L.ring_one(R_DH, s), q3, "psn_lg_gemm_kernel<bwd:ae1x^T>");
    };
    // Side stream for the batch weight-gradient GEMMs (see lg_overlap_wgrad): capped at the SMs a layer launch leaves free, they run one
    // batch behind the sweep on the other half of the (double-depth) ring.
    static cudaStream_t side = nullptr;
    static cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    static int side_dev = -1;
    int cur_dev = 0;
    PSN_CUDA(cudaGetDevice(&cur_dev));
    if (overlap && (!side || side_dev != cur_dev)) {        // (one process per GPU: created once; re-created if the device changed)
        side_dev = cur_dev;
        int lo = 0, hi = 0;
        PSN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PSN_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, lo));
        for (int i = 0; i < 2; i++) {
            PSN_CUDA(cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming));
            PSN_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
        }
    }
    bool done_pending[2] = {false, false};
    if (overlap) {                                           // the side stream starts behind everything already queued on `stream`
        PSN_CUDA(cudaEventRecord(ev_ready[0], stream));
        PSN_CUDA(cudaStreamWaitEvent(side, ev_ready[0], 0));
    }
    // weight-gradient products of the steps / grid points of one ring batch: grid points [j_lo, j_hi)
    auto flush = [&](int j_lo, int j_hi) -> int {
        const int de_lo = j_lo < 1 ? 1 : j_lo, n_de = j_hi - de_lo;            // DE steps in the batch
        const int half = (j_lo / ring) & 1;
        cudaStream_t ws_stream = overlap ? side : stream;
        float* wslabs = overlap ? w + L.slabs2 : slabs;
        const int cap = overlap ? lg_idle_sms(p) : 148;
        if (overlap) {
            PSN_CUDA(cudaEventRecord(ev_ready[half], stream));
            PSN_CUDA(cudaStreamWaitEvent(side, ev_ready[half], 0));
        }
        int rc = PSNODE_OK;
        if (n_de > 0) {
            const int s0 = de_lo % depth;
            rc = lg_wgrad(buf(L.ring_dk(s0, 0)), H, BHa, H, buf(L.ring_a1(s0, 0)), H, BHa, H, n_de * NS, B, th + o_W2, H, 1, wslabs, c.err, ws_stream, nullptr, 0,
                          cap);
            if (rc != PSNODE_OK) return rc;
            rc = lg_wgrad(buf(L.ring_d1(s0, 0)), H, BHa, H, buf(L.ring_y(s0, 0)), H, BHa, H, n_de * NS, B, th + o_W1 + 2 * S, 3 * S, 1, wslabs, c.err, ws_stream,
                          nullptr, 0, cap);
            if (rc != PSNODE_OK) return rc;
            if (dae) {
                rc = lg_wgrad(buf(L.ring_one(R_DSUM, s0)), H, BHa, H, buf(L.ring_one(R_I0, s0)), H, BHa, H, n_de, B, th + o_W1 + 2 * S + X + Z + V, 3 * S, 1,
                              wslabs, c.err, ws_stream, nullptr, 0, cap);
                if (rc != PSNODE_OK) return rc;
            }
        }
        if (dae && j_hi > j_lo) {
            const int s0 = j_lo % depth, n = j_hi - j_lo;
            rc = lg_wgrad(buf(L.ring_one(R_GI, s0)), H, BHa, H, buf(L.ring_one(R_H, s0)), H, BHa, H, n, B, th + o_A2, H, 1, wslabs, c.err, ws_stream, nullptr, 0,
                          cap);
            if (rc != PSNODE_OK) return rc;
            rc = lg_wgrad(buf(L.ring_one(R_DH, s0)), H, BHa, H, p->x_sol.p + (int64_t)j_lo * p->x_sol.st, p->x_sol.sb, p->x_sol.st, H, n, B,
                          th + o_A1 + S, lda, 1, wslabs, c.err, ws_stream, nullptr, 0, cap);
            if (rc != PSNODE_OK) return rc;
        }
        if (overlap) {
            PSN_CUDA(cudaEventRecord(ev_done[half], side));
            done_pending[half] = true;
            // the sweep now moves on to the batch below, which lives in the OTHER half of the ring: that half must have been consumed
            if (done_pending[half ^ 1]) { PSN_CUDA(cudaStreamWaitEvent(stream, ev_done[half ^ 1], 0)); done_pending[half ^ 1] = false; }
        }
        return PSNODE_OK;
    };

    // upstream gradient rows dL/dx_sol[j], dL/di_sol[j]: the caller's tensors, or formed on the fly from the trajectory (fused masked MSE)
    const bool fuse_x = a->fuse_x.target.p != nullptr, fuse_i = dae && a->fuse_i.target.p != nullptr;
    const PsnFuse fx = psn_make_fuse(a->fuse_x, p->x_sol), fi = psn_make_fuse(a->fuse_i, p->i_sol);
    const float* upx = nullptr; const float* upi = nullptr;
    int64_t upx_sb = H, upi_sb = H;
    auto upstream_row = [&](int j) {
        if (fuse_x || fuse_i) {
            psn_lg_fuse_row_kernel<<<nblk, 256, 0, stream>>>(B, H, j, fx, fuse_x ? buf(BUF_UPX) : nullptr, fi, fuse_i ? buf(BUF_UPI) : nullptr);
            psn_count_launch("psn_lg_fuse_row_kernel");
        }
        if (fuse_x) { upx = buf(BUF_UPX); upx_sb = H; } else { upx = a->gx.p + (int64_t)j * a->gx.st; upx_sb = a->gx.sb; }
        if (fuse_i) { upi = buf(BUF_UPI); upi_sb = H; } else if (dae) { upi = a->gi.p + (int64_t)j * a->gi.st; upi_sb = a->gi.sb; }
    };
    // ---- start: adjoints of the last grid point ----
    {
        const int s = (T - 1) % depth;
        upstream_row(T - 1);
        psn_lg_bwd_init_kernel<<<nblk, 256, 0, stream>>>(B, H, upx, upx_sb, dae ? upi : nullptr, upi_sb, buf(BUF_GX), buf(L.ring_one(R_GI, s)),
                                                         buf(BUF_ACCGI));
        psn_count_launch("psn_lg_bwd_init_kernel");
    }
    auto sweep_point = [&](int j) -> int {            // AE at grid point j, then (j >= 1) the reverse of step j
        const int s = j % depth;
        if (dae) ae_backward(j, s);
        if (j >= 1) {
            upstream_row(j - 1);
            const float* tc = t + (int64_t)j * p->t.st;
            const float* tp = t + (int64_t)(j - 1) * p->t.st;
            psn_lg_bwd_head_kernel<<<nblk, 256, 0, stream>>>(B, H, p->x_sol.p + (int64_t)(j - 1) * p->x_sol.st, p->x_sol.sb,
                                                             dae ? p->i_sol.p + (int64_t)(j - 1) * p->i_sol.st : nullptr, p->i_sol.sb,
                                                             buf(L.ring_y(s, 0)), buf(L.ring_one(R_I0, s)), buf(BUF_GX), tc, tp, p->t.sb, c_last,
                                                             buf(L.ring_dk(s, NS - 1)), buf(BUF_ACCDK));
            psn_count_launch("psn_lg_bwd_head_kernel");
            // ---- recompute the stages of step j ----
            if (dae) {
                if (E > 0) {                                                   // event at j-1: i0 = ae(x_{j-1}, jumped inputs)  (my_solvers.py:108-110)
                    LgParams q = c.base();
                    q.mode = LG_HIDDEN;
                    gated(q, j - 1);
                    q.add1_jump = w + L.pj_ae; q.add1_jump_sr = BH;
                    q.out = buf(BUF_HEV);
                    gemm(M_AE1X, L.ring_y(s, 0), q, "psn_lg_gemm_kernel<bwd:ae1,event>");
                    LgParams q2 = c.base();
                    q2.bias = p->ae.b[1];
                    gated(q2, j - 1);
                    q2.out = buf(L.ring_one(R_I0, s));
                    gemm(M_AE2, BUF_HEV, q2, "psn_lg_gemm_kernel<bwd:ae2,event>");
                }
                LgParams qg = c.base();                                        // G = F_i i0
                qg.out = buf(BUF_G);
                gemm(M_DE1I, L.ring_one(R_I0, s), qg, "psn_lg_gemm_kernel<bwd:de1_i>");
            }
            for (int e = 0; e < NS; e++) {
                LgParams q1 = c.base();
                q1.mode = LG_HIDDEN;
                q1.add1 = dpre_de + (int64_t)(j - 1 - r0) * BH;
                if (E > 0) { q1.ev = p->event_idx; q1.ev_j = j - 1; q1.add1_jump = w + L.pj_de; q1.add1_jump_sr = BH; }
                if (dae) q1.add2 = buf(BUF_G);
                q1.out = buf(L.ring_a1(s, e));
                gemm(M_DE1X, L.ring_y(s, e), q1, "psn_lg_gemm_kernel<bwd:de1>");
                if (e + 1 < NS) {                                              // next stage input (the last stage's slope is not needed)
                    LgParams q2 = c.base();
                    q2.mode = LG_RK;
                    q2.bias = p->de.b[1];
                    q2.method = p->method; q2.stage = e;
                    q2.x0 = buf(L.ring_y(s, 0)); q2.k1 = buf(BUF_K1); q2.k2 = buf(BUF_K2); q2.k3 = buf(BUF_K3);
                    q2.t_cur = tc; q2.t_prev = tp; q2.t_sb = p->t.sb;
                    q2.out = buf(L.ring_y(s, e + 1));
                    gemm(M_DE2, L.ring_a1(s, e), q2, "psn_lg_gemm_kernel<bwd:de2,rk>");
                }
            }
            // ---- reverse through the stages ----
            for (int e = NS - 1; e >= 0; e--) {
                LgParams q1 = c.base();                                        // delta1_e = (W2^T dk_e) * ELU'(a1_e); dsum (+)= delta1_e
                q1.mode = LG_DELTA;
                q1.add1 = buf(L.ring_a1(s, e));
                q1.out = buf(L.ring_d1(s, e));
                q1.out2 = buf(L.ring_one(R_DSUM, s)); q1.out2_ld = H; q1.out2_acc = e == NS - 1 ? 0 : 1;
                q1.acc = buf(BUF_DCDE);
                gemm(M_T_DE2, L.ring_dk(s, e), q1, "psn_lg_gemm_kernel<bwd:de2^T>");
                LgParams q2 = c.base();                                        // dy_e = F_x^T delta1_e -> Runge-Kutta adjoint
                q2.mode = LG_BRK;
                q2.x0 = buf(BUF_GX); q2.k1 = buf(BUF_DY3); q2.k2 = buf(BUF_DY2); q2.k3 = buf(BUF_DY1);
                q2.t_cur = tc; q2.t_prev = tp; q2.t_sb = p->t.sb;
                if (e == 0) {
                    q2.stage = EPI_BSUM;
                    q2.add1 = upx; q2.add1_ld = upx_sb;
                    q2.out = buf(BUF_GX);
                } else {
                    q2.stage = p->method == PSNODE_MIDPOINT ? (int)EPI_BMID1 : (e == 3 ? (int)EPI_BRK3 : (e == 2 ? (int)EPI_BRK2 : (int)EPI_BRK1));
                    q2.out = buf(L.ring_dk(s, e - 1));
                    q2.acc = buf(BUF_ACCDK);
                }
                gemm(M_T_DE1X, L.ring_d1(s, e), q2, "psn_lg_gemm_kernel<bwd:de1x^T>");
            }
            if (dae) {
                LgParams qi = c.base();                                        // gi0 = F_i^T dsum
                qi.out = buf(BUF_GI0);
                gemm(M_T_DE1I, L.ring_one(R_DSUM, s), qi, "psn_lg_gemm_kernel<bwd:de1i^T>");
                if (E > 0) {                                                   // the event's own evaluation of i0: back into x_{j-1} and the jump rows
                    LgParams q1 = c.base();
                    q1.mode = LG_DELTA;
                    gated(q1, j - 1);
                    q1.add1 = buf(BUF_HEV);
                    q1.out = buf(BUF_DHEV);
                    q1.out2_jump = w + L.dpj_ae; q1.out2_jump_sr = BH; q1.out2_ld = H;
                    q1.acc = buf(BUF_DCAE);
                    gemm(M_T_AE2, BUF_GI0, q1, "psn_lg_gemm_kernel<bwd:ae2^T,event>");
                    LgParams q2 = c.base();
                    gated(q2, j - 1);
                    q2.add1 = buf(BUF_GX);
                    q2.out = buf(BUF_GX);
                    gemm(M_T_AE1X, BUF_DHEV, q2, "psn_lg_gemm_kernel<bwd:ae1x^T,event>");
                    st = lg_wgrad(buf(BUF_GI0), H, BHa, H, buf(BUF_HEV), H, BHa, H, 1, B, th + o_A2, H, 1, slabs, c.err, stream, p->event_idx, j - 1);
                    if (st != PSNODE_OK) return st;
                    st = lg_wgrad(buf(BUF_DHEV), H, BHa, H, buf(L.ring_y(s, 0)), H, BHa, H, 1, B, th + o_A1 + S, lda, 1, slabs, c.err, stream, p->event_idx,
                                  j - 1);
                    if (st != PSNODE_OK) return st;
                }
            }
        }
        if (j % ring == 0) {
            st = flush(j, j + ring < T ? j + ring : T);
            if (st != PSNODE_OK) return st;
        }
        if (j >= 1) {
            const int sn = (j - 1) % depth;
            psn_lg_bwd_tail_kernel<<<nblk, 256, 0, stream>>>(B, H, E > 0 ? p->event_idx : nullptr, j - 1, buf(L.ring_one(R_DSUM, s)),
                                                             dpre_de + (int64_t)(j - 1 - r0) * BH, w + L.dpj_de, BH, dae ? buf(BUF_GI0) : nullptr,
                                                             dae ? upi : nullptr, upi_sb, dae ? buf(L.ring_one(R_GI, sn)) : nullptr, buf(BUF_ACCGI));
            psn_count_launch("psn_lg_bwd_tail_kernel");
        }
        return PSNODE_OK;
    };
    // the chunk's share of the hoisted gradients: input-series rows (d_z[r] += F_z^T d pre_de[r] and += A1z^T d pre_ae[r], the two halves of
    // the [F_z^T | A1z^T] plane on their own because a row's two parts can belong to different chunks) and the weight gradients
    // dF_z / dF_v += sum_r d pre_de[r] (x) z[r] / v[r],  dA1z / dA1v likewise
    auto chunk_grads = [&](int de_rows, int ae_first, int ae_rows) -> int {    // DE rows r0 .. r0+de_rows-1; AE rows ae_first .. ae_first+ae_rows-1
        // Order of the two parts of a row r: its DE part is always written first (same chunk, or -- for the chunk's top row -- by the
        // later chunk that was swept earlier), so the DE part is a plain store and the AE part accumulates; only row T-1 has no DE part
        auto series_part = [&](int which, const CUtensorMap& m_src, int src_row0, int rows, int grid_row0, int a_c0, bool accumulate, float* out,
                               int64_t out_sr, int64_t out_sb, const char* name) {
            if (!out || rows <= 0) return;
            LgParams q = c.base();
            q.R = rows; q.b_r0 = src_row0; q.a_c0 = a_c0;
            q.out = out + (int64_t)grid_row0 * out_sr; q.out_sr = out_sr; q.out_ld = out_sb;
            if (accumulate) { q.add1 = q.out; q.add1_sr = out_sr; q.add1_ld = out_sb; }
            c.launch(which, m_src, m_src, q, name);
        };
        if (Z > 0) series_part(M_T_DZ, m_dpde, 0, de_rows, r0, 0, false, a->d_z.p, a->d_z.st, a->d_z.sb, "psn_lg_gemm_kernel<bwd:d_z,de>");
        if (dae) {
            series_part(M_T_DV, m_dpde, 0, de_rows, r0, 0, false, a->d_v.p, a->d_v.st, a->d_v.sb, "psn_lg_gemm_kernel<bwd:d_v,de>");
            if (Z > 0) series_part(M_T_DZ, m_dpae, ae_first - r0, ae_rows, ae_first, H, true, a->d_z.p, a->d_z.st, a->d_z.sb, "psn_lg_gemm_kernel<bwd:d_z,ae>");
            series_part(M_T_DV, m_dpae, ae_first - r0, ae_rows, ae_first, H, true, a->d_v.p, a->d_v.st, a->d_v.sb, "psn_lg_gemm_kernel<bwd:d_v,ae>");
        }
        int rc = PSNODE_OK;
        if (de_rows > 0) {
            if (Z > 0) {
                rc = lg_wgrad(dpre_de, H, BH, H, p->z.p + (int64_t)r0 * p->z.st, p->z.sb, p->z.st, H, de_rows, B, th + o_W1 + 2 * S + X, 3 * S, 1, slabs,
                              c.err, stream);
                if (rc != PSNODE_OK) return rc;
            }
            if (dae) {
                rc = lg_wgrad(dpre_de, H, BH, H, p->v.p + (int64_t)r0 * p->v.st, p->v.sb, p->v.st, H, de_rows, B, th + o_W1 + 2 * S + X + Z, 3 * S, 1, slabs,
                              c.err, stream);
                if (rc != PSNODE_OK) return rc;
            }
        }
        if (dae && ae_rows > 0) {
            const float* pp = dpre_ae + (int64_t)(ae_first - r0) * BH;
            if (Z > 0) {
                rc = lg_wgrad(pp, H, BH, H, p->z.p + (int64_t)ae_first * p->z.st, p->z.sb, p->z.st, H, ae_rows, B, th + o_A1 + S + X, lda, 1, slabs, c.err,
                              stream);
                if (rc != PSNODE_OK) return rc;
            }
            rc = lg_wgrad(pp, H, BH, H, p->v.p + (int64_t)ae_first * p->v.st, p->v.sb, p->v.st, H, ae_rows, B, th + o_A1 + S + X + Z, lda, 1, slabs, c.err,
                          stream);
            if (rc != PSNODE_OK) return rc;
        }
        return PSNODE_OK;
    };
    // the last grid row has no DE part (z[T-1] / v[T-1] only enter the AE): it starts at zero
    if (a->d_z.p) {
        psn_lg_zero_rows_kernel<<<nblk, 256, 0, stream>>>(1, B, H, a->d_z.p + (int64_t)(T - 1) * a->d_z.st, 0, a->d_z.sb);
        psn_count_launch("psn_lg_zero_rows_kernel");
    }
    if (dae && a->d_v.p) {
        psn_lg_zero_rows_kernel<<<nblk, 256, 0, stream>>>(1, B, H, a->d_v.p + (int64_t)(T - 1) * a->d_v.st, 0, a->d_v.sb);
        psn_count_launch("psn_lg_zero_rows_kernel");
    }
    if (T == 1) {
        if (dae) c.project(M_AE1ZV, c.m_z, c.m_v, 0, 1, w + L.c_ae, dpre_ae, "psn_lg_gemm_kernel<bwd:pre_ae>");
        st = sweep_point(0);
        if (st != PSNODE_OK) return st;
        st = chunk_grads(0, 0, dae ? 1 : 0);
        if (st != PSNODE_OK) return st;
    }
    for (int ck = T > 1 ? (T - 2) / RC : -1; ck >= 0; ck--) {
        r0 = ck * RC;
        const int r1 = r0 + RC < T - 1 ? r0 + RC : T - 1, n = r1 - r0;
        c.project(M_DE1ZV, c.m_z, c.m_v, r0, n, w + L.c_de, dpre_de, "psn_lg_gemm_kernel<bwd:pre_de>");
        if (dae) {
            const int first = r0 == 0 ? 0 : r0 + 1;
            c.project(M_AE1ZV, c.m_z, c.m_v, first, r1 - first + 1, w + L.c_ae, dpre_ae + (int64_t)(first - r0) * BH, "psn_lg_gemm_kernel<bwd:pre_ae>");
        }
        for (int j = r1; j > r0; j--) {
            st = sweep_point(j);
            if (st != PSNODE_OK) return st;
        }
        if (r0 == 0) {
            st = sweep_point(0);
            if (st != PSNODE_OK) return st;
        }
        const int ae_first = r0 == 0 ? 0 : r0 + 1;
        st = chunk_grads(n, ae_first, dae ? r1 - ae_first + 1 : 0);
        if (st != PSNODE_OK) return st;
    }
    PSN_CUDA(cudaGetLastError());
    for (int i = 0; i < 2; i++)
        if (done_pending[i]) { PSN_CUDA(cudaStreamWaitEvent(stream, ev_done[i], 0)); done_pending[i] = false; }

    // ---- gradient of the initial state ----
    if (a->d_x0) {
        psn_lg_copy_rows_kernel<<<nblk, 256, 0, stream>>>(B, H, buf(BUF_GX), a->d_x0, a->d_x0_sb);
        psn_count_launch("psn_lg_copy_rows_kernel");
    }
    // ---- jump-row gradients: d_zjump[k] = F_z^T d pre_de_jump[k] + A1z^T d pre_ae_jump[k], and their share of dF_z / dF_v / dA1z / dA1v ----
    auto jump_grad = [&](int which, float* out, int64_t out_se, int64_t out_sb, const char* name) {
        if (!out || E <= 0) return;
        LgParams q = c.base();
        q.R = E; q.nsrc = dae ? 2 : 1;
        q.out = out; q.out_sr = out_se; q.out_ld = out_sb;
        c.launch(which, m_dpjde, dae ? m_dpjae : m_dpjde, q, name);
    };
    if (Z > 0) jump_grad(M_T_DZ, a->d_zjump, a->d_zj_se, a->d_zj_sb, "psn_lg_gemm_kernel<bwd:d_zjump>");
    if (dae) jump_grad(M_T_DV, a->d_vjump, a->d_vj_se, a->d_vj_sb, "psn_lg_gemm_kernel<bwd:d_vjump>");
    if (E > 0) {
        if (Z > 0) {
            st = lg_wgrad(w + L.dpj_de, H, BH, H, p->z_jump, p->zj_sb, p->zj_se, H, E, B, th + o_W1 + 2 * S + X, 3 * S, 1, slabs, c.err, stream);
            if (st != PSNODE_OK) return st;
        }
        if (dae) {
            st = lg_wgrad(w + L.dpj_de, H, BH, H, p->v_jump, p->vj_sb, p->vj_se, H, E, B, th + o_W1 + 2 * S + X + Z, 3 * S, 1, slabs, c.err, stream);
            if (st != PSNODE_OK) return st;
            if (Z > 0) {
                st = lg_wgrad(w + L.dpj_ae, H, BH, H, p->z_jump, p->zj_sb, p->zj_se, H, E, B, th + o_A1 + S + X, lda, 1, slabs, c.err, stream);
                if (st != PSNODE_OK) return st;
            }
            st = lg_wgrad(w + L.dpj_ae, H, BH, H, p->v_jump, p->vj_sb, p->vj_se, H, E, B, th + o_A1 + S + X + Z, lda, 1, slabs, c.err, stream);
            if (st != PSNODE_OK) return st;
        }
    }
    // ---- per-trajectory layer-1 constants: c_de = (W_a - W_b) a0 + b1, c_ae = A1a a0 + ab1 ----
    st = lg_wgrad(buf(BUF_DCDE), H, BH, H, p->a0, p->a0_sb, 0, S, 1, B, th + o_W1, 3 * S, 1, slabs, c.err, stream);      // dW_a = dc (x) a0
    if (st != PSNODE_OK) return st;
    if (dae) {
        st = lg_wgrad(buf(BUF_DCAE), H, BH, H, p->a0, p->a0_sb, 0, S, 1, B, th + o_A1, lda, 1, slabs, c.err, stream);
        if (st != PSNODE_OK) return st;
    }
    psn_lg_unfold_w1_kernel<<<(H * S + 255) / 256, 256, 0, stream>>>(th + o_W1, H, S);
    psn_count_launch("psn_lg_unfold_w1_kernel");
    psn_lg_colsum_kernel<<<H / 32, 1024, 0, stream>>>(buf(BUF_DCDE), B, H, th + o_b1);
    psn_lg_colsum_kernel<<<H / 32, 1024, 0, stream>>>(buf(BUF_ACCDK), B, H, th + o_b2);
    psn_count_launch("psn_lg_colsum_kernel"); psn_count_launch("psn_lg_colsum_kernel");
    if (dae) {
        psn_lg_colsum_kernel<<<H / 32, 1024, 0, stream>>>(buf(BUF_DCAE), B, H, th + o_ab1);
        psn_lg_colsum_kernel<<<H / 32, 1024, 0, stream>>>(buf(BUF_ACCGI), B, H, th + o_ab2);
        psn_count_launch("psn_lg_colsum_kernel"); psn_count_launch("psn_lg_colsum_kernel");
    }
    if (a->d_a0) {                                                             // d_a0 = (W_a - W_b)^T dc_de + A1a^T dc_ae
        LgParams q = c.base();
        q.nsrc = dae ? 2 : 1;
        q.out = a->d_a0; q.out_ld = a->d_a0_sb;
        LgParams qq = q;
        qq.b_r0 = 0;
        CUtensorMap m_dcde, m_dcae;
        if (!lg_make_map(&m_dcde, buf(BUF_DCDE), H, B, H, 1, 0) || !lg_make_map(&m_dcae, buf(BUF_DCAE), H, B, H, 1, 0))
            return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer path, d_a0)");
        c.launch(M_T_DA0, m_dcde, m_dcae, qq, "psn_lg_gemm_kernel<bwd:d_a0>");
    }
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

// Test hook (not part of the public header): the MN-major weight-gradient GEMM alone.  out[M x K] = sum_{slot, n} P[slot][n][:]^T Q[slot][n][:].
extern "C" int psnode_debug_lg_wgrad(const float* P, int64_t p_sn, int64_t p_ss, int M, const float* Q, int64_t q_sn, int64_t q_ss, int K, int nslots,
                                     int N, float* out, void* ws, int64_t ws_bytes, void* stream) {
    if ((M != 128 && M != 256 && M != 512) || (K % 128) != 0 || K < 128) return PSNODE_EINVAL;
    if (ws_bytes < 256 + lg_wgrad_slab_floats(M, K) * 4) return PSNODE_EWORKSPACE;
    int* err = static_cast<int*>(ws);
    float* slabs = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + 256);
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, static_cast<cudaStream_t>(stream)));
    return lg_wgrad(P, p_sn, p_ss, M, Q, q_sn, q_ss, K, nslots, N, out, K, 0, slabs, err, static_cast<cudaStream_t>(stream));
}
extern "C" int64_t psnode_debug_lg_wgrad_workspace(int M, int K) { return 256 + lg_wgrad_slab_floats(M, K) * 4; }
