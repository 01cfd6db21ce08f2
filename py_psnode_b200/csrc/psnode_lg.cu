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

constexpr int TM = 128, TN = 128;
constexpr int NST = 3;
constexpr int SLAB = TN * 128;          // 16 KB: 128 rows x 32 fp32
constexpr int LG_THREADS = 320;         // 8 split / epilogue warps + TMA producer warp + MMA issuer warp
constexpr int NPART = 2;                // K-partials (truncating fp32 accumulate: short chains)

enum { LG_PLAIN = 0, LG_HIDDEN = 1, LG_RK = 2, LG_DELTA = 3 };
// epilogue kinds (template parameter of the GEMM kernel)
enum { EPI_PLAIN = 0, EPI_HIDDEN, EPI_EULER, EPI_MID0, EPI_MID1, EPI_RK0, EPI_RK1, EPI_RK2, EPI_RK3, EPI_DELTA };
// EPI_DELTA (reverse pass): out = D * ELU'(act) with act = the recorded post-ELU activation (add1 operand); out2 (+)= out

struct __align__(1024) LgSmem {
    unsigned char a_hi[NST][SLAB], a_lo[NST][SLAB], b_hi[NST][SLAB], b_lo[NST][SLAB];
    uint64_t full[NST], split[NST], done[NST];
    uint32_t tmem_base;
};

__device__ long long g_lg_dbg[64];        // clock64 stamps of one CTA (PSNODE_LG_DBG=<cta index + 1>): phase breakdown on the device

struct LgParams {
    int dbg, rotate;
    int N, R, nbt;                      // trajectories per row, rows (1 for a layer launch), n-tiles per row
    int nsrc, kchunks;                  // B sources (K segments) and 32-wide chunks per source
    int mode;
    // event handling (neural_base.py:52-65, 180-196): ev[ev_j] = index of the event that fires when leaving grid point ev_j, or -1
    const int32_t* ev; int ev_j; int skip_unless_event; int skip_if_event;
    int b_r0;                           // added to the row coordinate of the B operand (slot of a ring of activation buffers)
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
    const float* t_cur; const float* t_prev; int64_t t_sb;   // t[j], t[j-1] rows (element n at n * t_sb)
    int* err;
};

template <int EPI>
__global__ void __launch_bounds__(LG_THREADS, 1) psn_lg_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                                                                   const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1,
                                                                   const __grid_constant__ LgParams q) {
    // programmatic dependent launch: everything above the first read of the previous launch's output may overlap its tail
    asm volatile("griddepcontrol.wait;" ::: "memory");
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
                mbar_expect_tx(&sm.full[s], 3 * SLAB);
                tma_load_3d(sm.a_hi[s], &map_a_hi, ck * 32, mblk * TM, 0, &sm.full[s]);
                tma_load_3d(sm.a_lo[s], &map_a_lo, ck * 32, mblk * TM, 0, &sm.full[s]);
                tma_load_3d(sm.b_hi[s], src == 0 ? &map_b0 : &map_b1, kc * 32, b0, r + q.b_r0, &sm.full[s]);
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
#pragma unroll
            for (int e = 0; e < SLAB / 16 / 256; e++) {
                const int idx = tid + e * 256;
                float4 lo;
                const float4 hi = split4_hi(h4[idx], lo);
                h4[idx] = hi;
                l4[idx] = lo;
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
    if (!mbar_wait(&sm.done[(nchunk - 1) % NST], (uint32_t)(((nchunk - 1) / NST) & 1))) { atomicExch(q.err, 13); __trap(); }
    tc_fence_after();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // the next layer's grid may be set up while this epilogue runs
    stamp();

    // ---- epilogue: lane = output feature, columns = trajectories ---------------------------------------------------------
    // Operands of the fused epilogue (hoisted layer-1 half, G, or the Runge-Kutta state) are fetched 16 columns ahead of the
    // arithmetic, so their L2 latency is paid once per tile.  The epilogue kind is a template parameter and the batch loop is
    // rolled: the first version (run-time mode switch inside fully unrolled loops) was 16 000 instructions of straight-line
    // code executed once per CTA and spent 38 000 of its 54 000 cycles waiting for instruction fetches (`no_inst` stalls).
    {
        constexpr bool rk = EPI >= EPI_EULER && EPI <= EPI_RK3;
        const int m = mblk * TM + 32 * wq + lane;
        const float* add1 = q.add1;
        int64_t add1_row = (int64_t)r * q.add1_sr;
        if (evk >= 0 && q.add1_jump) { add1 = q.add1_jump; add1_row = (int64_t)evk * q.add1_jump_sr; }
        const float bias = q.bias ? __ldg(q.bias + m) : 0.0f;
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
        float* px0 = rk ? q.x0 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk1 = rk ? q.k1 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk2 = rk ? q.k2 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        float* pk3 = rk ? q.k3 + (int64_t)ncol0 * q.st_ld + m : nullptr;
        const float* ptc = rk ? q.t_cur + (int64_t)ncol0 * q.t_sb : nullptr;
        const float* ptp = rk ? q.t_prev + (int64_t)ncol0 * q.t_sb : nullptr;
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
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int c = 16 * bt + i;
                if (!full_tile && c >= nlive) continue;
                float v = (t0[i] + t1[i]) + bias;
                if constexpr (EPI == EPI_DELTA) {
                    v = (t0[i] + t1[i]) * psn_elu_grad_from_out(e.p0[i]);
                    pout[c * old] = v;
                    if (pout2) pout2[c * o2ld] = q.out2_acc ? pout2[c * o2ld] + v : v;
                } else if constexpr (!rk) {
                    v = (v + e.p0[i]) + e.p1[i];
                    if constexpr (EPI == EPI_HIDDEN) v = psn_elu(v);
                    pout[c * old] = v;
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
        Buf cur, nxt;
        fetch(0, cur);
#pragma unroll 1
        for (int bt = 0; bt < 4; bt++) {
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
}

// ---- one-time preparation per call ------------------------------------------------------------------------------------------
// dst_hi / dst_lo [m][k] (k < K) = split_tf32(W[m * ldw + col0 + k] (+ W[m * ldw + col1 + k] if col1 >= 0))
__global__ void psn_lg_prep_kernel(const float* __restrict__ W, int ldw, int col0, int col1, int M, int K, float* __restrict__ hi, float* __restrict__ lo) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * K) return;
    const int m = idx / K, k = idx - m * K;
    float w = __ldg(W + (int64_t)m * ldw + col0 + k);
    if (col1 >= 0) w += __ldg(W + (int64_t)m * ldw + col1 + k);
    float h, l;
    split_tf32(w, h, l);
    hi[idx] = h;
    lo[idx] = l;
}
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
    int dbg, lbo, sbo, kadv;        // debugging / probing of the MN-major descriptor fields (bytes)
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
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)(4096 >> 4) << 16;          // LBO: next block of 32 features (one TMA box of 32 rows x 128 B)
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO: next group of 8 reduction rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
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
        const uint32_t idesc = q.dbg == 4 ? make_idesc_tf32(TM, TN) : (q.dbg == 5 ? (make_idesc_tf32(TM, TN) | (1u << 15)) : (q.dbg == 6 ? (make_idesc_tf32(TM, TN) | (1u << 16)) : make_idesc_tf32_mn(TM, TN)));
        for (int c = 0; c < n; c++) {
            const int s = c % NST;
            const int pair = c >> 1;
            if ((c & 1) == 0 && pair >= 2 && !mbar_wait(&sm.drained[pair & 1], (uint32_t)(((pair - 2) >> 1) & 1))) { atomicExch(q.err, 22); __trap(); }
            if (!mbar_wait(&sm.split[s], (uint32_t)((c / NST) & 1))) { atomicExch(q.err, 23); __trap(); }
            if (elect_one()) {
                tc_fence_after();
                const uint64_t dp_hi = make_desc_mn(smem_u32(sm.p_hi[s]), q.lbo, q.sbo), dp_lo = make_desc_mn(smem_u32(sm.p_lo[s]), q.lbo, q.sbo);
                const uint64_t dq_hi = make_desc_mn(smem_u32(sm.q_hi[s]), q.lbo, q.sbo), dq_lo = make_desc_mn(smem_u32(sm.q_lo[s]), q.lbo, q.sbo);
                const uint64_t kadv = (uint64_t)(q.kadv >> 4);
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
            if (q.dbg == 3 && c == 0 && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {      // debugging aid
                float* f = reinterpret_cast<float*>(q.err);
                q.err[1] = n; q.err[2] = cps; q.err[3] = (int)total;
                f[4] = reinterpret_cast<const float*>(sm.p_hi[s])[0]; f[5] = reinterpret_cast<const float*>(sm.p_hi[s])[33];
                f[6] = reinterpret_cast<const float*>(sm.q_hi[s])[0]; f[7] = reinterpret_cast<const float*>(sm.p_lo[s])[0];
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
        if (q.dbg == 3 && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
            float* f = reinterpret_cast<float*>(q.err);
            f[8] = acc[0]; f[9] = acc[1]; f[10] = acc[63];
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
             int64_t out_ld, int accumulate, float* slabs, int* err, cudaStream_t stream, const int32_t* ev = nullptr, int ev_j = 0) {
    CUtensorMap mp, mq;
    if (!lg_make_map32(&mp, P, M, N, p_sn, nslots, p_ss) || !lg_make_map32(&mq, Q, K, N, q_sn, nslots, q_ss))
        return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer weight gradients)");
    WgParams q;
    q.nslots = nslots; q.N = N;
    q.mblks = M / TM; q.kblks = K / TN;
    const int ntiles = q.mblks * q.kblks;
    const int64_t total = (int64_t)nslots * ((N + 31) / 32);
    int nsplit = 148 / ntiles;
    if (nsplit > WG_NSPLIT_MAX) nsplit = WG_NSPLIT_MAX;
    if (nsplit > total) nsplit = (int)total;
    if (nsplit < 1) nsplit = 1;
    q.nsplit = nsplit;
    q.slabs = slabs; q.err = err;
    q.ev = ev; q.ev_j = ev_j;
    q.dbg = std::getenv("PSNODE_WG_DBG") ? std::atoi(std::getenv("PSNODE_WG_DBG")) : 0;
    q.lbo = std::getenv("PSNODE_WG_LBO") ? std::atoi(std::getenv("PSNODE_WG_LBO")) : 4096;
    q.sbo = std::getenv("PSNODE_WG_SBO") ? std::atoi(std::getenv("PSNODE_WG_SBO")) : 512;
    q.kadv = std::getenv("PSNODE_WG_KADV") ? std::atoi(std::getenv("PSNODE_WG_KADV")) : 1024;
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
    static const LgKernFn kerns[10] = {psn_lg_gemm_kernel<EPI_PLAIN>, psn_lg_gemm_kernel<EPI_HIDDEN>, psn_lg_gemm_kernel<EPI_EULER>,
                                       psn_lg_gemm_kernel<EPI_MID0>, psn_lg_gemm_kernel<EPI_MID1>, psn_lg_gemm_kernel<EPI_RK0>,
                                       psn_lg_gemm_kernel<EPI_RK1>, psn_lg_gemm_kernel<EPI_RK2>, psn_lg_gemm_kernel<EPI_RK3>,
                                       psn_lg_gemm_kernel<EPI_DELTA>};
    return kerns;
}
int lg_gemm_smem() { return (int)sizeof(LgSmem) + 1024; }
int lg_prepare_kernels() {
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < 10; i++) PSN_CUDA(cudaFuncSetAttribute(lg_kernels()[i], cudaFuncAttributeMaxDynamicSharedMemorySize, lg_gemm_smem()));
        attr_set = true;
    }
    return PSNODE_OK;
}
int lg_epi_of(const LgParams& q) {
    if (q.mode == LG_PLAIN) return (int)EPI_PLAIN;
    if (q.mode == LG_HIDDEN) return (int)EPI_HIDDEN;
    if (q.mode == LG_DELTA) return (int)EPI_DELTA;
    if (q.method == PSNODE_EULER) return (int)EPI_EULER;
    if (q.method == PSNODE_MIDPOINT) return q.stage == 0 ? (int)EPI_MID0 : (int)EPI_MID1;
    return (int)EPI_RK0 + q.stage;
}
bool lg_use_pdl() {
    static const bool v = std::getenv("PSNODE_LG_PDL") ? std::atoi(std::getenv("PSNODE_LG_PDL")) != 0 : true;
    return v;
}
// one GEMM launch: A planes (hi, lo maps), one or two B sources
int lg_launch_gemm(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b0, const CUtensorMap& b1, const LgParams& q, int mblks,
                   cudaStream_t stream, const char* name) {
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

struct LgLayout {
    int H, KZV, nmat;
    int64_t err, wts_hi[7], wts_lo[7], c_de, c_ae, pre_de, pre_ae, x0, k1, k2, k3, ycur, a1, hbuf, icur, G, total;
    int64_t rows_de, rows_ae;
};
enum { M_DE1X = 0, M_DE1ZV, M_DE1I, M_DE2, M_AE1X, M_AE1ZV, M_AE2 };

LgLayout lg_layout(const psnode_problem* p) {
    LgLayout L;
    const bool dae = p->kind == PSNODE_DAE;
    const int H = p->X, E = p->event_idx ? p->E : 0;
    L.H = H;
    L.KZV = p->Z + p->V;
    L.nmat = dae ? 7 : 4;
    int64_t o = 64;
    L.err = 0;
    const int kt[7] = {H, L.KZV, H, H, H, L.KZV, H};
    for (int i = 0; i < 7; i++) {
        L.wts_hi[i] = o; o += al((int64_t)H * kt[i]);
        L.wts_lo[i] = o; o += al((int64_t)H * kt[i]);
    }
    const int64_t BH = (int64_t)p->B * H;
    L.c_de = o; o += al(BH);
    L.c_ae = o; o += al(BH);
    L.rows_de = (p->T > 1 ? p->T - 1 : 0) + E;
    L.rows_ae = dae ? (int64_t)p->T + E : 0;
    L.pre_de = o; o += al(L.rows_de * BH);
    L.pre_ae = o; o += al(L.rows_ae * BH);
    L.x0 = o; o += al(BH); L.k1 = o; o += al(BH); L.k2 = o; o += al(BH); L.k3 = o; o += al(BH);
    L.ycur = o; o += al(BH); L.a1 = o; o += al(BH); L.hbuf = o; o += al(BH); L.icur = o; o += al(BH); L.G = o; o += al(BH);
    L.total = o;
    return L;
}

bool view_ok(const float* p, int64_t s0, int64_t s1) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (s0 & 3) == 0 && (s1 & 3) == 0; }

}  // namespace

bool psn_lg_supports(const psnode_problem* p) {
    if (p->teacher_x || p->teacher_i) return false;
    const int H = p->X;
    if (H != 128 && H != 256) return false;
    const bool dae = p->kind == PSNODE_DAE;
    if (p->Z != H) return false;                                   // (the z_dim == 0 script variant runs on the generic kernels)
    if (dae && (p->V != H || p->I != H)) return false;
    const int S = p->X + p->Z + p->V + p->I;
    if (p->de.n_layers != 2 || p->de.in_dim[0] != 3 * S || p->de.out_dim[0] != H || p->de.out_dim[1] != H) return false;
    if (dae && (p->ae.n_layers != 2 || p->ae.in_dim[0] != S + p->X + p->Z + p->V || p->ae.out_dim[0] != H || p->ae.out_dim[1] != H)) return false;
    if (!view_ok(p->z.p, p->z.st, p->z.sb)) return false;
    if (dae && !view_ok(p->v.p, p->v.st, p->v.sb)) return false;
    if (p->event_idx) {
        if (!view_ok(p->z_jump, p->zj_sb, p->zj_se)) return false;
        if (dae && !view_ok(p->v_jump, p->vj_sb, p->vj_se)) return false;
    }
    if ((p->x_sol.sb & 3) || (p->x_sol.st & 3)) return false;
    return lg_encode_fn() != nullptr;
}

int64_t psn_lg_forward_workspace(const psnode_problem* p) { return lg_layout(p).total * 4; }

int psn_lg_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const LgLayout L = lg_layout(p);
    if (ws == nullptr || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    float* w = static_cast<float*>(ws);
    int* err = reinterpret_cast<int*>(w + L.err);
    const bool dae = p->kind == PSNODE_DAE;
    const int H = L.H, B = p->B, T = p->T, E = p->event_idx ? p->E : 0;
    const int S = p->X + p->Z + p->V + p->I;
    const int64_t BH = (int64_t)B * H;
    const int nstages = psw_nstages(p->method);
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    // ---- weights: folded, split into tf32 hi / lo planes ----
    const float* W1 = p->de.W[0];
    const int ld1 = 3 * S;
    auto prep = [&](int which, const float* W, int ldw, int col0, int col1, int K) {
        const int n = H * K;
        psn_lg_prep_kernel<<<(n + 255) / 256, 256, 0, stream>>>(W, ldw, col0, col1, H, K, w + L.wts_hi[which], w + L.wts_lo[which]);
        psn_count_launch("psn_lg_prep_kernel");
    };
    prep(M_DE1X, W1, ld1, S, 2 * S, H);                                       // F_x = (W_b + W_c)[:, 0:X]
    prep(M_DE1ZV, W1, ld1, S + p->X, 2 * S + p->X, L.KZV);                    // [F_z | F_v]
    prep(M_DE2, p->de.W[1], H, 0, -1, H);
    if (dae) {
        prep(M_DE1I, W1, ld1, S + p->X + p->Z + p->V, 2 * S + p->X + p->Z + p->V, H);   // F_i
        const int lda = S + p->X + p->Z + p->V;
        prep(M_AE1X, p->ae.W[0], lda, S, -1, H);
        prep(M_AE1ZV, p->ae.W[0], lda, S + p->X, -1, L.KZV);
        prep(M_AE2, p->ae.W[1], H, 0, -1, H);
        psn_lg_const_kernel<<<(B + 7) / 8, H, 8 * S * 4, stream>>>(p->ae.W[0], lda, -1, p->ae.b[0], p->a0, p->a0_sb, S, B, H, w + L.c_ae);
        psn_count_launch("psn_lg_const_kernel");
    }
    psn_lg_const_kernel<<<(B + 7) / 8, H, 8 * S * 4, stream>>>(W1, ld1, S, p->de.b[0], p->a0, p->a0_sb, S, B, H, w + L.c_de);
    psn_count_launch("psn_lg_const_kernel");
    PSN_CUDA(cudaGetLastError());

    // ---- tensor maps (once per call) ----
    CUtensorMap mw_hi[7], mw_lo[7], m_z, m_v, m_zj, m_vj, m_y, m_a1, m_h, m_i;
    const int kt[7] = {H, L.KZV, H, H, H, L.KZV, H};
    bool ok = true;
    for (int i = 0; i < 7; i++) {
        if (!dae && (i == M_DE1I || i >= M_AE1X)) continue;
        ok = ok && lg_make_map(&mw_hi[i], w + L.wts_hi[i], kt[i], H, kt[i], 1, 0) && lg_make_map(&mw_lo[i], w + L.wts_lo[i], kt[i], H, kt[i], 1, 0);
    }
    ok = ok && lg_make_map(&m_z, p->z.p, H, B, p->z.sb, T, p->z.st);
    if (dae) ok = ok && lg_make_map(&m_v, p->v.p, H, B, p->v.sb, T, p->v.st);
    if (E > 0) {
        ok = ok && lg_make_map(&m_zj, p->z_jump, H, B, p->zj_sb, E, p->zj_se);
        if (dae) ok = ok && lg_make_map(&m_vj, p->v_jump, H, B, p->vj_sb, E, p->vj_se);
    }
    ok = ok && lg_make_map(&m_y, w + L.ycur, H, B, H, 1, 0) && lg_make_map(&m_a1, w + L.a1, H, B, H, 1, 0);
    if (dae) ok = ok && lg_make_map(&m_h, w + L.hbuf, H, B, H, 1, 0) && lg_make_map(&m_i, w + L.icur, H, B, H, 1, 0);
    if (!ok) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled (layer path)");

    const int smem = (int)sizeof(LgSmem) + 1024;
    using KernFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const LgParams);
    if (lg_prepare_kernels() != PSNODE_OK) return PSNODE_ECUDA;
    const LgKernFn* kerns = lg_kernels();
    auto epi_of = [&](const LgParams& q) { return lg_epi_of(q); };
    const int nbt = (B + TN - 1) / TN;
    static const int dbg_cta = std::getenv("PSNODE_LG_DBG") ? std::atoi(std::getenv("PSNODE_LG_DBG")) : 0;
    static const int rotate = std::getenv("PSNODE_LG_ROTATE") ? std::atoi(std::getenv("PSNODE_LG_ROTATE")) : 1;
    auto base = [&]() {
        LgParams q;
        std::memset(&q, 0, sizeof(q));
        q.N = B; q.R = 1; q.nbt = nbt; q.nsrc = 1; q.kchunks = H / 32; q.mode = LG_PLAIN;
        q.add1_ld = H; q.add2_ld = H; q.out_ld = H; q.st_ld = H;
        q.err = err;
        q.dbg = dbg_cta; q.rotate = rotate;
        return q;
    };
    auto launch = [&](int which, const CUtensorMap& b0, const CUtensorMap& b1, const LgParams& q, const char* name) -> int {
        return lg_launch_gemm(mw_hi[which], mw_lo[which], b0, b1, q, H / TM, stream, name);
    };
    // ---- hoisted layer-1 halves over the whole series ----
    auto project = [&](int which, const CUtensorMap& bz, const CUtensorMap& bv, int R, const float* cadd, float* out, const char* name) {
        if (R <= 0) return;
        LgParams q = base();
        q.R = R; q.nsrc = dae ? 2 : 1; q.kchunks = H / 32;
        q.add1 = cadd; q.add1_sr = 0; q.add1_ld = H;
        q.out = out; q.out_sr = BH; q.out_ld = H;
        launch(which, bz, dae ? bv : bz, q, name);
    };
    project(M_DE1ZV, m_z, m_v, T - 1, w + L.c_de, w + L.pre_de, "psn_lg_gemm_kernel<pre_de>");
    if (E > 0) project(M_DE1ZV, m_zj, m_vj, E, w + L.c_de, w + L.pre_de + (int64_t)(T - 1) * BH, "psn_lg_gemm_kernel<pre_de_jump>");
    if (dae) {
        project(M_AE1ZV, m_z, m_v, T, w + L.c_ae, w + L.pre_ae, "psn_lg_gemm_kernel<pre_ae>");
        if (E > 0) project(M_AE1ZV, m_zj, m_vj, E, w + L.c_ae, w + L.pre_ae + (int64_t)T * BH, "psn_lg_gemm_kernel<pre_ae_jump>");
    }
    // ---- initial state ----
    {
        const float* src = dae ? p->x_init : p->x.p;
        const int64_t sb = dae ? p->x_init_sb : p->x.sb;
        psn_lg_init_kernel<<<(int)((BH + 255) / 256), 256, 0, stream>>>(src, sb, B, H, w + L.x0, w + L.ycur, p->x_sol.p, p->x_sol.sb);
        psn_count_launch("psn_lg_init_kernel");
    }
    // algebraic evaluation i = ae(x, z, v) on the current state (ycur holds x at step boundaries)
    auto ae_eval = [&](const float* pre_row, int jrow, bool event_only, int ev_j) {
        LgParams q = base();
        q.mode = LG_HIDDEN;
        q.add1 = pre_row; q.add1_sr = 0;
        if (event_only) {
            q.ev = p->event_idx; q.ev_j = ev_j; q.skip_unless_event = 1;
            q.add1_jump = w + L.pre_ae + (int64_t)T * BH; q.add1_jump_sr = BH;
        }
        q.out = w + L.hbuf;
        launch(M_AE1X, m_y, m_y, q, event_only ? "psn_lg_gemm_kernel<ae1,event>" : "psn_lg_gemm_kernel<ae1>");
        LgParams q2 = base();
        q2.bias = p->ae.b[1];
        if (event_only) { q2.ev = p->event_idx; q2.ev_j = ev_j; q2.skip_unless_event = 1; }
        q2.out = w + L.icur;
        if (jrow >= 0) { q2.out2 = p->i_sol.p + (int64_t)jrow * p->i_sol.st; q2.out2_ld = p->i_sol.sb; }
        launch(M_AE2, m_h, m_h, q2, event_only ? "psn_lg_gemm_kernel<ae2,event>" : "psn_lg_gemm_kernel<ae2>");
    };
    if (dae) ae_eval(w + L.pre_ae, 0, false, 0);                               // i_0 = ae(x_0, z[0], v[0])  (my_solvers.py:95)

    for (int j = 1; j < T; j++) {
        if (dae) {
            if (E > 0) ae_eval(nullptr, -1, true, j - 1);                      // event: i_0 re-evaluated with the jumped inputs (:108-110)
            LgParams qg = base();                                              // G = F_i i0
            qg.out = w + L.G;
            launch(M_DE1I, m_i, m_i, qg, "psn_lg_gemm_kernel<de1_i>");
        }
        for (int e = 0; e < nstages; e++) {
            LgParams q1 = base();
            q1.mode = LG_HIDDEN;
            q1.add1 = w + L.pre_de + (int64_t)(j - 1) * BH;
            if (E > 0) { q1.ev = p->event_idx; q1.ev_j = j - 1; q1.add1_jump = w + L.pre_de + (int64_t)(T - 1) * BH; q1.add1_jump_sr = BH; }
            if (dae) q1.add2 = w + L.G;
            q1.out = w + L.a1;
            launch(M_DE1X, m_y, m_y, q1, "psn_lg_gemm_kernel<de1>");
            LgParams q2 = base();
            q2.mode = LG_RK;
            q2.bias = p->de.b[1];
            q2.method = p->method; q2.stage = e;
            q2.x0 = w + L.x0; q2.k1 = w + L.k1; q2.k2 = w + L.k2; q2.k3 = w + L.k3;
            q2.t_cur = p->t.p + (int64_t)j * p->t.st; q2.t_prev = p->t.p + (int64_t)(j - 1) * p->t.st; q2.t_sb = p->t.sb;
            q2.out = w + L.ycur;
            q2.out2 = p->x_sol.p + (int64_t)j * p->x_sol.st; q2.out2_ld = p->x_sol.sb;
            launch(M_DE2, m_a1, m_a1, q2, "psn_lg_gemm_kernel<de2,rk>");
        }
        if (dae) ae_eval(w + L.pre_ae + (int64_t)j * BH, j, false, 0);          // i_j = ae(x_j, z[j], v[j])  (:121)
    }
    PSN_CUDA(cudaGetLastError());
    if (dbg_cta) {                            // debugging aid only: synchronises
        long long h[64];
        cudaStreamSynchronize(stream);
        cudaMemcpyFromSymbol(h, g_lg_dbg, sizeof(h));
        std::fprintf(stderr, "psn_lg stamps (cycles since kernel start, last launch, CTA %d):", dbg_cta - 1);
        for (int i = 1; i < (int)h[63] && i < 32; i++) std::fprintf(stderr, " %lld", h[i] - h[0]);
        std::fprintf(stderr, "\n");
    }
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
