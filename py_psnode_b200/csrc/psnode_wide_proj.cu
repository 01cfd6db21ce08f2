// psnode_wide_proj.cu -- row GEMM over a whole input series, the hoisted half of layer 1 of the latent `*_02` nets and its
// transpose in the reverse sweep (psnode_wide.cuh):
//     forward : pre[r][b][:] = F_z . z[r][b][:] + c[b][:]          (F_z = (W_b + W_c)[:, X:X+Z], c = (W_a - W_b) a0 + b1)
//     reverse : d_z[r][b][:] = F_z^T . d_pre[r][b][:]
// where z is the (T, B, 128) latent input series of integrate_ODE (neural_dae/my_solvers.py:52-80) as produced by the
// script's z_encoder (neural_00_ODE_02_direct_encode.py:75-78) -- a strided view of batch-major storage.
//
// This is where BASELINE's "stages the trajectory's external-input time series through TMA" applies: the series tile of a
// CTA (64 trajectories x 128 floats of one grid row) is fetched by cp.async.bulk.tensor (UTMALDG) through a 3-D tensor map
// over the strided view (feature, trajectory, row), SWIZZLE_128B, 3 stages deep, so the tile lands in shared memory already
// in the canonical K-major UMMA layout and is the B operand as it is.  Threads only split it into tf32 hi / lo parts in
// place (3xTF32).  Orientation: D[neuron m][trajectory] = A . tile^T with A (hi, lo) resident in TMEM (TS MMAs, M = 128,
// N = 64), K = 128 split over 4 partial accumulators, so a warp's 32 lanes hold 32 consecutive output features of one
// trajectory and every global store is a full 128-byte line.
// Roofline: HBM (512 B read + 512 B written per trajectory-row); 3 x 2 x 128 x 128 tf32 FLOP per row keeps the tensor
// pipe at ~1/3 of the time the bytes need.
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int TB = 64;                       // trajectories per tile (MMA N)
constexpr int NSTAGE = 3;
constexpr int SLAB = TB * 128;               // bytes of one 32-feature slab: TB rows x 128 B
constexpr int TILE_BYTES = 4 * SLAB;         // 32 KB
constexpr int WORK_THREADS = 256;            // 8 warps: split the tiles, run the epilogue
constexpr int PROJ_THREADS = WORK_THREADS + 64;   // + TMA producer warp + MMA issuer warp
constexpr int NPP = 2;                       // K-partials per accumulator buffer (two buffers: the epilogue of tile i overlaps the MMAs of tile i + 1)
constexpr int TM_A_HI = 0, TM_A_LO = 128, TM_ACC = 256;

struct __align__(1024) ProjSmem {
    unsigned char hi[NSTAGE][TILE_BYTES];    // TMA destination (raw fp32), overwritten in place by the tf32 hi parts
    unsigned char lo[NSTAGE][TILE_BYTES];
    uint64_t full[NSTAGE];                   // producer -> workers (transaction count)
    uint64_t split[NSTAGE];                  // workers -> issuer (8 warp arrivals)
    uint64_t done[NSTAGE];                   // issuer (tcgen05.commit) -> producer (stage free) and workers (accumulator complete)
    uint64_t drained[2];                     // workers -> issuer: accumulator buffer may be overwritten
    uint32_t tmem_base;
};

struct ProjParams {
    int R, B, nbt;
    const float* W; const float* W2; int ldw; int transpose;
    const float* add; int64_t add_sb;
    float* out; int64_t out_sr, out_sb;
    const float* gen_raw; int64_t gen_sr, gen_sb; int gen_w;     // generated tiles (encoder fusion); gen_raw == NULL: TMA tiles
    const float* gen_W; const float* gen_b;
    int* err;
};

// Warp roles: warps 0..7 split the tiles and run the epilogue, warp 8 is the TMA producer, warp 9 the MMA issuer; each role waits
// only on the mbarrier of the role before it.  (Version 1 ran load-wait -> split -> CTA barrier -> MMA -> wait -> epilogue -> CTA
// barrier serially per tile: 36 % of the copy bandwidth.)
__global__ void __launch_bounds__(PROJ_THREADS, 1) psn_wide_proj_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ProjParams q) {
    extern __shared__ unsigned char smem_raw[];
    ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = cw & 3, hh = (cw >> 2) & 1;
    const int ntiles = q.R * q.nbt;
    const int my_n = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.split[s], WORK_THREADS / 32); mbar_init(&sm.done[s], 1); }
        mbar_init(&sm.drained[0], WORK_THREADS / 32); mbar_init(&sm.drained[1], WORK_THREADS / 32);
        fence_mbar_init();
        if (!q.gen_raw) prefetch_tmap(&tmap);
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;

    // ---- A (hi, lo) -> TMEM: lane m = output feature, columns = reduction index -----------------------------------------
    if (cw < WORK_THREADS / 32) {
        const int m = 32 * wq + lane;
        for (int ch = 0; ch < 8; ch++) {
            const int k0 = 64 * hh + 8 * ch;
            float hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int k = k0 + i;
                const int64_t off = q.transpose ? (int64_t)k * q.ldw + m : (int64_t)m * q.ldw + k;
                float w = __ldg(q.W + off);
                if (q.W2) w += __ldg(q.W2 + off);
                split_tf32(w, hi[i], lo[i]);
            }
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_HI + k0, hi);
            tmem_st_32x32b_x8(tmem + lane_base + TM_A_LO + k0, lo);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (cw == WORK_THREADS / 32) {
        // ---- TMA producer (idle when the tiles are generated) ----
        if (!q.gen_raw && elect_one()) {
            for (int i = 0; i < my_n; i++) {
                const int s = i % NSTAGE;
                if (i >= NSTAGE && !mbar_wait(&sm.done[s], (uint32_t)(((i - NSTAGE) / NSTAGE) & 1))) { atomicExch(q.err, 4); __trap(); }
                const int tile = blockIdx.x + i * gridDim.x;
                const int r = tile / q.nbt, b0 = (tile - r * q.nbt) * TB;
                mbar_expect_tx(&sm.full[s], TILE_BYTES);
#pragma unroll
                for (int cb = 0; cb < 4; cb++) tma_load_3d(sm.hi[s] + cb * SLAB, &tmap, cb * 32, b0, r, &sm.full[s]);
            }
        }
        __syncwarp();
    } else if (cw == WORK_THREADS / 32 + 1) {
        // ---- MMA issuer: partial accumulator p of buffer (i & 1) <- K slabs 2p, 2p + 1 (32 features each) ----
        const uint32_t idesc = make_idesc_tf32(128, TB);
        for (int i = 0; i < my_n; i++) {
            const int s = i % NSTAGE;
            if (i >= 2 && !mbar_wait(&sm.drained[i & 1], (uint32_t)(((i - 2) >> 1) & 1))) { atomicExch(q.err, 5); __trap(); }
            if (!mbar_wait(&sm.split[s], (uint32_t)((i / NSTAGE) & 1))) { atomicExch(q.err, 6); __trap(); }
            if (elect_one()) {
                tc_fence_after();
                const uint64_t d_hi = make_desc_sw128(smem_u32(sm.hi[s])), d_lo = make_desc_sw128(smem_u32(sm.lo[s]));
                const uint32_t acc0 = tmem + TM_ACC + (uint32_t)((i & 1) * NPP * TB);
#pragma unroll 1
                for (int p = 0; p < NPP; p++) {
                    uint32_t accumulate = 0;
#pragma unroll
                    for (int term = 0; term < 3; term++) {        // small terms first
#pragma unroll
                        for (int sl = 0; sl < 2; sl++) {
                            const int slab = 2 * p + sl;
                            const uint32_t a_col = (term == 0 ? TM_A_LO : TM_A_HI) + 32 * slab;
                            const uint64_t bd = (term == 1 ? d_lo : d_hi) + (uint64_t)((slab * SLAB) >> 4);
#pragma unroll
                            for (int kk = 0; kk < 4; kk++) {
                                mma_tf32_ts(acc0 + TB * p, tmem + a_col + 8 * kk, bd + (uint64_t)(2 * kk), idesc, accumulate);
                                accumulate = 1;
                            }
                        }
                    }
                }
                mma_commit(&sm.done[s]);
            }
            __syncwarp();
        }
    } else {
        // ---- workers: split tile i, then the epilogue of tile i - 1 while the MMAs of tile i run ----
        const int m = 32 * wq + lane;
        float cadd_cur[32], cadd_prev[32];
        int r_prev = 0, b0_prev = 0;
        auto load_cadd = [&](int b0, float (&c)[32]) {
#pragma unroll
            for (int e = 0; e < 32; e++) {
                const int b = min(b0 + 32 * hh + e, q.B - 1);
                c[e] = q.add ? __ldg(q.add + (int64_t)b * q.add_sb + m) : 0.0f;
            }
        };
        // sum the partials, add the per-trajectory constant, store 128-byte lines
        auto epilogue = [&](int pi, int r, int b0, const float (&c)[32]) {
            if (!mbar_wait(&sm.done[pi % NSTAGE], (uint32_t)((pi / NSTAGE) & 1))) { atomicExch(q.err, 3); __trap(); }
            tc_fence_after();
            const uint32_t acc0 = tmem + lane_base + TM_ACC + (uint32_t)((pi & 1) * NPP * TB);
#pragma unroll
            for (int ch = 0; ch < 4; ch++) {
                const int n0 = 32 * hh + 8 * ch;
                float t0[8], t1[8];
                tmem_ld_32x32b_x8(acc0 + n0, t0);
                tmem_ld_32x32b_x8(acc0 + TB + n0, t1);
                tmem_ld_wait();
#pragma unroll
                for (int i2 = 0; i2 < 8; i2++) {
                    const int b = b0 + n0 + i2;
                    if (b < q.B) q.out[(int64_t)r * q.out_sr + (int64_t)b * q.out_sb + m] = (t0[i2] + t1[i2]) + c[8 * ch + i2];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.drained[pi & 1]);
        };
        for (int i = 0; i < my_n; i++) {
            const int s = i % NSTAGE;
            const int tile = blockIdx.x + i * gridDim.x;
            const int r = tile / q.nbt, b0 = (tile - r * q.nbt) * TB;
            // the per-trajectory constants of this thread's 32 outputs are fetched before anything is waited on (they were on the
            // critical path of the epilogue: long_scoreboard 9.4 -> 1.5 warps per issue cycle, 2.09 -> 0.9 ms at the cfg4 shard)
            load_cadd(b0, cadd_cur);
            if (q.gen_raw) {
                // Encoder fusion: the tile is computed here instead of fetched, in the layout TMA delivers (slab cb = features 32 cb ..,
                // 128-byte row n, 16-byte unit u at position u ^ (n & 7)).  Stage s is free: every warp ran the epilogue of tile i - NSTAGE,
                // i.e. saw its MMAs complete, two iterations ago.
                float4* h4 = reinterpret_cast<float4*>(sm.hi[s]);
                float4* l4 = reinterpret_cast<float4*>(sm.lo[s]);
                float raw[2][8];
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    const int b = min(b0 + (tid >> 3) + 32 * hf, q.B - 1);
#pragma unroll
                    for (int cc = 0; cc < 8; cc++) raw[hf][cc] = cc < q.gen_w ? __ldg(q.gen_raw + (int64_t)r * q.gen_sr + (int64_t)b * q.gen_sb + cc) : 0.0f;
                }
#pragma unroll
                for (int e = 0; e < TILE_BYTES / 16 / WORK_THREADS; e++) {
                    const int idx = tid + e * WORK_THREADS;
                    const int cb = idx >> 9, n = (idx & 511) >> 3, k0 = 32 * cb + (((idx & 7) ^ (n & 7)) << 2);
                    float v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        float pre = __ldg(q.gen_b + k0 + u);
#pragma unroll
                        for (int cc = 0; cc < 8; cc++)
                            if (cc < q.gen_w) pre = fmaf(__ldg(q.gen_W + (k0 + u) * q.gen_w + cc), raw[e & 1][cc], pre);
                        v[u] = psn_elu(pre);
                    }
                    float4 lo;
                    const float4 hi = split4_hi(make_float4(v[0], v[1], v[2], v[3]), lo);
                    h4[idx] = hi;
                    l4[idx] = lo;
                }
            } else {
            if (!mbar_wait(&sm.full[s], (uint32_t)((i / NSTAGE) & 1))) { atomicExch(q.err, 2); __trap(); }
            {   // raw fp32 tile -> tf32 hi (in place) and lo parts; elementwise, so the swizzled layout is preserved
                float4* h4 = reinterpret_cast<float4*>(sm.hi[s]);
                float4* l4 = reinterpret_cast<float4*>(sm.lo[s]);
#pragma unroll
                for (int e = 0; e < TILE_BYTES / 16 / WORK_THREADS; e++) {
                    const int idx = tid + e * WORK_THREADS;
                    float4 lo;
                    const float4 hi = split4_hi(h4[idx], lo);
                    h4[idx] = hi;
                    l4[idx] = lo;
                }
            }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.split[s]);
            if (i >= 1) epilogue(i - 1, r_prev, b0_prev, cadd_prev);
#pragma unroll
            for (int e = 0; e < 32; e++) cadd_prev[e] = cadd_cur[e];
            r_prev = r; b0_prev = b0;
        }
        if (my_n > 0) epilogue(my_n - 1, r_prev, b0_prev, cadd_prev);
    }
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 512);
}

// CUDA-core version of the same job (fp32 FMA): the independent cross-check of the TMA / tensor-core kernel and its A/B partner
// (PSNODE_WIDE_PROJ=simple).  One CTA = 16 trajectories of one row; thread m keeps 16 accumulators.
__global__ void __launch_bounds__(128) psn_wide_proj_simple_kernel(const float* __restrict__ in, int64_t in_sr, int64_t in_sb, ProjParams q) {
    extern __shared__ float sms[];
    float* AT = sms;                    // [k][m]
    float* tile = sms + 128 * 128;      // [n][k]
    const int m = threadIdx.x;
    for (int k = 0; k < 128; k++) {
        const int64_t off = q.transpose ? (int64_t)k * q.ldw + m : (int64_t)m * q.ldw + k;
        float w = __ldg(q.W + off);
        if (q.W2) w += __ldg(q.W2 + off);
        AT[k * 128 + m] = w;
    }
    const int nb16 = (q.B + 15) / 16;
    for (int item = blockIdx.x; item < q.R * nb16; item += gridDim.x) {
        const int r = item / nb16, b0 = (item - r * nb16) * 16;
        __syncthreads();
        for (int n = 0; n < 16; n++) {
            const int b = min(b0 + n, q.B - 1);
            tile[n * 128 + m] = __ldg(in + (int64_t)r * in_sr + (int64_t)b * in_sb + m);
        }
        __syncthreads();
        float acc[16];
#pragma unroll
        for (int n = 0; n < 16; n++) acc[n] = 0.0f;
        for (int k = 0; k < 128; k++) {
            const float a = AT[k * 128 + m];
#pragma unroll
            for (int n = 0; n < 16; n++) acc[n] = fmaf(a, tile[n * 128 + k], acc[n]);
        }
        for (int n = 0; n < 16; n++) {
            const int b = b0 + n;
            if (b < q.B) {
                float v = acc[n];
                if (q.add) v += __ldg(q.add + (int64_t)b * q.add_sb + m);
                q.out[(int64_t)r * q.out_sr + (int64_t)b * q.out_sb + m] = v;
            }
        }
    }
}

// c[b][m] = b1[m] + sum_k (W_a - W_b)[m][k] a0[b][k]   (K = S = 256); one CTA per 8 trajectories, thread = neuron
__global__ void __launch_bounds__(128) psn_wide_const_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                                               const float* __restrict__ a0, int64_t a0_sb, int B, float* __restrict__ c) {
    constexpr int S = 2 * PSW_H;
    __shared__ float a[8][S];
    const int m = threadIdx.x, b0 = blockIdx.x * 8;
    for (int e = m; e < 8 * S; e += 128) {
        const int n = e / S, k = e - n * S;
        a[n][k] = __ldg(a0 + (int64_t)min(b0 + n, B - 1) * a0_sb + k);
    }
    __syncthreads();
    float acc[8];
    const float bias = __ldg(b1 + m);
#pragma unroll
    for (int n = 0; n < 8; n++) acc[n] = bias;
    for (int k = 0; k < S; k++) {
        const float w = __ldg(W1 + (int64_t)m * 3 * S + k) - __ldg(W1 + (int64_t)m * 3 * S + S + k);
#pragma unroll
        for (int n = 0; n < 8; n++) acc[n] = fmaf(w, a[n][k], acc[n]);
    }
    for (int n = 0; n < 8; n++)
        if (b0 + n < B) c[(int64_t)(b0 + n) * PSW_H + m] = acc[n];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

bool use_simple_proj() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("PSNODE_WIDE_PROJ"); v = (e && std::strcmp(e, "simple") == 0) ? 1 : 0; }
    return v == 1;
}

}  // namespace

bool psn_wide_proj_view_ok(const float* p, int64_t sr, int64_t sb) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (sr & 3) == 0 && (sb & 3) == 0 && sr > 0 && sb > 0;
}

int psn_wide_proj(const PswProjJob& job, int* err_flag, cudaStream_t stream, const char* name) {
    if (job.R <= 0 || job.B <= 0) return PSNODE_OK;
    ProjParams q;
    q.R = job.R; q.B = job.B; q.nbt = (job.B + TB - 1) / TB;
    q.W = job.W; q.W2 = job.W2; q.ldw = job.ldw; q.transpose = job.transpose;
    q.add = job.add; q.add_sb = job.add_sb;
    q.out = job.out; q.out_sr = job.out_sr; q.out_sb = job.out_sb;
    q.gen_raw = job.gen_raw; q.gen_sr = job.gen_sr; q.gen_sb = job.gen_sb; q.gen_w = job.gen_w; q.gen_W = job.gen_W; q.gen_b = job.gen_b;
    q.err = err_flag;
    int dev = 0, sms = 148;
    PSN_CUDA(cudaGetDevice(&dev));
    PSN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!job.gen_raw && (use_simple_proj() || !psn_wide_proj_view_ok(job.in, job.in_sr, job.in_sb) || encode_fn() == nullptr)) {
        const int smem = (128 * 128 + 16 * 128) * 4;
        PSN_CUDA(cudaFuncSetAttribute(psn_wide_proj_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        psn_wide_proj_simple_kernel<<<sms * 2, 128, smem, stream>>>(job.in, job.in_sr, job.in_sb, q);
        psn_count_launch("psn_wide_proj_simple_kernel");
        PSN_CUDA(cudaGetLastError());
        return PSNODE_OK;
    }
    CUtensorMap tmap;
    std::memset(&tmap, 0, sizeof(tmap));
    if (!job.gen_raw) {
    const cuuint64_t gdim[3] = {128, (cuuint64_t)job.B, (cuuint64_t)job.R};
    const cuuint64_t gstr[2] = {(cuuint64_t)job.in_sb * 4, (cuuint64_t)job.in_sr * 4};
    const cuuint32_t box[3] = {32, TB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(job.in), gdim, gstr, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return psn_cuda_fail(cudaErrorInvalidValue, "cuTensorMapEncodeTiled");
    }
    const int ntiles = q.R * q.nbt;
    const int grid = ntiles < sms ? ntiles : sms;
    const int smem = (int)sizeof(ProjSmem) + 1024;
    PSN_CUDA(cudaFuncSetAttribute(psn_wide_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    psn_wide_proj_kernel<<<grid, PROJ_THREADS, smem, stream>>>(tmap, q);
    psn_count_launch(name);
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

int psn_wide_const(const float* W1, const float* b1, const float* a0, int64_t a0_sb, int B, float* c, cudaStream_t stream) {
    psn_wide_const_kernel<<<(B + 7) / 8, 128, 0, stream>>>(W1, b1, a0, a0_sb, B, c);
    psn_count_launch("psn_wide_const_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
