// psnode_wide.cuh -- shared definitions of the tensor-core path for the LATENT nets of the `*_02_direct_encode` scripts
// (neural_00_ODE_02_direct_encode.py:49-57,70: DE_Func(x_dim = z_dim = hidden_dim = H), two Linear layers
//  L(3*2H -> H) . ELU . L(H -> H); BASELINE configs[3]: H = 128).  Here the stage MLP is a genuine dense GEMM chain.
//
// Folding: with s = [y ; z] (stage state y, zero-order-held input z) and a0 = all_initial,
//     W1 . cat(a0, s - a0, s) + b1 = F_x . y + F_z . z + c,     F = (W_b + W_c),  c = (W_a - W_b) . a0 + b1
// F_z . z + c is constant across the stages of a step and independent of the state, so it is hoisted out of the time loop:
//     pre[r][b][:] = F_z . z[r][b] + c[b]     for every grid row r = 0..T-2 and every event row T-1+k (z_jump[:,k])
// is ONE plain GEMM over the whole series (psnode_wide_proj.cu; its input tiles are staged by TMA), and the time loop
// (psnode_wide_fwd.cu) runs 2 dependent 128x128 GEMMs per stage with both weight matrices resident in TMEM.
//
// All products are 3xTF32 (psnode_tc.cuh): W_lo.a_hi + W_hi.a_lo + W_hi.a_hi, fp32 accumulation in TMEM, K split over 4
// partial accumulators (the tensor core's fp32 accumulate truncates: short chains keep the bias at the fp32 level).
//
// Group = 16 trajectories (MMA N = 16), M = 128 = one neuron / state element per TMEM lane.  Thread (lane m, half h) of a
// group owns elements (m, n = 8h + i), i = 0..7, of every 128 x 16 tile -- accumulators (tcgen05.ld.32x32b.x8), stage
// algebra, trajectory rows and tape blocks alike.
//
// Tape / operand block layout ("K-major over trajectories"): a 128 x 16 block is stored as the canonical no-swizzle UMMA
// tile with rows = m and K = n:  float offset(m, n) = (m/8)*128 + (n/4)*32 + (m%8)*4 + (n%4)   (2048 floats = 8 KB),
// i.e. LBO = 128 B, SBO = 512 B.  A thread's 8 elements are two float4 (n/4 = 2h, 2h+1), and a bulk copy of the 8 KB
// brings the block into shared memory ready to be an MMA operand of the weight-gradient products (psnode_wide_grad.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "psnode_internal.cuh"
#include "psnode_tc.cuh"

constexpr int PSW_H = 128;                 // latent width = state width = held-input width = hidden width
constexpr int PSW_N = 16;                  // trajectories per group
constexpr int PSW_BLOCK = PSW_H * PSW_N;   // floats per 128 x 16 block
constexpr int PSW_GROUPS_PER_CTA = 2;
constexpr int PSW_GROUP_THREADS = 256;
constexpr int PSW_FWD_REC = 2 * PSW_BLOCK; // forward tape record per (group, step, stage): a1 block, y block
constexpr int PSW_BWD_REC = 2 * PSW_BLOCK; // reverse tape record per (group, step, stage): delta2 block, delta1 block
constexpr int PSW_STEP_REC = 2 * PSW_BLOCK;// reverse tape record per (group, step): sum_e delta1 block, held-input z block

static inline int psw_ngroups(int B) { return (B + PSW_N - 1) / PSW_N; }
static inline int psw_nstages(int method) { return method == PSNODE_EULER ? 1 : (method == PSNODE_MIDPOINT ? 2 : 4); }
static inline int64_t psw_bpad(int B) { return (int64_t)psw_ngroups(B) * PSW_N; }
// rows of the hoisted pre-activation buffer: one per step (grid rows 0..T-2) + one per event
static inline int64_t psw_pre_rows(int T, int E) { return (int64_t)(T > 1 ? T - 1 : 0) + (E > 0 ? E : 0); }
static inline int64_t psw_pre_floats(int B, int T, int E) { return psw_pre_rows(T, E) * psw_bpad(B) * PSW_H; }
static inline int64_t psw_tape_floats(int B, int T, int method) {
    return (int64_t)psw_ngroups(B) * (T > 1 ? T - 1 : 0) * psw_nstages(method) * PSW_FWD_REC;
}

// 4-layer ODE_01 net (psnode_wide4_fwd.cu / psnode_wide4_bwd.cu): forward tape record per (group, step, stage) = a1 | a2 | a3 blocks +
// the stage input y as [16 trajectories][16 state rows]; reverse tape record = delta2 | delta3 blocks
constexpr int PSW4_YBLK = PSW_N * 16;
constexpr int PSW4_FWD_REC = 3 * PSW_BLOCK + PSW4_YBLK;
constexpr int PSW4_BWD_REC = 2 * PSW_BLOCK;
static inline int64_t psw4_tape_floats(int B, int T, int method) {
    return (int64_t)psw_ngroups(B) * (T > 1 ? T - 1 : 0) * psw_nstages(method) * PSW4_FWD_REC;
}

__host__ __device__ __forceinline__ int psw_block_off(int m, int n) { return (m >> 3) * 128 + (n >> 2) * 32 + (m & 7) * 4 + (n & 3); }

bool psn_wide_supports(const psnode_problem* p);
int64_t psn_wide_forward_workspace(const psnode_problem* p);
int psn_wide_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
bool psn_wide_bwd_supports(const psnode_problem* p, const psnode_adjoint* a);
int64_t psn_wide_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);
int psn_wide_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
// the 4-layer ODE_01 net at hidden <= 128 (X <= 16, Z <= 8) on the same machinery (psnode_wide4_fwd.cu; `tape` instantiation for the sweep below)
bool psn_wide4_supports(const psnode_problem* p);
bool psn_wide4_auto(const psnode_problem* p);
int64_t psn_wide4_forward_workspace(const psnode_problem* p);
int psn_wide4_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
// its tape-based tensor-core reverse sweep (psnode_wide4_bwd.cu): parameter, x[0] and all_initial gradients
bool psn_wide4_bwd_enabled();
bool psn_wide4_bwd_supports(const psnode_problem* p, const psnode_adjoint* a);
int64_t psn_wide4_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);
int psn_wide4_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
int psn_wide_grad_pairs(const float* a0, int64_t a0_stride, const float* b0, int64_t b0_stride, const float* a1, int64_t a1_stride,
                        const float* b1, int64_t b1_stride, int64_t nrec, int ncta, float* slabs, int* err, int cta0[3], cudaStream_t stream);

// per-layer tcgen05 GEMM launches for the latent nets that do not fit one SM (psnode_lg.cu: DAE_02 / ODE_02 with H = 128 / 256)
bool psn_lg_supports(const psnode_problem* p);
int64_t psn_lg_forward_workspace(const psnode_problem* p);
int psn_lg_forward(const psnode_problem* p, void* ws, int64_t ws_bytes, cudaStream_t stream);
// recomputing reverse sweep on the same GEMM kernel (all parameter, initial-state, all_initial, input-series and jump gradients)
bool psn_lg_bwd_supports(const psnode_problem* p, const psnode_adjoint* a);
int64_t psn_lg_backward_workspace(const psnode_problem* p, const psnode_adjoint* a);
int psn_lg_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream);
// encoders / decoders of the `*_02_direct_encode` models fused with the integration, time chunk by time chunk (psnode_forward_encoded)
bool psn_lg_encoded_supports(const psnode_problem* p, const psnode_codec* c);
int64_t psn_lg_encoded_workspace(const psnode_problem* p, const psnode_codec* c);
int psn_lg_forward_encoded(const psnode_problem* p, const psnode_codec* c, void* ws, int64_t ws_bytes, cudaStream_t stream);
// the same entry for the latent ODE_02 net (X = Z = H = 128) on the wide kernels: generated projection tiles, TMEM-resident time loop
bool psn_wide_encoded_supports(const psnode_problem* p, const psnode_codec* c);
int64_t psn_wide_encoded_workspace(const psnode_problem* p, const psnode_codec* c);
int psn_wide_forward_encoded(const psnode_problem* p, const psnode_codec* c, void* ws, int64_t ws_bytes, cudaStream_t stream);
// decoder Linear(H -> H) . ELU . Linear(H -> width <= 128) over R contiguous latent rows (R, B, H) -> out rows (psnode_lg.cu)
int64_t psn_lg_decode_workspace(int B, int H);
int psn_lg_decode(const psnode_mlp* dec, int width, const float* src, int R, int B, int H, const psnode_series_out* out, void* ws, int64_t ws_bytes,
                  cudaStream_t stream);

// ---- row GEMM over a whole series (psnode_wide_proj.cu) ------------------------------------------------------------------
// out[r][b][0:128] = A . in[r][b][0:128] (+ add[b][0:128]),  A[m][k] = W[m*ldw + k] or (transpose) W[k*ldw + m], optionally
// A = W + W2 (the folded F = W_b + W_c).  `in` is a strided (R, B, 128) view read through a TMA tensor map.
struct PswProjJob {
    const float* in; int64_t in_sr, in_sb;        // element strides between rows r / trajectories b (multiples of 4)
    int R, B;
    const float* W; const float* W2; int ldw; int transpose;
    const float* add; int64_t add_sb;             // nullable
    float* out; int64_t out_sr, out_sb;
    int zero_rows_from;                           // rows r >= this are written as zeros (no GEMM); R if none
    // encoder fusion (psnode_forward_encoded): `in` is not read; its tile is generated in shared memory as the hidden layer of the input
    // encoder, in[r][b][k] = ELU(sum_c gen_W[k][c] * gen_raw[r][b][c] + gen_b[k]), from the raw (R, B, gen_w <= 8) series
    const float* gen_raw = nullptr; int64_t gen_sr = 0, gen_sb = 0; int gen_w = 0;
    const float* gen_W = nullptr; const float* gen_b = nullptr;
};
int psn_wide_proj(const PswProjJob& job, int* err_flag, cudaStream_t stream, const char* name);
// c[b][m] = b1[m] + sum_k (W1[m][k] - W1[m][S + k]) * a0[b][k],  S = 256 (fp32 FMA; B x 128 outputs)
int psn_wide_const(const float* W1, const float* b1, const float* a0, int64_t a0_sb, int B, float* c, cudaStream_t stream);

namespace psn_tc {

// ---- 32x32b TMEM access: thread t of warp w touches lane 32*(w%4) + t, consecutive 32-bit columns --------------------------
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}

// ---- TMA: 1-D bulk copy and 3-D tiled tensor copy, completion on an mbarrier ------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) { asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory"); }

// shared-memory descriptor of a K-major SWIZZLE_128B operand slab (rows x 32 tf32 = 128 B per row, 8-row atoms of 1024 B,
// as written by a TMA box {32, rows} with CU_TENSOR_MAP_SWIZZLE_128B); a K = 8 step inside the slab advances the start by 32 B
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                    // LBO: unused for swizzled K-major layouts
    d |= (uint64_t)(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// explicit round-to-nearest hi/lo split in place for finite values (same arithmetic as split_tf32_fast)
__device__ __forceinline__ float4 split4_hi(const float4 v, float4& lo) {
    float4 hi;
    split_tf32_fast(v.x, hi.x, lo.x); split_tf32_fast(v.y, hi.y, lo.y);
    split_tf32_fast(v.z, hi.z, lo.z); split_tf32_fast(v.w, hi.w, lo.w);
    return hi;
}

}  // namespace psn_tc
