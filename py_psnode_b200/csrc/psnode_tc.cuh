// psnode_tc.cuh -- thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma kind::tf32 with
// shared-memory operands, TMEM allocation / loads, mbarrier completion, and the operand-tile address maps.
//
// Operand tiles are K-major, no swizzle ("interleaved" canonical layout): a tile of R rows x K fp32 columns is a grid of
// 8-row x 16-byte core matrices (8 rows x 4 tf32, 128 contiguous bytes, row r at r*16);  core matrices adjacent in K are
// LBO bytes apart, 8-row groups are SBO bytes apart:
//     byte(row, k) = (row / 8) * SBO + (k / 4) * LBO + (row % 8) * 16 + (k % 4) * 4
// One tcgen05.mma kind::tf32 consumes K = 8 (two core matrices); the descriptor start address advances 2*LBO per k-step.
// Descriptor fields (bit layout as in CUTLASS cute/arch/mma_sm100_desc.hpp, UMMA::SmemDescriptor / InstrDescriptor):
//   smem desc : [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version = 1 | [61,64) layout = 0 (no swizzle)
//   instr desc: [4,6) D fmt (1 = f32) | [7,10) A fmt (2 = tf32) | [10,13) B fmt (2 = tf32) | bit 15/16 A/B major (0 = K)
//               | [17,23) N>>3 | [24,29) M>>4
// Accumulator D (M = 64, cta_group::1) sits in TMEM at lane (m % 16) + 32 * (m / 16), column n: warp w of a warpgroup reads
// its 16 rows with tcgen05.ld.16x256b -- thread t gets rows 16w + t/4 (+8) and columns 2*(t%4) (+1) (+8 per repeat).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace psn_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint32_t tile_byte(int row, int k, int lbo, int sbo) {
    return (uint32_t)((row >> 3) * sbo + (k >> 2) * lbo + (row & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int lbo, int sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// One lane of a fully converged warp; tcgen05.mma / commit issued under this predicate compile to a single UTCHMMA
// (issued from `if (threadIdx.x == 0)` instead, ptxas wraps every MMA in an ELECT/BRA.U.ANY serialisation loop, ~50 cycles).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: returns false if the phase did not complete within ~1e9 SM cycles (about half a second; the caller flags
// an error and bails out instead of hanging the GPU).
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 1000000000ll) return false;
    return true;
}

// ---- TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, both K-major, one K = 8 slice
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when every previously issued tcgen05.mma of this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_16x256b_x1(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr) : "memory");
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                   "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T : A (M = 64 rows at lanes (m%16)+32*(m/16), K = 8 consecutive 32-bit columns) read from TMEM
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest split of an fp32 value into two tf32-representable parts: v ~= hi + lo, |v - hi - lo| <= 2^-22 |v|
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = tf32_rn(v);
    lo = tf32_rn(v - hi);
}
// Same split in 3 instructions for finite inputs (cvt.rna.tf32.f32 is expanded by ptxas into ~8 integer instructions with
// NaN/Inf handling): hi = magnitude rounded half-up to 10 mantissa bits, lo = v - hi exactly (<= 13 significant bits; the
// tensor core ignores the low 13 mantissa bits of a tf32 operand, i.e. truncates lo: |error| <= 2^-21 |v|, sign-symmetric).
__device__ __forceinline__ void split_tf32_fast(float v, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
    lo = v - hi;
}

}  // namespace psn_tc
