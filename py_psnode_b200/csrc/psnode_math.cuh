// psnode_math.cuh -- scalar math shared by every kernel: the ELU activation of the reference's MLPs
// (nn.ELU(), alpha = 1, e.g. neural_00_ODE_01_no_encode.py:61-64) with expm1 accuracy.
//
// torch's CPU ELU evaluates expm1 (F.elu(-1e-8) == -1e-8, SURVEY.md section 7 "ELU numerics"); CUDA's
// expm1f is equally accurate but measured ~33 issue slots per warp on B200 (bench_micro/micro.cu), a
// quarter of a 64-wide GEMV row.  psn_expm1_neg is a 17-instruction Cody-Waite + degree-6 polynomial
// replacement, valid for x <= 0 only (the only branch ELU needs), max error 1 ulp against the correctly rounded
// result, checked exhaustively over every negative float by tests/csrc/check_expm1.c (host build of this
// same header: the sequence uses only IEEE fma/add/mul and integer ops, so host and device agree bit for bit).
#pragma once

#if defined(__CUDACC__)
#define PSN_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define PSN_HD static inline
#endif

PSN_HD float psn_i2f(int v) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}
PSN_HD int psn_f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int v; memcpy(&v, &f, 4); return v;
#endif
}

// expm1(x) for x <= 0.  x = k*ln2 + r, |r| <= ln2/2;  expm1(x) = 2^k * expm1(r) + (2^k - 1).
PSN_HD float psn_expm1_neg(float x) {
    x = fmaxf(x, -17.5f);                                   // exp(-17.5) < 2^-25: result is exactly -1 below this
    float kf = fmaf(x, 1.4426950408889634f, 12582912.0f);   // round-to-nearest via the 1.5*2^23 magic constant
    const int ki = psn_f2i(kf);                             // low mantissa bits hold k (two's complement)
    kf -= 12582912.0f;
    float r = fmaf(kf, -0.693145751953125f, x);             // ln2 split: hi part has 9 trailing zero bits
    r = fmaf(kf, -1.42860682030941723212e-6f, r);
    float q = fmaf(r, 1.9841270e-4f, 1.3888889e-3f);        // expm1(r) = r + r^2 * q(r)
    q = fmaf(q, r, 8.3333333e-3f);
    q = fmaf(q, r, 4.1666668e-2f);
    q = fmaf(q, r, 1.6666667e-1f);
    q = fmaf(q, r, 0.5f);
    const float p = fmaf(r * r, q, r);
    const float t = psn_i2f((int)((unsigned)ki << 23) + 0x3f800000);   // 2^k, k in [-26, 0]
    return fmaf(t, p, t - 1.0f);
}

// ELU(v), alpha = 1, branch free.  The select is written `v <= 0 ? e : v` so that a NaN input comes out as NaN, as in
// torch's nn.ELU (fminf(NaN, 0) = 0 would otherwise turn it into 0 and hide a diverged trajectory from the scripts'
// `loss != loss` guards, utils.py:33-42).
PSN_HD float psn_elu(float v) {
    const float e = psn_expm1_neg(fminf(v, 0.0f));
    return v <= 0.0f ? e : v;
}
// derivative of ELU expressed through its OUTPUT y: 1 for y > 0, y + 1 (= exp(v)) otherwise.
PSN_HD float psn_elu_grad_from_out(float y) { return y > 0.0f ? 1.0f : y + 1.0f; }

#if defined(__CUDACC__)
// ---- packed pairs (Blackwell fma/add/mul.rn.f32x2: two IEEE fp32 operations per issue slot) -------------------------------
// psn_elu2 evaluates psn_elu on two values with exactly the operation sequence of the scalar routine above (every packed
// instruction is two independent round-to-nearest fp32 operations), so its results are bit-identical to psn_elu.
typedef unsigned long long psn_u64;
__device__ __forceinline__ psn_u64 psn_pack2(float x, float y) {
    psn_u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}
__device__ __forceinline__ void psn_unpack2(psn_u64 v, float& x, float& y) { asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ psn_u64 psn_fma2(psn_u64 a, psn_u64 b, psn_u64 c) {
    psn_u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ psn_u64 psn_add2(psn_u64 a, psn_u64 b) {
    psn_u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ psn_u64 psn_mul2(psn_u64 a, psn_u64 b) {
    psn_u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ psn_u64 psn_dup2(float c) { return psn_pack2(c, c); }

__device__ __forceinline__ void psn_elu2(float v0, float v1, float& o0, float& o1) {
    const float x0 = fmaxf(fminf(v0, 0.0f), -17.5f), x1 = fmaxf(fminf(v1, 0.0f), -17.5f);
    const psn_u64 x = psn_pack2(x0, x1);
    psn_u64 kf = psn_fma2(x, psn_dup2(1.4426950408889634f), psn_dup2(12582912.0f));
    float kf0, kf1;
    psn_unpack2(kf, kf0, kf1);
    const int ki0 = __float_as_int(kf0), ki1 = __float_as_int(kf1);
    kf = psn_add2(kf, psn_dup2(-12582912.0f));
    psn_u64 r = psn_fma2(kf, psn_dup2(-0.693145751953125f), x);
    r = psn_fma2(kf, psn_dup2(-1.42860682030941723212e-6f), r);
    psn_u64 q = psn_fma2(r, psn_dup2(1.9841270e-4f), psn_dup2(1.3888889e-3f));
    q = psn_fma2(q, r, psn_dup2(8.3333333e-3f));
    q = psn_fma2(q, r, psn_dup2(4.1666668e-2f));
    q = psn_fma2(q, r, psn_dup2(1.6666667e-1f));
    q = psn_fma2(q, r, psn_dup2(0.5f));
    const psn_u64 p = psn_fma2(psn_mul2(r, r), q, r);
    const float t0 = __int_as_float((int)((unsigned)ki0 << 23) + 0x3f800000), t1 = __int_as_float((int)((unsigned)ki1 << 23) + 0x3f800000);
    const psn_u64 t = psn_pack2(t0, t1);
    const psn_u64 e = psn_fma2(t, p, psn_add2(t, psn_dup2(-1.0f)));
    float e0, e1;
    psn_unpack2(e, e0, e1);
    o0 = v0 <= 0.0f ? e0 : v0;      // NaN propagates (see psn_elu)
    o1 = v1 <= 0.0f ? e1 : v1;
}
#endif
