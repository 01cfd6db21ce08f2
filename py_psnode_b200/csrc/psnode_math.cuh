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

// ELU(v), alpha = 1, branch free.
PSN_HD float psn_elu(float v) {
    const float e = psn_expm1_neg(fminf(v, 0.0f));
    return v > 0.0f ? v : e;
}
// derivative of ELU expressed through its OUTPUT y: 1 for y > 0, y + 1 (= exp(v)) otherwise.
PSN_HD float psn_elu_grad_from_out(float y) { return y > 0.0f ? 1.0f : y + 1.0f; }
