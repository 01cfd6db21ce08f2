// tc_probe.cu -- validates the tcgen05 kind::tf32 operand layouts / descriptors used by the tensor-core integrator and
// measures (a) the numerical error of 1xTF32 and 3xTF32 products against fp64 and (b) the latency of one dependent
// "layer" round trip (24 MMAs -> commit -> mbarrier wait -> tcgen05.ld -> operand rewrite -> fences).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_probe tc_probe.cu ; run on a B200.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../py_psnode_b200/csrc/psnode_tc.cuh"

using namespace psn_tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 64, K = 64;
constexpr int LBO_A = 128, SBO_A = (K / 4) * LBO_A;
constexpr int LBO_B = 144;

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ W, const float* __restrict__ A, float* __restrict__ out1,
                                                    float* __restrict__ out3, int iters, long long* cycles, int* err, int nterms, int nacc) {
    constexpr int SBO_B = (K / 4) * LBO_B;
    extern __shared__ __align__(128) unsigned char smem[];
    float* whi = reinterpret_cast<float*>(smem);
    float* wlo = whi + M * K;
    unsigned char* ahi = reinterpret_cast<unsigned char*>(wlo + M * K);
    unsigned char* alo = ahi + (N / 8) * SBO_B;
    __shared__ __align__(8) uint64_t bar, bar4;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = tid; e < M * K; e += 128) {
        const int m = e / K, k = e % K;
        float hi, lo;
        split_tf32(W[e], hi, lo);
        whi[tile_byte(m, k, LBO_A, SBO_A) / 4] = hi;
        wlo[tile_byte(m, k, LBO_A, SBO_A) / 4] = lo;
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        float hi, lo;
        split_tf32(A[e], hi, lo);
        *reinterpret_cast<float*>(ahi + tile_byte(n, k, LBO_B, SBO_B)) = hi;
        *reinterpret_cast<float*>(alo + tile_byte(n, k, LBO_B, SBO_B)) = lo;
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar4, 4); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc_tf32(M, N);
    const uint64_t dwhi = make_desc(smem_u32(whi), LBO_A, SBO_A), dwlo = make_desc(smem_u32(wlo), LBO_A, SBO_A);
    const uint64_t dahi = make_desc(smem_u32(ahi), LBO_B, SBO_B), dalo = make_desc(smem_u32(alo), LBO_B, SBO_B);
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    {   // weights -> TMEM columns [128,192) = hi, [192,256) = lo ; thread t of warp w owns rows 16w + t/4 (+8), cols 2(t%4)+{0,1} (+8)
        const int r0 = 16 * warp + lane / 4, c0 = 2 * (lane % 4);
        for (int half = 0; half < 2; half++)
            for (int cb = 0; cb < K / 16; cb++) {
                float w8[8];
                for (int i = 0; i < 8; i++) {
                    const int row = r0 + ((i >> 1) & 1) * 8, col = 16 * cb + c0 + (i & 1) + (i >> 2) * 8;
                    float hi, lo;
                    split_tf32(W[row * K + col], hi, lo);
                    w8[i] = half ? lo : hi;
                }
                tmem_st_16x256b_x2(taddr + 128 + 64 * half + 16 * cb, w8);
            }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    uint32_t phase = 0;
    bool ok = true;
    float v[N / 2 < 4 ? 4 : N / 2];

    auto load_d = [&]() {
        if constexpr (N == 8) { float t4[4]; tmem_ld_16x256b_x1(taddr, t4); tmem_ld_wait(); for (int i = 0; i < 4; i++) v[i] = t4[i]; }
        else {
#pragma unroll
            for (int cb = 0; cb < N / 16; cb++) { float t8[8]; tmem_ld_16x256b_x2(taddr + 16 * cb, t8); for (int i = 0; i < 8; i++) v[8 * cb + i] = t8[i]; }
            tmem_ld_wait();
        }
    };
    auto store_d = [&](float* out) {
        const int m0 = 16 * warp + lane / 4, c = 2 * (lane % 4);
#pragma unroll
        for (int r = 0; r < N / 8; r++) {
            out[(m0) * N + 8 * r + c] = v[4 * r + 0];
            out[(m0) * N + 8 * r + c + 1] = v[4 * r + 1];
            out[(m0 + 8) * N + 8 * r + c] = v[4 * r + 2];
            out[(m0 + 8) * N + 8 * r + c + 1] = v[4 * r + 3];
        }
    };
    // ---- 1xTF32: raw fp32 bits fed to the tensor core (shows what the hardware does with the low 13 mantissa bits) ----
    // (re-uses the hi tiles: hi is already tf32-representable, so this is the plain single-pass product)
    if (tid == 0) {
        for (int ks = 0; ks < K / 8; ks++)
            mma_tf32(tmem, dwhi + (uint64_t)((ks * 2 * LBO_A) >> 4), dahi + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc, ks > 0);
        mma_commit(&bar);
    }
    ok = mbar_wait(&bar, phase); phase ^= 1;
    tc_fence_after();
    if (ok) { load_d(); store_d(out1); }
    tc_fence_before();
    __syncthreads();
    // ---- 3xTF32: small terms first ----
    auto issue3 = [&]() {
        int cnt = 0;
        for (int term = 0; term < 3; term++) {
            const uint64_t da = term == 0 ? dwlo : dwhi, db = term == 1 ? dalo : dahi;
            for (int ks = 0; ks < K / 8; ks++) {
                // nacc independent accumulators (TMEM column blocks of N) break the D -> D dependency chain
                mma_tf32(tmem + (cnt % nacc) * N, da + (uint64_t)((ks * 2 * LBO_A) >> 4), db + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc,
                         cnt < nacc ? 0 : 1);
                cnt++;
            }
        }
        mma_commit(&bar);
    };
    if (ok && tid == 0) { tc_fence_after(); issue3(); }
    if (ok) { ok = mbar_wait(&bar, phase); phase ^= 1; }
    tc_fence_after();
    if (ok) { load_d(); store_d(out3); }
    tc_fence_before();
    __syncthreads();
    // ---- latency of a dependent layer round trip, with a per-phase breakdown taken by thread 0 ----
    long long acc[6] = {0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters && ok; it++) {
        const long long s0 = clock64();
        if (nacc == -6) {       // ONE issuing warp, weights in TMEM, all 24 MMAs fully unrolled into ONE accumulator (round 2:
                                // re-measures the single-issuer cost after the elect_one + uniform-descriptor fixes)
            if (warp == 0) {
                if (elect_one()) {
                    tc_fence_after();
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint32_t ta = tmem + 128 + (term == 0 ? 64 : 0);
                        const uint64_t db = term == 1 ? dalo : dahi;
#pragma unroll
                        for (int ks = 0; ks < K / 8; ks++)
                            mma_tf32_ts(tmem, ta + 8 * ks, db + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc, (term | ks) ? 1 : 0);
                    }
                    mma_commit(&bar);
                }
                __syncwarp();
            }
        } else if (nacc == -5) {       // as -4, but the weights (A operand) live in TMEM: only B is fetched from shared memory
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint32_t ta = tmem + 128 + (term == 0 ? 64 : 0);
                    const uint64_t db = term == 1 ? dalo : dahi;
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const int ks = 2 * warp + kk;
                        mma_tf32_ts(tmem + warp * N, ta + 8 * ks, db + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc, (term | kk) ? 1 : 0);
                    }
                }
                mma_commit(&bar4);
            }
            __syncwarp();
        } else if (nacc == -4) {       // every warp issues its own 2 k-steps x 3 terms into its own accumulator; bar expects 4 commits
            if (elect_one()) {
                tc_fence_after();
#pragma unroll
                for (int term = 0; term < 3; term++) {
                    const uint64_t da = term == 0 ? dwlo : dwhi, db = term == 1 ? dalo : dahi;
#pragma unroll
                    for (int kk = 0; kk < 2; kk++) {
                        const int ks = 2 * warp + kk;
                        mma_tf32(tmem + warp * N, da + (uint64_t)((ks * 2 * LBO_A) >> 4), db + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc,
                                 (term | kk) ? 1 : 0);
                    }
                }
                mma_commit(&bar4);
            }
            __syncwarp();
        } else if (warp == 0 && elect_one()) {
            tc_fence_after();
            if (nterms == 3) issue3();
            else {
                for (int ks = 0; ks < K / 8; ks++)
                    mma_tf32(tmem, dwhi + (uint64_t)((ks * 2 * LBO_A) >> 4), dahi + (uint64_t)((ks * 2 * LBO_B) >> 4), idesc, ks > 0);
                mma_commit(&bar);
            }
        }
        const long long s1 = clock64();
        ok = mbar_wait((nacc == -4 || nacc == -5) ? &bar4 : &bar, phase); phase ^= 1;
        tc_fence_after();
        const long long s2 = clock64();
        load_d();
        if (nacc == -4 || nacc == -5) {       // sum the 4 partial accumulators
            float vv[N / 2 < 4 ? 4 : N / 2];
            for (int i = 0; i < (N / 2 < 4 ? 4 : N / 2); i++) vv[i] = v[i];
            for (int a = 1; a < 4; a++) {
                if constexpr (N == 8) { float t4[4]; tmem_ld_16x256b_x1(taddr + a * N, t4); tmem_ld_wait(); for (int i = 0; i < 4; i++) vv[i] += t4[i]; }
                else {
#pragma unroll
                    for (int cb = 0; cb < N / 16; cb++) { float t8[8]; tmem_ld_16x256b_x2(taddr + a * N + 16 * cb, t8); tmem_ld_wait(); for (int i = 0; i < 8; i++) vv[8 * cb + i] += t8[i]; }
                }
            }
            for (int i = 0; i < (N / 2 < 4 ? 4 : N / 2); i++) v[i] = vv[i];
            if (it == 0) store_d(out3);
        }
        const long long s3 = clock64();
        // rewrite the B operand like the real epilogue does (values kept bounded)
        const int m0 = 16 * warp + lane / 4, c = 2 * (lane % 4);
#pragma unroll
        for (int r = 0; r < N / 8; r++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int m = m0 + (q >> 1) * 8, n = 8 * r + c + (q & 1);
                float hi, lo;
                split_tf32(v[4 * r + q] * 0.05f, hi, lo);
                *reinterpret_cast<float*>(ahi + tile_byte(n, m, LBO_B, SBO_B)) = hi;
                *reinterpret_cast<float*>(alo + tile_byte(n, m, LBO_B, SBO_B)) = lo;
            }
        const long long s4 = clock64();
        fence_async_smem();
        tc_fence_before();
        const long long s5 = clock64();
        __syncthreads();
        const long long s6 = clock64();
        acc[0] += s1 - s0; acc[1] += s2 - s1; acc[2] += s3 - s2; acc[3] += s4 - s3; acc[4] += s5 - s4; acc[5] += s6 - s5;
    }
    long long t1 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; for (int i = 0; i < 6; i++) cycles[1 + i] = acc[i]; if (!ok) err[0] = 1; }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int N>
void run(int iters, int nterms, int nacc = 1) {
    std::vector<float> W(M * K), A(N * K);
    srand(1);
    for (auto& x : W) x = (rand() / (float)RAND_MAX - 0.5f) * 0.25f;
    for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    float *dW, *dA, *d1, *d3; long long* dc; int* de;
    CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&d1, M * N * 4)); CK(cudaMalloc(&d3, M * N * 4));
    CK(cudaMalloc(&dc, 64)); CK(cudaMalloc(&de, 4)); CK(cudaMemset(de, 0, 4));
    CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    const int smem = 2 * M * K * 4 + 2 * (N / 8) * (K / 4) * LBO_B;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<N><<<1, 128, smem>>>(dW, dA, d1, d3, iters, dc, de, nterms, nacc);
    CK(cudaDeviceSynchronize());
    std::vector<float> o1(M * N), o3(M * N); long long cyc[7]; int err;
    CK(cudaMemcpy(o1.data(), d1, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o3.data(), d3, M * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc, dc, 56, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
    double e1 = 0, e3 = 0, ef = 0, b3 = 0, scale = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double ref = 0; float f = 0;
            for (int k = 0; k < K; k++) { ref += (double)W[m * K + k] * (double)A[n * K + k]; f = fmaf(W[m * K + k], A[n * K + k], f); }
            e1 = fmax(e1, fabs(o1[m * N + n] - ref)); e3 = fmax(e3, fabs(o3[m * N + n] - ref)); ef = fmax(ef, fabs((double)f - ref));
            b3 += (o3[m * N + n] - ref) * (ref >= 0 ? 1 : -1); scale = fmax(scale, fabs(ref));
        }
    printf("N=%2d  timeout=%d  max|ref|=%.3f  err 1xTF32 %.3e  3xTF32 %.3e  fp32-fma %.3e  mean signed 3x err (toward +|ref|) %.3e   layer round trip %.1f cycles (%d iters)\n",
           N, err, scale, e1, e3, ef, b3 / (M * N), iters ? (double)cyc[0] / iters : 0.0, iters);
    printf("      nacc=%d terms=%d  thread-0 breakdown (cycles/iter): issue %.0f | wait %.0f | tmem ld %.0f | split+sts %.0f | fences %.0f | syncthreads %.0f\n", nacc, nterms,
           (double)cyc[1] / iters, (double)cyc[2] / iters, (double)cyc[3] / iters, (double)cyc[4] / iters, (double)cyc[5] / iters, (double)cyc[6] / iters);
}

int main() {
    run<8>(2000, 3);
    run<16>(2000, 3);
    run<32>(2000, 3);
    run<16>(2000, 1);
    run<16>(2000, 3, 2);
    run<16>(2000, 3, 4);
    run<32>(2000, 3, 4);
    run<8>(2000, 3, 4);
    run<8>(2000, 3, -4);
    run<16>(2000, 3, -4);
    run<32>(2000, 3, -4);
    run<8>(2000, 3, -5);
    run<16>(2000, 3, -5);
    run<32>(2000, 3, -5);
    run<8>(2000, 3, -6);
    run<16>(2000, 3, -6);
    run<32>(2000, 3, -6);
    return 0;
}
