// Microbenchmarks that decide the thread mapping of the fused RK4/MLP kernel (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o micro micro.cu
// Not part of the product path; results are recorded in DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float2 unpack(u64 v) {
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64 pack(float x, float y) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y));
    return r;
}

// ---------------------------------------------------------------- A: scalar FFMA throughput
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    float x = a + threadIdx.x * 1e-6f, y = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct register operands per FFMA (GEMM-like: acc += a[m]*b[n])
__global__ void k_ffma_tile(float* out, int iters, float a0, float b0) {
    float acc[4][4];
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = a0 + threadIdx.x * 1e-6f + i; b[i] = b0 + i * 0.5f; }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 4; i++) { a[i] += 1e-9f; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---------------------------------------------------------------- B: packed f32x2 FMA throughput
__global__ void k_ffma2(float* out, int iters, float a, float b) {
    u64 acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = pack(threadIdx.x * 0.001f + i, i);
    u64 x = pack(a + threadIdx.x * 1e-6f, a), y = pack(b, b * 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma2(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { float2 f = unpack(acc[i]); s += f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2_tile(float* out, int iters, float a0, float b0) {
    u64 acc[4][4];
    u64 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = pack(a0 + threadIdx.x * 1e-6f + i, a0); b[i] = pack(b0 + i * 0.5f, b0); }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = pack(i, j);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fma2(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 4; i++) { a[i] ^= (u64)it; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { float2 f = unpack(acc[i][j]); s += f.x + f.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---------------------------------------------------------------- C/D/E: LDS throughput
// mode 0: LDS.128 broadcast (all lanes same address); 1: LDS.128 distinct conflict-free;
// 2: LDS.32 broadcast; 3: LDS.64 broadcast; 4: LDS.32 distinct; 5: LDS.128, 4 distinct addrs/warp (8 lanes each)
template <int MODE>
__global__ void k_lds(float* out, int iters, long long* cyc) {
    extern __shared__ float4 sm4[];
    float* sm = (float*)sm4;
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 0.25f;
    __syncthreads();
    int lane = threadIdx.x & 31;
    int base;
    if (MODE == 0 || MODE == 2 || MODE == 3) base = 0;
    else if (MODE == 1) base = lane * 4;
    else if (MODE == 4) base = lane;
    else base = (lane >> 3) * 68;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int off = (base + u * 128 + (it & 7) * 4) & 8191;
            if (MODE == 0 || MODE == 1 || MODE == 5) {
                float4 v = *reinterpret_cast<const float4*>(sm + off);
                s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
            } else if (MODE == 3) {
                float2 v = *reinterpret_cast<const float2*>(sm + off);
                s0 += v.x; s1 += v.y;
            } else {
                s0 += sm[off];
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- F: dependent 64x64 layer chain, several mappings
constexpr int H = 64;
constexpr int WS = 68;   // padded row stride (floats) for weights and activations

__device__ __forceinline__ float elu(float v) { return v > 0.f ? v : expm1f(v); }

// lanes = trajectories (TB <= 32), NW warps each owning TN = 64/NW neurons; weights via broadcast LDS.128
template <int NW, bool PACKED>
__global__ void __launch_bounds__(NW * 32) k_chain_lanes_traj(const float* __restrict__ Wg, float* out, int nlayers, int TB, long long* cyc) {
    constexpr int TN = H / NW;
    extern __shared__ float4 sm4[];
    float* W = (float*)sm4;              // [64][WS]
    float* act = W + H * WS;             // [2][32][WS]
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) W[(i / H) * WS + (i % H)] = Wg[i];
    for (int i = threadIdx.x; i < 2 * 32 * WS; i += blockDim.x) act[i] = 0.01f * (i % 97) - 0.4f;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    bool active = lane < TB;
    long long t0 = clock64();
    for (int l = 0; l < nlayers; l++) {
        const float* ain = act + (l & 1) * 32 * WS + lane * WS;
        float* aout = act + ((l + 1) & 1) * 32 * WS + lane * WS;
        if (active) {
            if (!PACKED) {
                float acc[TN];
#pragma unroll
                for (int n = 0; n < TN; n++) acc[n] = 0.1f;
#pragma unroll 4
                for (int k = 0; k < H; k += 4) {
                    float4 a = *reinterpret_cast<const float4*>(ain + k);
#pragma unroll
                    for (int n = 0; n < TN; n++) {
                        float4 wv = *reinterpret_cast<const float4*>(W + (w * TN + n) * WS + k);
                        acc[n] = fmaf(a.x, wv.x, acc[n]);
                        acc[n] = fmaf(a.y, wv.y, acc[n]);
                        acc[n] = fmaf(a.z, wv.z, acc[n]);
                        acc[n] = fmaf(a.w, wv.w, acc[n]);
                    }
                }
#pragma unroll
                for (int n = 0; n < TN; n++) aout[w * TN + n] = elu(acc[n]);
            } else {
                u64 acc[TN];
#pragma unroll
                for (int n = 0; n < TN; n++) acc[n] = pack(0.1f, 0.f);
#pragma unroll 4
                for (int k = 0; k < H; k += 4) {
                    ulonglong2 a = *reinterpret_cast<const ulonglong2*>(ain + k);
#pragma unroll
                    for (int n = 0; n < TN; n++) {
                        ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(W + (w * TN + n) * WS + k);
                        acc[n] = fma2(a.x, wv.x, acc[n]);
                        acc[n] = fma2(a.y, wv.y, acc[n]);
                    }
                }
#pragma unroll
                for (int n = 0; n < TN; n++) { float2 f = unpack(acc[n]); aout[w * TN + n] = elu(f.x + f.y); }
            }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (active) out[(blockIdx.x * NW + w) * 32 + lane] = act[(nlayers & 1) * 32 * WS + lane * WS + w * TN];
}

// lanes = neurons: thread = (traj group tg of TM trajectories, neuron set of TN neurons {nidx + 32*q... interleaved});
// activations broadcast (warp shares tg), weights distinct per lane.
template <int TM, int TN, int NTG, bool PACKED>
__global__ void __launch_bounds__(NTG * (H / TN)) k_chain_lanes_neuron(const float* __restrict__ Wg, float* out, int nlayers, long long* cyc) {
    constexpr int NPG = H / TN;          // threads per trajectory group
    constexpr int TB = TM * NTG;
    extern __shared__ float4 sm4[];
    float* W = (float*)sm4;              // [64][WS]
    float* act = W + H * WS;             // [2][TB][WS]
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) W[(i / H) * WS + (i % H)] = Wg[i];
    for (int i = threadIdx.x; i < 2 * TB * WS; i += blockDim.x) act[i] = 0.01f * (i % 97) - 0.4f;
    __syncthreads();
    int tg = threadIdx.x / NPG, nidx = threadIdx.x % NPG;
    long long t0 = clock64();
    for (int l = 0; l < nlayers; l++) {
        const float* ain = act + (l & 1) * TB * WS + tg * TM * WS;
        float* aout = act + ((l + 1) & 1) * TB * WS + tg * TM * WS;
        if (!PACKED) {
            float acc[TM][TN];
#pragma unroll
            for (int m = 0; m < TM; m++)
#pragma unroll
                for (int n = 0; n < TN; n++) acc[m][n] = 0.1f;
#pragma unroll 2
            for (int k = 0; k < H; k += 4) {
                float4 wv[TN];
#pragma unroll
                for (int n = 0; n < TN; n++) wv[n] = *reinterpret_cast<const float4*>(W + (nidx + n * NPG) * WS + k);
#pragma unroll
                for (int m = 0; m < TM; m++) {
                    float4 a = *reinterpret_cast<const float4*>(ain + m * WS + k);
#pragma unroll
                    for (int n = 0; n < TN; n++) {
                        acc[m][n] = fmaf(a.x, wv[n].x, acc[m][n]);
                        acc[m][n] = fmaf(a.y, wv[n].y, acc[m][n]);
                        acc[m][n] = fmaf(a.z, wv[n].z, acc[m][n]);
                        acc[m][n] = fmaf(a.w, wv[n].w, acc[m][n]);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < TM; m++)
#pragma unroll
                for (int n = 0; n < TN; n++) aout[m * WS + nidx + n * NPG] = elu(acc[m][n]);
        } else {
            u64 acc[TM][TN];
#pragma unroll
            for (int m = 0; m < TM; m++)
#pragma unroll
                for (int n = 0; n < TN; n++) acc[m][n] = pack(0.1f, 0.f);
#pragma unroll 2
            for (int k = 0; k < H; k += 4) {
                ulonglong2 wv[TN];
#pragma unroll
                for (int n = 0; n < TN; n++) wv[n] = *reinterpret_cast<const ulonglong2*>(W + (nidx + n * NPG) * WS + k);
#pragma unroll
                for (int m = 0; m < TM; m++) {
                    ulonglong2 a = *reinterpret_cast<const ulonglong2*>(ain + m * WS + k);
#pragma unroll
                    for (int n = 0; n < TN; n++) {
                        acc[m][n] = fma2(a.x, wv[n].x, acc[m][n]);
                        acc[m][n] = fma2(a.y, wv[n].y, acc[m][n]);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < TM; m++)
#pragma unroll
                for (int n = 0; n < TN; n++) { float2 f = unpack(acc[m][n]); aout[m * WS + nidx + n * NPG] = elu(f.x + f.y); }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = act[(nlayers & 1) * TB * WS + tg * TM * WS + nidx];
}

// weights held in registers: thread = neuron j holds W[j][0..63]; loops over TB trajectories; activations broadcast LDS.128
template <int TBT>
__global__ void __launch_bounds__(64) k_chain_wreg(const float* __restrict__ Wg, float* out, int nlayers, long long* cyc) {
    extern __shared__ float4 sm4[];
    float* act = (float*)sm4;             // [2][TBT][WS]
    for (int i = threadIdx.x; i < 2 * TBT * WS; i += blockDim.x) act[i] = 0.01f * (i % 97) - 0.4f;
    float wr[H];
#pragma unroll
    for (int k = 0; k < H; k++) wr[k] = Wg[threadIdx.x * H + k];
    __syncthreads();
    long long t0 = clock64();
    for (int l = 0; l < nlayers; l++) {
        const float* ain = act + (l & 1) * TBT * WS;
        float* aout = act + ((l + 1) & 1) * TBT * WS;
#pragma unroll 2
        for (int m = 0; m < TBT; m++) {
            float a0 = 0.1f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < H; k += 4) {
                float4 a = *reinterpret_cast<const float4*>(ain + m * WS + k);
                a0 = fmaf(a.x, wr[k], a0); a1 = fmaf(a.y, wr[k + 1], a1);
                a2 = fmaf(a.z, wr[k + 2], a2); a3 = fmaf(a.w, wr[k + 3], a3);
            }
            aout[m * WS + threadIdx.x] = elu((a0 + a1) + (a2 + a3));
        }
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = act[(nlayers & 1) * TBT * WS + threadIdx.x];
}

// ---------------------------------------------------------------- expm1f cost
__global__ void k_expm1(float* out, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = -seed * (i + 1) - threadIdx.x * 1e-4f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = expm1f(v[i]) * 0.5f - 0.1f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float* d_out; static long long* d_cyc; static float* d_W;
static int nsm;

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}
double avg_cyc(int n) {
    std::vector<long long> h(n); CK(cudaMemcpy(h.data(), d_cyc, n * sizeof(long long), cudaMemcpyDeviceToHost));
    double s = 0; for (auto v : h) s += v; return s / n;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); nsm = p.multiProcessorCount;
    int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    printf("device %s SMs %d maxclk %d kHz smem/blk optin %zu\n", p.name, nsm, clk_khz, p.sharedMemPerBlockOptin);
    CK(cudaMalloc(&d_out, 1 << 24)); CK(cudaMalloc(&d_cyc, 8192 * 8)); CK(cudaMalloc(&d_W, H * H * 4));
    std::vector<float> hW(H * H); for (int i = 0; i < H * H; i++) hW[i] = ((i * 2654435761u) % 1000) / 1000.f * 0.25f - 0.125f;
    CK(cudaMemcpy(d_W, hW.data(), H * H * 4, cudaMemcpyHostToDevice));

    // A/B: FMA throughput vs warps per SM
    int iters = 20000;
    for (int bs : {128, 256, 512}) {
        float ms = time_ms([&] { k_ffma<<<nsm, bs>>>(d_out, iters, 1.0001f, 0.5f); });
        double fma = (double)nsm * bs * 16.0 * iters;
        printf("ffma_scalar(acc*x+y)   bs=%4d: %.3f ms  %.1f FMA/clk/SM @maxclk  %.2f TFLOP/s\n", bs, ms, fma / (ms * 1e-3) / nsm / (clk_khz * 1e3), 2 * fma / ms / 1e9);
        ms = time_ms([&] { k_ffma_tile<<<nsm, bs>>>(d_out, iters, 1.0001f, 0.5f); });
        printf("ffma_tile(a[i]*b[j]+c) bs=%4d: %.3f ms  %.1f FMA/clk/SM @maxclk  %.2f TFLOP/s\n", bs, ms, fma / (ms * 1e-3) / nsm / (clk_khz * 1e3), 2 * fma / ms / 1e9);
        ms = time_ms([&] { k_ffma2<<<nsm, bs>>>(d_out, iters, 1.0001f, 0.5f); });
        printf("ffma2(acc*x+y)         bs=%4d: %.3f ms  %.1f FMA/clk/SM @maxclk  %.2f TFLOP/s\n", bs, ms, 2 * fma / (ms * 1e-3) / nsm / (clk_khz * 1e3), 4 * fma / ms / 1e9);
        ms = time_ms([&] { k_ffma2_tile<<<nsm, bs>>>(d_out, iters, 1.0001f, 0.5f); });
        printf("ffma2_tile             bs=%4d: %.3f ms  %.1f FMA/clk/SM @maxclk  %.2f TFLOP/s\n", bs, ms, 2 * fma / (ms * 1e-3) / nsm / (clk_khz * 1e3), 4 * fma / ms / 1e9);
    }
    {
        float ms = time_ms([&] { k_expm1<<<nsm, 256>>>(d_out, 2000, 0.3f); });
        double n = (double)nsm * 256 * 8 * 2000;
        printf("expm1f: %.3f ms -> %.2f clk per warp-expm1f per SMSP (issue-equivalent)\n", ms, (ms * 1e-3) * (clk_khz * 1e3) / (n / 32 / nsm / 4));
    }
    // LDS
    auto run_lds = [&](auto kern, const char* name) {
        for (int bs : {32, 128, 256, 512}) {
            int it = 2000;
            kern<<<nsm, bs, 32768>>>(d_out, it, d_cyc); CK(cudaDeviceSynchronize());
            kern<<<nsm, bs, 32768>>>(d_out, it, d_cyc); CK(cudaDeviceSynchronize());
            double c = avg_cyc(nsm);
            printf("%s bs=%3d: %.2f cyc per warp-LDS (SM-wide)\n", name, bs, c / (double(it) * 16 * (bs / 32)));
        }
    };
    run_lds(k_lds<0>, "LDS.128 broadcast      ");
    run_lds(k_lds<1>, "LDS.128 distinct       ");
    run_lds(k_lds<5>, "LDS.128 4-addr(8 lanes)");
    run_lds(k_lds<3>, "LDS.64  broadcast      ");
    run_lds(k_lds<2>, "LDS.32  broadcast      ");
    run_lds(k_lds<4>, "LDS.32  distinct       ");

    // F: layer chains
    int NL = 2000;
    auto report = [&](const char* name, int grid, int trajPerCta, int ctasPerSm, float ms) {
        double c = avg_cyc(grid);
        double macs = (double)grid * trajPerCta * NL * 4096.0;
        printf("%-44s grid=%4d traj/CTA=%2d: %8.1f cyc/layer  %.3f ms  %.1f MAC/clk/SM  (%.1f%% of 128)\n", name, grid, trajPerCta, c / NL, ms,
               macs / (ms * 1e-3) / nsm / (clk_khz * 1e3), 100.0 * macs / (ms * 1e-3) / nsm / (clk_khz * 1e3) / 128.0);
    };
#define SMEM_LT ((H * WS + 2 * 32 * WS) * 4)
#define RUN_LT(NW, PK, TB, CPS) { \
        CK(cudaFuncSetAttribute(k_chain_lanes_traj<NW, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LT)); \
        int grid = nsm * CPS; \
        float ms = time_ms([&] { k_chain_lanes_traj<NW, PK><<<grid, NW * 32, SMEM_LT>>>(d_W, d_out, NL, TB, d_cyc); }); \
        report("lanes=traj NW=" #NW " packed=" #PK " CTAs/SM=" #CPS, grid, TB, CPS, ms); }
    RUN_LT(4, false, 28, 1) RUN_LT(4, true, 28, 1) RUN_LT(8, false, 28, 1) RUN_LT(8, true, 28, 1)
    RUN_LT(16, false, 28, 1) RUN_LT(16, true, 28, 1)
    RUN_LT(4, false, 32, 2) RUN_LT(4, true, 32, 2) RUN_LT(8, true, 32, 2)
    RUN_LT(4, true, 32, 4)
#define RUN_LN(TM, TN, NTG, PK, CPS) { \
        int smem = (H * WS + 2 * TM * NTG * WS) * 4; \
        CK(cudaFuncSetAttribute(k_chain_lanes_neuron<TM, TN, NTG, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        int grid = nsm * CPS; \
        float ms = time_ms([&] { k_chain_lanes_neuron<TM, TN, NTG, PK><<<grid, NTG * (H / TN), smem>>>(d_W, d_out, NL, d_cyc); }); \
        report("lanes=neuron TM=" #TM " TN=" #TN " NTG=" #NTG " packed=" #PK " CTAs/SM=" #CPS, grid, TM * NTG, CPS, ms); }
    RUN_LN(7, 1, 4, false, 1) RUN_LN(7, 1, 4, true, 1)
    RUN_LN(7, 2, 4, false, 1) RUN_LN(7, 2, 4, true, 1)
    RUN_LN(7, 2, 2, false, 2) RUN_LN(7, 2, 2, true, 2)
    RUN_LN(7, 1, 2, false, 2) RUN_LN(7, 1, 2, true, 2)
    RUN_LN(4, 2, 8, false, 1) RUN_LN(4, 2, 8, true, 1)
    RUN_LN(4, 4, 8, false, 1) RUN_LN(4, 4, 8, true, 1)
    RUN_LN(8, 2, 4, false, 1) RUN_LN(8, 2, 4, true, 1)
    RUN_LN(8, 4, 4, false, 1) RUN_LN(8, 4, 4, true, 1)
    RUN_LN(7, 4, 4, false, 1) RUN_LN(7, 4, 4, true, 1)
    RUN_LN(7, 4, 2, true, 2) RUN_LN(7, 2, 1, true, 4) RUN_LN(7, 1, 1, true, 4)
#define RUN_WR(TBT, CPS) { \
        int smem = (2 * TBT * WS) * 4; \
        int grid = nsm * CPS; \
        float ms = time_ms([&] { k_chain_wreg<TBT><<<grid, 64, smem>>>(d_W, d_out, NL, d_cyc); }); \
        report("weights-in-regs TB=" #TBT " CTAs/SM=" #CPS, grid, TBT, CPS, ms); }
    RUN_WR(7, 4) RUN_WR(4, 7) RUN_WR(14, 2) RUN_WR(2, 14) RUN_WR(28, 1)
    printf("done\n");
    return 0;
}
