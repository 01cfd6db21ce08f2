// psnode_wide_grad.cu -- weight gradients of the latent `*_02` nets from the operand blocks the two sweeps recorded, and the
// host side of the whole wide reverse pass (psnode_wide.cuh).
//
// Autograd in the reference accumulates dW += delta . act^T once per Linear call of the unrolled loop
// (loss.backward(), neural_00_ODE_02_direct_encode.py:274); here the three products
//     dW2 = sum_{step, stage} delta2 . a1^T      dF_x = sum_{step, stage} delta1 . y^T      dF_z = sum_{step} (sum_e delta1) . z^T
// are ONE pass over the tapes with full-rate tcgen05 MMAs: every 128 x 16 block is already a canonical K-major UMMA tile
// (K = the 16 trajectories of a group), so a TMA bulk copy brings it into shared memory as the operand it is; threads only
// split it into tf32 hi / lo (3xTF32).  M = N = 128, K = 16 per record.  The tensor core's fp32 accumulation truncates, so an
// accumulator is drained into registers (round-to-nearest adds) every 2 records (12 accumulations) while the MMAs of the
// next pair run into the other accumulator.  Each CTA owns one product and a contiguous range of records; per-CTA slabs are
// summed in a fixed order (deterministic, no floating-point atomics).
// Bound: HBM (16 KB of operands per 6 MMAs = 384 tensor-pipe cycles per SM).
#include <cstddef>
#include "psnode_wide.cuh"

namespace {
using namespace psn_tc;

constexpr int H = PSW_H;
constexpr int NSLOT = 3;                           // ring slots; a slot holds a PAIR of records
constexpr int BLK_BYTES = PSW_BLOCK * 4;           // 8 KB
constexpr int SPLIT_WARPS = 16;                    // splitter / drain warps (one float4 of a block per thread)
constexpr int GRAD_THREADS = (SPLIT_WARPS + 2) * 32;   // + TMA producer warp + MMA issuer warp

// One slot = two records: per-iteration fixed costs (mbarrier waits, proxy fence, MMA issue, refill) are ~1400 cycles whatever the
// amount of work, which held the one-record-per-iteration version at 3.3 TB/s (50 % of the copy bandwidth) with the tensor pipe
// 29 % busy; a pair halves that overhead per byte.  The two records of a pair are one accumulation chain (12 MMAs).
struct __align__(128) Slot {
    float a_hi[2][PSW_BLOCK], a_lo[2][PSW_BLOCK], b_hi[2][PSW_BLOCK], b_lo[2][PSW_BLOCK];
};
struct __align__(128) GradSmem {
    Slot slot[NSLOT];
    uint64_t full[NSLOT];          // producer -> splitters (transaction count)
    uint64_t split[NSLOT];         // splitters -> issuer   (16 warp arrivals)
    uint64_t done[NSLOT];          // issuer (tcgen05.commit) -> producer (slot free) and drainers (accumulator complete)
    uint64_t drained[2];           // drainers -> issuer: accumulator buffer may be overwritten
    uint32_t tmem_base;
};

struct GradParams {
    // product r: A blocks at a_base[r] + rec * a_stride[r], B blocks at b_base[r] + rec * b_stride[r], nrec[r] records,
    // CTAs [cta0[r], cta0[r+1]) work on it
    const float* a_base[3]; const float* b_base[3];
    int64_t a_stride[3], b_stride[3];
    int64_t nrec[3];
    int cta0[4];
    float* slabs;                                  // [gridDim.x][128][128]
    int* err;
};

// Warp roles (as psn_lg_gemm_kernel): warps 0..15 split the operand blocks and drain the accumulators, warp 16 is the TMA producer,
// warp 17 the MMA issuer; every role waits only on the mbarrier of the role before it, so loads, splits, MMAs and drains of
// different pairs overlap up to the depth of the ring.  (Version 1: one CTA barrier per pair and thread 0 refilling inside the loop,
// 3.7 TB/s = 57 % of the copy bandwidth.)
__global__ void __launch_bounds__(GRAD_THREADS, 1) psn_wide_grad_kernel(const __grid_constant__ GradParams q) {
    extern __shared__ unsigned char smem_raw[];
    GradSmem& sm = *reinterpret_cast<GradSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    const int tid = threadIdx.x, lane = tid & 31;
    const int cw = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int wq = cw & 3, cc = (cw >> 2) & 3;
    const int role = blockIdx.x >= q.cta0[2] ? 2 : (blockIdx.x >= q.cta0[1] ? 1 : 0);
    const int ncta = q.cta0[role + 1] - q.cta0[role], me = blockIdx.x - q.cta0[role];
    const int64_t per = (q.nrec[role] + ncta - 1) / ncta;
    const int64_t r0 = per * me, r1 = r0 + per < q.nrec[role] ? r0 + per : q.nrec[role];
    const int n = r1 > r0 ? (int)(r1 - r0) : 0;            // records of this CTA
    const int np = (n + 1) / 2;                            // pairs (the last one may hold a single record)
    const float* abase = q.a_base[role];
    const float* bbase = q.b_base[role];
    const int64_t astr = q.a_stride[role], bstr = q.b_stride[role];

    if (tid == 0) {
        for (int s = 0; s < NSLOT; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.split[s], SPLIT_WARPS); mbar_init(&sm.done[s], 1); }
        mbar_init(&sm.drained[0], SPLIT_WARPS); mbar_init(&sm.drained[1], SPLIT_WARPS);
        fence_mbar_init();
    }
    if (cw == 0) tmem_alloc(&sm.tmem_base, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;
    const uint32_t lane_base = (uint32_t)(32 * wq) << 16;

    if (cw == SPLIT_WARPS) {
        // ---- TMA producer ----
        if (elect_one()) {
            for (int i = 0; i < np; i++) {
                const int s = i % NSLOT;
                if (i >= NSLOT && !mbar_wait(&sm.done[s], (uint32_t)(((i - NSLOT) / NSLOT) & 1))) { atomicExch(q.err, 7); __trap(); }
                const int cnt = (2 * i + 1 < n) ? 2 : 1;
                mbar_expect_tx(&sm.full[s], (uint32_t)(2 * cnt * BLK_BYTES));
                for (int k = 0; k < cnt; k++) {
                    bulk_g2s(sm.slot[s].a_hi[k], abase + (r0 + 2 * i + k) * astr, BLK_BYTES, &sm.full[s]);
                    bulk_g2s(sm.slot[s].b_hi[k], bbase + (r0 + 2 * i + k) * bstr, BLK_BYTES, &sm.full[s]);
                }
            }
        }
        __syncwarp();
    } else if (cw == SPLIT_WARPS + 1) {
        // ---- MMA issuer ----
        const uint32_t idesc = make_idesc_tf32(H, H);
        for (int i = 0; i < np; i++) {
            const int s = i % NSLOT;
            Slot& sl = sm.slot[s];
            const int cnt = (2 * i + 1 < n) ? 2 : 1;
            if (i >= 2 && !mbar_wait(&sm.drained[i & 1], (uint32_t)(((i - 2) >> 1) & 1))) { atomicExch(q.err, 10); __trap(); }
            if (!mbar_wait(&sm.split[s], (uint32_t)((i / NSLOT) & 1))) { atomicExch(q.err, 11); __trap(); }
            if (elect_one()) {
                tc_fence_after();
                const uint32_t d = tmem + (uint32_t)((i & 1) * H);
                uint32_t accumulate = 0;
                for (int k = 0; k < cnt; k++) {
                    // K-major no-swizzle tiles, rows = 128, K = 16: LBO = 128 B, SBO = 512 B; a K = 8 step advances 256 B
                    const uint64_t da_hi = make_desc(smem_u32(sl.a_hi[k]), 128, 512), da_lo = make_desc(smem_u32(sl.a_lo[k]), 128, 512);
                    const uint64_t db_hi = make_desc(smem_u32(sl.b_hi[k]), 128, 512), db_lo = make_desc(smem_u32(sl.b_lo[k]), 128, 512);
#pragma unroll
                    for (int term = 0; term < 3; term++) {
                        const uint64_t ad = term == 0 ? da_lo : da_hi;
                        const uint64_t bd = term == 1 ? db_lo : db_hi;
#pragma unroll
                        for (int ks = 0; ks < 2; ks++) {
                            mma_tf32(d, ad + (uint64_t)(16 * ks), bd + (uint64_t)(16 * ks), idesc, accumulate);
                            accumulate = 1;
                        }
                    }
                }
                mma_commit(&sm.done[s]);
            }
            __syncwarp();
        }
    } else {
        // ---- splitters / drainers ----
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; i++) acc[i] = 0.0f;
        // drain the accumulator of pair `pi` (this thread: lane m = 32 wq + lane, columns 32 cc .. 32 cc + 31) into the register sums
        auto drain = [&](int pi) {
            if (!mbar_wait(&sm.done[pi % NSLOT], (uint32_t)((pi / NSLOT) & 1))) { atomicExch(q.err, 8); __trap(); }
            tc_fence_after();
            const int buf = pi & 1;
            float v0[16], v1[16];
            tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)(buf * H + 32 * cc), v0);
            tmem_ld_32x32b_x16(tmem + lane_base + (uint32_t)(buf * H + 32 * cc + 16), v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i++) { acc[i] += v0[i]; acc[16 + i] += v1[i]; }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.drained[buf]);
        };
        for (int i = 0; i < np; i++) {
            const int s = i % NSLOT;
            Slot& sl = sm.slot[s];
            const int cnt = (2 * i + 1 < n) ? 2 : 1;
            if (!mbar_wait(&sm.full[s], (uint32_t)((i / NSLOT) & 1))) { atomicExch(q.err, 6); __trap(); }
            for (int k = 0; k < cnt; k++) {          // tf32 hi (in place) / lo split: 512 float4 per block
                float4* ah = reinterpret_cast<float4*>(sl.a_hi[k]); float4* al = reinterpret_cast<float4*>(sl.a_lo[k]);
                float4* bh = reinterpret_cast<float4*>(sl.b_hi[k]); float4* bl = reinterpret_cast<float4*>(sl.b_lo[k]);
                float4 lo;
                float4 hi = split4_hi(ah[tid], lo);
                ah[tid] = hi; al[tid] = lo;
                hi = split4_hi(bh[tid], lo);
                bh[tid] = hi; bl[tid] = lo;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.split[s]);
            if (i >= 1) drain(i - 1);                // the previous pair's accumulator, while this pair's MMAs run
        }
        if (np > 0) drain(np - 1);
        float* slab = q.slabs + (int64_t)blockIdx.x * H * H + (int64_t)(32 * wq + lane) * H + 32 * cc;
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(slab + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
    }
    tc_fence_before();
    __syncthreads();
    if (cw == 0) tmem_dealloc(tmem, 256);
}

// dc[b][m] = sum_rows dpre[row][b][m]   (gradient of the per-trajectory layer-1 constant c)
__global__ void psn_wide_colsum_kernel(const float* __restrict__ dpre, int64_t sr, int rows, int64_t ncol, float* __restrict__ dc) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncol) return;
    float s = 0.0f;
    for (int r = 0; r < rows; r++) s += __ldcs(dpre + (int64_t)r * sr + c);
    dc[c] = s;
}
// dca[m][k] = sum_b dc[b][m] a0[b][k]   (k < 256): block = k, thread = m
__global__ void __launch_bounds__(128) psn_wide_dca_kernel(const float* __restrict__ dc, const float* __restrict__ a0, int64_t a0_sb, int B,
                                                             float* __restrict__ dca) {
    const int k = blockIdx.x, m = threadIdx.x;
    float s0 = 0.0f, s1 = 0.0f;
    int b = 0;
    for (; b + 1 < B; b += 2) {
        s0 = fmaf(__ldg(dc + (int64_t)b * H + m), __ldg(a0 + (int64_t)b * a0_sb + k), s0);
        s1 = fmaf(__ldg(dc + (int64_t)(b + 1) * H + m), __ldg(a0 + (int64_t)(b + 1) * a0_sb + k), s1);
    }
    if (b < B) s0 = fmaf(__ldg(dc + (int64_t)b * H + m), __ldg(a0 + (int64_t)b * a0_sb + k), s0);
    dca[m * 2 * H + k] = s0 + s1;
}
// d_a0[b][k] = sum_m dc[b][m] (W_a - W_b)[m][k]: block = 8 trajectories, thread = k (256)
__global__ void __launch_bounds__(256) psn_wide_da0_kernel(const float* __restrict__ dc, const float* __restrict__ W1, int B, float* __restrict__ d_a0,
                                                             int64_t d_a0_sb) {
    __shared__ float d[8][H];
    const int k = threadIdx.x, b0 = blockIdx.x * 8;
    for (int e = k; e < 8 * H; e += 256) d[e / H][e % H] = (b0 + e / H) < B ? dc[(int64_t)(b0 + e / H) * H + e % H] : 0.0f;
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; n++) acc[n] = 0.0f;
    for (int m = 0; m < H; m++) {
        const float w = __ldg(W1 + (int64_t)m * 6 * H + k) - __ldg(W1 + (int64_t)m * 6 * H + 2 * H + k);
#pragma unroll
        for (int n = 0; n < 8; n++) acc[n] = fmaf(w, d[n][m], acc[n]);
    }
    for (int n = 0; n < 8; n++)
        if (b0 + n < B) d_a0[(int64_t)(b0 + n) * d_a0_sb + k] = acc[n];
}
// d_theta = [dW1 (128 x 768) | db1 | dW2 (128 x 128) | db2], slabs summed in CTA order
__global__ void psn_wide_assemble_kernel(const float* __restrict__ slabs, int c0, int c1, int c2, int c3, const float* __restrict__ dca,
                                         const float* __restrict__ dc, int B, const float* __restrict__ db2_slab, int ngroups,
                                         float* __restrict__ d_theta) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nW1 = H * 6 * H, nW2 = H * H;
    auto slab_sum = [&](int from, int to, int m, int k) {
        float s = 0.0f;
        for (int c = from; c < to; c++) s += slabs[(int64_t)c * H * H + m * H + k];
        return s;
    };
    if (idx < nW1) {
        const int m = idx / (6 * H), col = idx - m * 6 * H, blk = col / (2 * H), k = col - blk * 2 * H;
        const float a = dca[m * 2 * H + k];
        if (blk == 0) { d_theta[idx] = a; return; }
        const float gsum = k < H ? slab_sum(c1, c2, m, k) : slab_sum(c2, c3, m, k - H);
        d_theta[idx] = blk == 1 ? gsum - a : gsum;
    } else if (idx < nW1 + H) {
        const int m = idx - nW1;
        float s = 0.0f;
        for (int b = 0; b < B; b++) s += dc[(int64_t)b * H + m];
        d_theta[idx] = s;
    } else if (idx < nW1 + H + nW2) {
        const int e = idx - nW1 - H;
        d_theta[idx] = slab_sum(c0, c1, e / H, e % H);
    } else if (idx < nW1 + H + nW2 + H) {
        const int m = idx - nW1 - H - nW2;
        float s = 0.0f;
        for (int gq = 0; gq < ngroups; gq++) s += db2_slab[(int64_t)gq * H + m];
        d_theta[idx] = s;
    }
}
__global__ void psn_wide_zero_row_kernel(float* p, int64_t sb, int B) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < B * H) p[(int64_t)(idx / H) * sb + (idx % H)] = 0.0f;
}

int64_t align64(int64_t floats) { return (floats + 63) & ~(int64_t)63; }

struct BwdLayout {
    int64_t err, btape, stape, dpre, db2, dc, dca, slabs, total;
    int nslab;
};
BwdLayout bwd_layout(const psnode_problem* p) {
    BwdLayout L;
    const int E = p->event_idx ? p->E : 0;
    const int64_t steps = p->T > 1 ? p->T - 1 : 0;
    const int64_t ng = psw_ngroups(p->B);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    L.nslab = sms < 3 ? 3 : sms;
    int64_t o = 0;
    L.err = o; o += 64;
    L.btape = o; o += align64(ng * steps * psw_nstages(p->method) * PSW_BWD_REC);
    L.stape = o; o += align64(ng * steps * PSW_STEP_REC);
    L.dpre = o; o += align64(psw_pre_floats(p->B, p->T, E) + PSW_BLOCK);
    L.db2 = o; o += align64(ng * PSW_H);
    L.dc = o; o += align64(psw_bpad(p->B) * PSW_H);
    L.dca = o; o += align64(PSW_H * 2 * PSW_H);
    L.slabs = o; o += align64((int64_t)L.nslab * PSW_H * PSW_H);
    L.total = o;
    return L;
}

}  // namespace

int psn_wide_bwd_sweep(const psnode_problem* p, const psnode_adjoint* a, const float* tape, float* btape, float* stape, float* dpre,
                       float* db2_slab, int* err, cudaStream_t stream);

bool psn_wide_bwd_supports(const psnode_problem* p, const psnode_adjoint* a) {
    if (!psn_wide_supports(p) || !p->tape || p->tape_floats < psw_tape_floats(p->B, p->T, p->method)) return false;
    if (a->d_xteach.p || a->d_iteach.p) return false;
    if (!a->gx.p && !a->fuse_x.target.p) return false;
    return true;
}

int64_t psn_wide_backward_workspace(const psnode_problem* p, const psnode_adjoint* a) {
    (void)a;
    return bwd_layout(p).total * 4;
}

int psn_wide_backward(const psnode_problem* p, const psnode_adjoint* a, void* ws, int64_t ws_bytes, cudaStream_t stream) {
    const BwdLayout L = bwd_layout(p);
    if (ws == nullptr || ws_bytes < L.total * 4) return PSNODE_EWORKSPACE;
    float* w = static_cast<float*>(ws);
    int* err = reinterpret_cast<int*>(w + L.err);
    const int B = p->B, T = p->T, E = p->event_idx ? p->E : 0;
    const int NST = psw_nstages(p->method);
    const int64_t bpad = psw_bpad(B), steps = T > 1 ? T - 1 : 0, ng = psw_ngroups(B);
    const float* W1 = p->de.W[0];
    PSN_CUDA(cudaMemsetAsync(err, 0, 256, stream));
    if (E > 0) PSN_CUDA(cudaMemsetAsync(w + L.dpre + steps * bpad * PSW_H, 0, (size_t)E * bpad * PSW_H * 4, stream));
    int st = psn_wide_bwd_sweep(p, a, p->tape, w + L.btape, w + L.stape, w + L.dpre, w + L.db2, err, stream);
    if (st != PSNODE_OK) return st;
    // ---- input-series / jump gradients: d_z[r] = F_z^T dpre[r] (TMA-staged row GEMM) ----
    if (a->d_z.p) {
        if (steps > 0) {
            PswProjJob job;
            job.in = w + L.dpre; job.in_sr = bpad * PSW_H; job.in_sb = PSW_H;
            job.R = (int)steps; job.B = B;
            job.W = W1 + 3 * PSW_H; job.W2 = W1 + 5 * PSW_H; job.ldw = 6 * PSW_H; job.transpose = 1;
            job.add = nullptr; job.add_sb = 0;
            job.out = a->d_z.p; job.out_sr = a->d_z.st; job.out_sb = a->d_z.sb;
            job.zero_rows_from = job.R;
            st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<d_z>");
            if (st != PSNODE_OK) return st;
        }
        psn_wide_zero_row_kernel<<<(B * PSW_H + 255) / 256, 256, 0, stream>>>(a->d_z.p + (int64_t)(T - 1) * a->d_z.st, a->d_z.sb, B);   // z[T-1] is never read
        psn_count_launch("psn_wide_zero_row_kernel");
        PSN_CUDA(cudaGetLastError());
    }
    if (a->d_zjump && E > 0) {
        PswProjJob job;
        job.in = w + L.dpre + steps * bpad * PSW_H; job.in_sr = bpad * PSW_H; job.in_sb = PSW_H;
        job.R = E; job.B = B;
        job.W = W1 + 3 * PSW_H; job.W2 = W1 + 5 * PSW_H; job.ldw = 6 * PSW_H; job.transpose = 1;
        job.add = nullptr; job.add_sb = 0;
        job.out = a->d_zjump; job.out_sr = a->d_zj_se; job.out_sb = a->d_zj_sb;
        job.zero_rows_from = job.R;
        st = psn_wide_proj(job, err, stream, "psn_wide_proj_kernel<d_zjump>");
        if (st != PSNODE_OK) return st;
    }
    // ---- weight gradients ----
    GradParams g;
    g.a_base[0] = w + L.btape;             g.b_base[0] = p->tape;               // dW2 = delta2 . a1^T
    g.a_base[1] = w + L.btape + PSW_BLOCK; g.b_base[1] = p->tape + PSW_BLOCK;   // dF_x = delta1 . y^T
    g.a_base[2] = w + L.stape;             g.b_base[2] = w + L.stape + PSW_BLOCK;   // dF_z = (sum_e delta1) . z^T
    g.a_stride[0] = g.a_stride[1] = PSW_BWD_REC; g.b_stride[0] = g.b_stride[1] = PSW_FWD_REC;
    g.a_stride[2] = g.b_stride[2] = PSW_STEP_REC;
    g.nrec[0] = g.nrec[1] = ng * steps * NST; g.nrec[2] = ng * steps;
    const int ncta = L.nslab;
    int c2n = ncta / (2 * NST + 1); if (c2n < 1) c2n = 1;
    const int c01 = (ncta - c2n) / 2;
    g.cta0[0] = 0; g.cta0[1] = c01; g.cta0[2] = 2 * c01; g.cta0[3] = ncta;
    g.slabs = w + L.slabs;
    g.err = err;
    {
        const int smem = (int)sizeof(GradSmem) + 128;
        PSN_CUDA(cudaFuncSetAttribute(psn_wide_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        psn_wide_grad_kernel<<<ncta, GRAD_THREADS, smem, stream>>>(g);
        psn_count_launch("psn_wide_grad_kernel");
        PSN_CUDA(cudaGetLastError());
    }
    const int rows = (int)psw_pre_rows(T, E);
    psn_wide_colsum_kernel<<<(int)((bpad * PSW_H + 255) / 256), 256, 0, stream>>>(w + L.dpre, bpad * PSW_H, rows, bpad * PSW_H, w + L.dc);
    psn_count_launch("psn_wide_colsum_kernel");
    psn_wide_dca_kernel<<<2 * PSW_H, 128, 0, stream>>>(w + L.dc, p->a0, p->a0_sb, B, w + L.dca);
    psn_count_launch("psn_wide_dca_kernel");
    if (a->d_a0) {
        psn_wide_da0_kernel<<<(B + 7) / 8, 256, 0, stream>>>(w + L.dc, W1, B, a->d_a0, a->d_a0_sb);
        psn_count_launch("psn_wide_da0_kernel");
    }
    const int ntheta = PSW_H * 6 * PSW_H + PSW_H + PSW_H * PSW_H + PSW_H;
    if (a->n_theta < ntheta) return PSNODE_EINVAL;
    psn_wide_assemble_kernel<<<(ntheta + 255) / 256, 256, 0, stream>>>(w + L.slabs, g.cta0[0], g.cta0[1], g.cta0[2], g.cta0[3], w + L.dca, w + L.dc, B,
                                                                       w + L.db2, (int)ng, a->d_theta);
    psn_count_launch("psn_wide_assemble_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}

// Two big-K products D_r = sum_rec A_r[rec] . B_r[rec]^T (128 x 128, 128 x 16 blocks) with the kernel above, half of the CTAs each: per-CTA slabs
// [0, cta0[1]) belong to product 0, [cta0[1], cta0[2]) to product 1 (psnode_wide4_bwd.cu: dW2 and dW3 of the 4-layer net)
int psn_wide_grad_pairs(const float* a0, int64_t a0_stride, const float* b0, int64_t b0_stride, const float* a1, int64_t a1_stride,
                        const float* b1, int64_t b1_stride, int64_t nrec, int ncta, float* slabs, int* err, int cta0[3], cudaStream_t stream) {
    GradParams g;
    g.a_base[0] = a0; g.b_base[0] = b0; g.a_stride[0] = a0_stride; g.b_stride[0] = b0_stride;
    g.a_base[1] = a1; g.b_base[1] = b1; g.a_stride[1] = a1_stride; g.b_stride[1] = b1_stride;
    g.a_base[2] = a1; g.b_base[2] = b1; g.a_stride[2] = a1_stride; g.b_stride[2] = b1_stride;
    g.nrec[0] = g.nrec[1] = nrec; g.nrec[2] = 0;
    if (ncta < 2) return PSNODE_EINVAL;
    g.cta0[0] = 0; g.cta0[1] = ncta / 2; g.cta0[2] = ncta; g.cta0[3] = ncta;
    g.slabs = slabs;
    g.err = err;
    cta0[0] = 0; cta0[1] = g.cta0[1]; cta0[2] = ncta;
    const int smem = (int)sizeof(GradSmem) + 128;
    PSN_CUDA(cudaFuncSetAttribute(psn_wide_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    psn_wide_grad_kernel<<<ncta, GRAD_THREADS, smem, stream>>>(g);
    psn_count_launch("psn_wide_grad_kernel");
    PSN_CUDA(cudaGetLastError());
    return PSNODE_OK;
}
