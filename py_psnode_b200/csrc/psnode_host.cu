// psnode_host.cu -- psnode_forward_host: the host-buffer entry point (what a maintainer binds when the batch lives in host
// memory, e.g. straight out of the reference's DataLoader): inputs reach the device, the problem is integrated, and the
// trajectory lands in the caller's host buffers before the call returns.
//
// Two transfer modes, chosen per buffer:
//   * pinned (cudaHostAlloc / cudaHostRegister'ed / torch pin_memory) buffers are mapped into the device address space, so
//     the integrator reads its inputs and writes its trajectory there directly ("zero copy"): the PCIe transfer is fused
//     into the kernel and overlaps it completely -- per step the kernel touches ~50 KB of inputs, prefetched one step
//     ahead, and writes 1 KB per 16 trajectories as full 128-byte lines;
//   * ordinary pageable buffers are staged through a cached device arena with cudaMemcpyAsync on `stream`.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "psnode_internal.cuh"

namespace {
struct Scratch {
    char* p = nullptr;
    size_t cap = 0;
    char* get(size_t n) {
        if (n > cap) {
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            if (cudaMalloc(&p, n) != cudaSuccess) return nullptr;
            cap = n;
        }
        return p;
    }
};
Scratch g_scratch;

// device alias of a device-visible pointer (pinned host, device or managed memory), nullptr for pageable host memory.
// Small pinned buffers (weights, all_initial, jump tables: read by every CTA in its prologue) are staged instead: one
// async copy is cheaper than hundreds of PCIe round trips.
constexpr size_t kZeroCopyMinBytes = size_t(1) << 20;
void* mapped_alias(const void* p, size_t bytes) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return at.devicePointer;
    if (at.type == cudaMemoryTypeHost && bytes >= kZeroCopyMinBytes) return at.devicePointer;
    return nullptr;
}

// contiguous element range [lo, hi) covered by a strided (T,B,width) view
struct Span { int64_t lo, hi; };
Span span_of(int64_t st, int64_t sb, int T, int B, int w) {
    int64_t lo = 0, hi = 0;
    const int64_t a = st * (T - 1), b = sb * (B - 1);
    if (a < 0) lo += a; else hi += a;
    if (b < 0) lo += b; else hi += b;
    hi += w;
    return {lo, hi};
}

struct Plan {
    struct Item { const char* src; size_t bytes; size_t off; char* alias; };
    std::vector<Item> items;
    size_t cur = 0;
    int64_t staged = 0, direct = 0;
    size_t reserve(size_t bytes) { size_t o = cur; cur += (bytes + 255) & ~size_t(255); return o; }
    // returns the item index; resolve() turns it into a device pointer once the arena exists
    int add(const void* src, size_t bytes) {
        Item it{static_cast<const char*>(src), bytes, 0, static_cast<char*>(mapped_alias(src, bytes))};
        if (it.alias) direct += (int64_t)bytes; else { it.off = reserve(bytes); staged += (int64_t)bytes; }
        items.push_back(it);
        return (int)items.size() - 1;
    }
    char* resolve(int idx, char* base) const { return items[idx].alias ? items[idx].alias : base + items[idx].off; }
};
}  // namespace

// ---- DMA path (PSNODE_HOST_PATH=dma) --------------------------------------------------------------------------------------------
// The zero-copy path makes every SM issue its own PCIe reads / 128-bit writes; with 8 GPUs behind one root complex that is
// what limits the end-to-end rate (SCALE r01: 7.6 ms at 1 GPU, 24.8 ms at 8).  Here the copy engines move the data instead:
// the grid is cut into time chunks, chunk c + 1's input rows are copied host -> device and chunk c - 1's trajectory rows
// device -> host on two side streams while chunk c integrates (the state at a chunk boundary is the last trajectory row;
// i_0 of a DAE chunk is re-evaluated from it exactly as my_solvers.py:95 does at t_0).  Dense time-major buffers only.
namespace {
struct DmaCtx {
    cudaStream_t s_in = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> in_ready, done;
    Scratch arena;
    bool init(int n) {
        if (!s_in && cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (!s_out && cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking) != cudaSuccess) return false;
        while ((int)in_ready.size() < n) {
            cudaEvent_t a, b;
            if (cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess) return false;
            if (cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) return false;
            in_ready.push_back(a); done.push_back(b);
        }
        return true;
    }
};
DmaCtx g_dma;
bool dense_tm(const psnode_series& sr, int B, int w) { return !sr.p || w == 0 || (sr.sb == w && sr.st == (int64_t)B * w); }
bool dense_tm_out(const psnode_series_out& sr, int B, int w) { return sr.p && sr.sb == w && sr.st == (int64_t)B * w; }
}  // namespace

static int forward_host_dma(const psnode_problem* hp, cudaStream_t s, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    psnode_problem p = *hp;
    const bool dae = p.kind == PSNODE_DAE;
    const int S = psn_S(&p), T = p.T, B = p.B;
    const int NCH = T >= 64 ? 8 : 1;
    if (!g_dma.init(NCH + 1)) return psn_cuda_fail(cudaErrorUnknown, "dma streams");
    // arena layout
    size_t cur = 0;
    auto reserve = [&](size_t bytes) { size_t o = cur; cur += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_t = reserve((size_t)T * B * 4), o_z = reserve((size_t)T * B * p.Z * 4), o_v = reserve((size_t)T * B * p.V * 4);
    const size_t o_x0 = reserve((size_t)B * p.X * 4), o_a0 = reserve((size_t)B * S * 4);
    const size_t o_ev = reserve((size_t)(T > 1 ? T - 1 : 1) * 4);
    const size_t o_zj = reserve(p.event_idx ? (size_t)B * p.E * p.Z * 4 : 0), o_vj = reserve(p.event_idx && dae ? (size_t)B * p.E * p.V * 4 : 0);
    size_t o_W[2][PSNODE_MAX_LAYERS], o_b[2][PSNODE_MAX_LAYERS];
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        const psnode_mlp& m = net ? p.ae : p.de;
        for (int l = 0; l < m.n_layers; l++) { o_W[net][l] = reserve((size_t)m.out_dim[l] * m.in_dim[l] * 4); o_b[net][l] = reserve((size_t)m.out_dim[l] * 4); }
    }
    const size_t o_xs = reserve((size_t)T * B * p.X * 4), o_is = reserve(dae ? (size_t)T * B * p.I * 4 : 0);
    const int64_t ws_bytes = psnode_forward_workspace(&p);
    const size_t o_ws = reserve((size_t)(ws_bytes > 0 ? ws_bytes : 256));
    char* base = g_dma.arena.get(cur);
    if (!base) return psn_cuda_fail(cudaErrorMemoryAllocation, "psnode_forward_host dma arena");
    int64_t up = 0, down = 0;
    auto h2d = [&](size_t off, const void* src, size_t bytes, cudaStream_t st) -> int {
        if (bytes == 0) return PSNODE_OK;
        PSN_CUDA(cudaMemcpyAsync(base + off, src, bytes, cudaMemcpyHostToDevice, st));
        up += (int64_t)bytes;
        return PSNODE_OK;
    };
    // per-call constants on the compute stream
    int r = h2d(o_a0, hp->a0, (size_t)B * S * 4, s);
    if (r == PSNODE_OK) r = h2d(o_x0, dae ? hp->x_init : hp->x.p, (size_t)B * p.X * 4, s);
    if (r == PSNODE_OK && p.event_idx) {
        r = h2d(o_ev, hp->event_idx, (size_t)(T > 1 ? T - 1 : 1) * 4, s);
        if (r == PSNODE_OK && p.Z) r = h2d(o_zj, hp->z_jump, (size_t)B * p.E * p.Z * 4, s);
        if (r == PSNODE_OK && dae && p.V) r = h2d(o_vj, hp->v_jump, (size_t)B * p.E * p.V * 4, s);
    }
    for (int net = 0; r == PSNODE_OK && net < (dae ? 2 : 1); net++) {
        const psnode_mlp& m = net ? hp->ae : hp->de;
        psnode_mlp& dm = net ? p.ae : p.de;
        for (int l = 0; r == PSNODE_OK && l < m.n_layers; l++) {
            r = h2d(o_W[net][l], m.W[l], (size_t)m.out_dim[l] * m.in_dim[l] * 4, s);
            if (r == PSNODE_OK) r = h2d(o_b[net][l], m.b[l], (size_t)m.out_dim[l] * 4, s);
            dm.W[l] = reinterpret_cast<const float*>(base + o_W[net][l]);
            dm.b[l] = reinterpret_cast<const float*>(base + o_b[net][l]);
        }
    }
    if (r != PSNODE_OK) return r;
    float* d_t = reinterpret_cast<float*>(base + o_t);
    float* d_z = reinterpret_cast<float*>(base + o_z);
    float* d_v = reinterpret_cast<float*>(base + o_v);
    float* d_xs = reinterpret_cast<float*>(base + o_xs);
    float* d_is = reinterpret_cast<float*>(base + o_is);
    // chunk c integrates steps (r0, r1]: grid rows r0..r1
    auto row0 = [&](int c) { return (int)((int64_t)(T - 1) * c / NCH); };
    auto copy_in = [&](int c) -> int {                 // input rows [lo, hi) of chunk c on the input stream
        const int lo = c == 0 ? 0 : row0(c) + 1, hi = row0(c + 1) + 1;
        if (hi <= lo) return PSNODE_OK;
        int rr = h2d(o_t + (size_t)lo * B * 4, hp->t.p + (size_t)lo * B, (size_t)(hi - lo) * B * 4, g_dma.s_in);
        if (rr == PSNODE_OK && p.Z) rr = h2d(o_z + (size_t)lo * B * p.Z * 4, hp->z.p + (size_t)lo * B * p.Z, (size_t)(hi - lo) * B * p.Z * 4, g_dma.s_in);
        if (rr == PSNODE_OK && dae && p.V) rr = h2d(o_v + (size_t)lo * B * p.V * 4, hp->v.p + (size_t)lo * B * p.V, (size_t)(hi - lo) * B * p.V * 4, g_dma.s_in);
        if (rr != PSNODE_OK) return rr;
        PSN_CUDA(cudaEventRecord(g_dma.in_ready[c], g_dma.s_in));
        return PSNODE_OK;
    };
    // the side streams start behind whatever the caller already queued on `s`
    PSN_CUDA(cudaEventRecord(g_dma.done[NCH], s));
    PSN_CUDA(cudaStreamWaitEvent(g_dma.s_in, g_dma.done[NCH], 0));
    PSN_CUDA(cudaStreamWaitEvent(g_dma.s_out, g_dma.done[NCH], 0));
    r = copy_in(0);
    if (r != PSNODE_OK) return r;
    for (int c = 0; c < NCH; c++) {
        if (c + 1 < NCH) { r = copy_in(c + 1); if (r != PSNODE_OK) return r; }
        const int r0 = row0(c), r1 = row0(c + 1);
        PSN_CUDA(cudaStreamWaitEvent(s, g_dma.in_ready[c], 0));
        psnode_problem q = p;
        q.T = r1 - r0 + 1;
        q.t = {d_t + (size_t)r0 * B, (int64_t)B, 1};
        q.z = {p.Z ? d_z + (size_t)r0 * B * p.Z : nullptr, (int64_t)B * p.Z, (int64_t)p.Z};
        q.v = {dae && p.V ? d_v + (size_t)r0 * B * p.V : nullptr, (int64_t)B * p.V, (int64_t)p.V};
        q.i = {nullptr, 0, 0};
        const float* xstart = c == 0 ? reinterpret_cast<const float*>(base + o_x0) : d_xs + (size_t)r0 * B * p.X;
        if (dae) { q.x_init = xstart; q.x_init_sb = p.X; q.x = {nullptr, 0, 0}; }
        else q.x = {xstart, (int64_t)B * p.X, (int64_t)p.X};
        q.a0 = reinterpret_cast<const float*>(base + o_a0); q.a0_sb = S;
        if (p.event_idx) {
            q.event_idx = reinterpret_cast<const int32_t*>(base + o_ev) + r0;
            q.z_jump = reinterpret_cast<const float*>(base + o_zj); q.zj_sb = (int64_t)p.E * p.Z; q.zj_se = p.Z;
            q.v_jump = reinterpret_cast<const float*>(base + o_vj); q.vj_sb = (int64_t)p.E * p.V; q.vj_se = p.V;
        }
        q.x_sol = {d_xs + (size_t)r0 * B * p.X, (int64_t)B * p.X, (int64_t)p.X};
        if (dae) q.i_sol = {d_is + (size_t)r0 * B * p.I, (int64_t)B * p.I, (int64_t)p.I};
        q.tape = nullptr; q.tape_floats = 0;
        if (q.T > 1 || c == 0) {
            const int st = psnode_forward(&q, base + o_ws, ws_bytes, s);
            if (st != PSNODE_OK) return st;
        }
        PSN_CUDA(cudaEventRecord(g_dma.done[c], s));
        PSN_CUDA(cudaStreamWaitEvent(g_dma.s_out, g_dma.done[c], 0));
        const int lo = c == 0 ? 0 : r0 + 1, hi = r1 + 1;       // trajectory rows this chunk finalised
        if (hi > lo) {
            PSN_CUDA(cudaMemcpyAsync(hp->x_sol.p + (size_t)lo * B * p.X, d_xs + (size_t)lo * B * p.X, (size_t)(hi - lo) * B * p.X * 4,
                                     cudaMemcpyDeviceToHost, g_dma.s_out));
            down += (int64_t)(hi - lo) * B * p.X * 4;
            if (dae) {
                PSN_CUDA(cudaMemcpyAsync(hp->i_sol.p + (size_t)lo * B * p.I, d_is + (size_t)lo * B * p.I, (size_t)(hi - lo) * B * p.I * 4,
                                         cudaMemcpyDeviceToHost, g_dma.s_out));
                down += (int64_t)(hi - lo) * B * p.I * 4;
            }
        }
    }
    PSN_CUDA(cudaStreamSynchronize(g_dma.s_out));
    PSN_CUDA(cudaStreamSynchronize(s));
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return PSNODE_OK;
}

extern "C" int psnode_forward_host(const psnode_problem* hp, void* stream, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!hp) return PSNODE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        const char* mode = std::getenv("PSNODE_HOST_PATH");
        const bool dae0 = hp->kind == PSNODE_DAE;
        if (mode && std::strcmp(mode, "dma") == 0 && hp->B >= 1 && hp->T >= 1 && !hp->teacher_x && !hp->teacher_i && hp->a0 &&
            hp->a0_sb == psn_S(hp) && dense_tm(hp->t, hp->B, 1) && dense_tm(hp->z, hp->B, hp->Z) && (!dae0 || dense_tm(hp->v, hp->B, hp->V)) &&
            dense_tm_out(hp->x_sol, hp->B, hp->X) && (!dae0 || (dense_tm_out(hp->i_sol, hp->B, hp->I) && hp->x_init && hp->x_init_sb == hp->X)) &&
            (dae0 || (hp->x.p && hp->x.sb == hp->X)) &&
            (!hp->event_idx || ((!hp->Z || (hp->zj_se == hp->Z && hp->zj_sb == (int64_t)hp->E * hp->Z)) &&
                                (!dae0 || !hp->V || (hp->vj_se == hp->V && hp->vj_sb == (int64_t)hp->E * hp->V)))))
            return forward_host_dma(hp, s, h2d_bytes, d2h_bytes);
    }
    psnode_problem p = *hp;
    const bool dae = p.kind == PSNODE_DAE;
    const int S = psn_S(&p);
    if (p.B < 1 || p.T < 1 || p.X < 1) return PSNODE_EINVAL;
    Plan plan;
    // ---- inputs ----------------------------------------------------------------------------------------
    struct SeriesRef { int idx = -1; int64_t lo = 0; };
    auto add_series = [&](const psnode_series& sr, int w, int T, bool needed) {
        SeriesRef r;
        if (!needed || !sr.p || w == 0) return r;
        const Span sp = span_of(sr.st, sr.sb, T, p.B, w);
        r.lo = sp.lo;
        r.idx = plan.add(sr.p + sp.lo, (size_t)(sp.hi - sp.lo) * 4);
        return r;
    };
    const SeriesRef r_t = add_series(p.t, 1, p.T, true);
    const SeriesRef r_x = add_series(p.x, p.X, p.teacher_x ? p.T : 1, !dae || p.teacher_x);   // ODE reads only x[0] unless teacher forcing
    const SeriesRef r_z = add_series(p.z, p.Z, p.T, true);
    const SeriesRef r_v = add_series(p.v, p.V, p.T, dae);
    const SeriesRef r_i = add_series(p.i, p.I, p.T, dae && p.teacher_i);
    if (!p.a0 || (dae && !p.x_init)) return PSNODE_EINVAL;
    const int i_xinit = dae ? plan.add(p.x_init, ((size_t)(p.B - 1) * p.x_init_sb + p.X) * 4) : -1;
    const int i_a0 = plan.add(p.a0, ((size_t)(p.B - 1) * p.a0_sb + S) * 4);
    int i_ev = -1, i_zj = -1, i_vj = -1;
    if (p.event_idx) {
        i_ev = plan.add(p.event_idx, (size_t)(p.T > 1 ? p.T - 1 : 1) * 4);
        if (p.Z) i_zj = plan.add(p.z_jump, ((size_t)(p.B - 1) * p.zj_sb + (size_t)(p.E - 1) * p.zj_se + p.Z) * 4);
        if (dae && p.V) i_vj = plan.add(p.v_jump, ((size_t)(p.B - 1) * p.vj_sb + (size_t)(p.E - 1) * p.vj_se + p.V) * 4);
    }
    int i_W[2][PSNODE_MAX_LAYERS], i_b[2][PSNODE_MAX_LAYERS];
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        const psnode_mlp& m = net ? p.ae : p.de;
        if (m.n_layers < 1 || m.n_layers > PSNODE_MAX_LAYERS) return PSNODE_EINVAL;
        for (int l = 0; l < m.n_layers; l++) {
            if (!m.W[l] || !m.b[l]) return PSNODE_EINVAL;
            i_W[net][l] = plan.add(m.W[l], (size_t)m.out_dim[l] * m.in_dim[l] * 4);
            i_b[net][l] = plan.add(m.b[l], (size_t)m.out_dim[l] * 4);
        }
    }
    // ---- outputs: written in place when the caller's buffer is device visible, else dense time-major in the arena -------
    if (!hp->x_sol.p || (dae && !hp->i_sol.p)) return PSNODE_EINVAL;
    const size_t xs_bytes = (size_t)p.T * p.B * p.X * 4, is_bytes = dae ? (size_t)p.T * p.B * p.I * 4 : 0;
    float* x_alias = static_cast<float*>(mapped_alias(hp->x_sol.p, xs_bytes));
    float* i_alias = dae ? static_cast<float*>(mapped_alias(hp->i_sol.p, is_bytes)) : nullptr;
    const size_t o_xsol = x_alias ? 0 : plan.reserve(xs_bytes);
    const size_t o_isol = (!dae || i_alias) ? 0 : plan.reserve(is_bytes);
    // ---- device problem ------------------------------------------------------------------------------------
    // The workspace size depends on shapes only, never on the pointer values.
    const int64_t ws_bytes = psnode_forward_workspace(&p);
    const size_t o_ws = plan.reserve((size_t)(ws_bytes > 0 ? ws_bytes : 256));
    char* base = g_scratch.get(plan.cur);
    if (!base) return psn_cuda_fail(cudaErrorMemoryAllocation, "psnode_forward_host scratch");
    for (const Plan::Item& it : plan.items)
        if (!it.alias) PSN_CUDA(cudaMemcpyAsync(base + it.off, it.src, it.bytes, cudaMemcpyDefault, s));
    auto dev_series = [&](psnode_series& sr, const SeriesRef& r) {
        sr.p = r.idx >= 0 ? reinterpret_cast<const float*>(plan.resolve(r.idx, base)) - r.lo : nullptr;
    };
    dev_series(p.t, r_t);
    dev_series(p.x, r_x);
    dev_series(p.z, r_z);
    dev_series(p.v, r_v);
    dev_series(p.i, r_i);
    if (dae) p.x_init = reinterpret_cast<const float*>(plan.resolve(i_xinit, base));
    p.a0 = reinterpret_cast<const float*>(plan.resolve(i_a0, base));
    if (p.event_idx) {
        p.event_idx = reinterpret_cast<const int32_t*>(plan.resolve(i_ev, base));
        if (i_zj >= 0) p.z_jump = reinterpret_cast<const float*>(plan.resolve(i_zj, base));
        if (i_vj >= 0) p.v_jump = reinterpret_cast<const float*>(plan.resolve(i_vj, base));
    }
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        psnode_mlp& m = net ? p.ae : p.de;
        for (int l = 0; l < m.n_layers; l++) {
            m.W[l] = reinterpret_cast<const float*>(plan.resolve(i_W[net][l], base));
            m.b[l] = reinterpret_cast<const float*>(plan.resolve(i_b[net][l], base));
        }
    }
    if (x_alias) p.x_sol.p = x_alias;     // strides stay the caller's
    else p.x_sol = {reinterpret_cast<float*>(base + o_xsol), (int64_t)p.B * p.X, (int64_t)p.X};
    if (dae) {
        if (i_alias) p.i_sol.p = i_alias;
        else p.i_sol = {reinterpret_cast<float*>(base + o_isol), (int64_t)p.B * p.I, (int64_t)p.I};
    }
    const int st = psnode_forward(&p, base + o_ws, ws_bytes, stream);
    if (st != PSNODE_OK) return st;
    // ---- results back ------------------------------------------------------------------------------------
    int64_t down = 0;
    auto copy_back = [&](const psnode_series_out& dst, const float* src, int w) -> int {
        if (dst.sb == w && dst.st == (int64_t)p.B * w) {
            PSN_CUDA(cudaMemcpyAsync(dst.p, src, (size_t)p.T * p.B * w * 4, cudaMemcpyDeviceToHost, s));
        } else if (dst.st == w && dst.sb == (int64_t)p.T * w) {   // batch-major destination: one 2-D copy per trajectory
            for (int b = 0; b < p.B; b++)
                PSN_CUDA(cudaMemcpy2DAsync(dst.p + (int64_t)b * dst.sb, (size_t)w * 4, src + (size_t)b * w, (size_t)p.B * w * 4,
                                           (size_t)w * 4, p.T, cudaMemcpyDeviceToHost, s));
        } else {
            return PSNODE_EUNSUPPORTED;
        }
        return PSNODE_OK;
    };
    if (!x_alias) { const int r = copy_back(hp->x_sol, p.x_sol.p, p.X); if (r != PSNODE_OK) return r; }
    down += (int64_t)xs_bytes;
    if (dae) {
        if (!i_alias) { const int r = copy_back(hp->i_sol, p.i_sol.p, p.I); if (r != PSNODE_OK) return r; }
        down += (int64_t)is_bytes;
    }
    PSN_CUDA(cudaStreamSynchronize(s));
    if (h2d_bytes) *h2d_bytes = plan.staged + plan.direct;     // bytes that crossed the bus host -> device (copied or read in place)
    if (d2h_bytes) *d2h_bytes = down;
    return PSNODE_OK;
}
