// psnode_host.cu -- psnode_forward_host: host-buffer entry point (stages inputs, integrates, copies results back).
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

// dense (rows, width) span covered by a strided (T,B,width) view: we copy the covering contiguous range
struct Span { int64_t lo, hi; };
Span span_of(int64_t st, int64_t sb, int T, int B, int w) {
    int64_t lo = 0, hi = 0;
    const int64_t a = st * (T - 1), b = sb * (B - 1);
    if (a < 0) lo += a; else hi += a;
    if (b < 0) lo += b; else hi += b;
    hi += w;
    return {lo, hi};
}
}  // namespace

extern "C" int psnode_forward_host(const psnode_problem* hp, void* stream, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!hp) return PSNODE_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    psnode_problem p = *hp;
    const bool dae = p.kind == PSNODE_DAE;
    const int S = psn_S(&p);
    // 1. plan the device arena
    struct Item { const void* src; size_t bytes; size_t off; };
    std::vector<Item> in;
    size_t cur = 0;
    auto reserve = [&](size_t bytes) { size_t o = cur; cur += (bytes + 255) & ~size_t(255); return o; };
    auto add_in = [&](const void* src, size_t bytes) { in.push_back({src, bytes, reserve(bytes)}); return in.back().off; };
    auto add_series = [&](const psnode_series& sr, int w, bool needed, size_t& off, int64_t& lo) {
        off = 0; lo = 0;
        if (!needed || !sr.p || w == 0) return;
        Span sp = span_of(sr.st, sr.sb, p.T, p.B, w);
        lo = sp.lo;
        off = add_in(sr.p + sp.lo, (size_t)(sp.hi - sp.lo) * 4);
    };
    size_t o_t, o_x, o_z, o_v, o_i; int64_t l_t, l_x, l_z, l_v, l_i;
    add_series(p.t, 1, true, o_t, l_t);
    // ODE needs only x[0] unless teacher forcing: stage the single initial row in that case
    psnode_series xs = p.x;
    int xT = p.T;
    if (!p.teacher_x) xT = 1;
    {
        o_x = 0; l_x = 0;
        if (xs.p && (!dae || p.teacher_x)) {
            Span sp = span_of(xs.st, xs.sb, xT, p.B, p.X);
            l_x = sp.lo; o_x = add_in(xs.p + sp.lo, (size_t)(sp.hi - sp.lo) * 4);
        }
    }
    add_series(p.z, p.Z, true, o_z, l_z);
    add_series(p.v, p.V, dae, o_v, l_v);
    add_series(p.i, p.I, dae && p.teacher_i, o_i, l_i);
    size_t o_xinit = 0, o_a0 = 0, o_ev = 0, o_zj = 0, o_vj = 0;
    if (dae) o_xinit = add_in(p.x_init, ((size_t)(p.B - 1) * p.x_init_sb + p.X) * 4);
    o_a0 = add_in(p.a0, ((size_t)(p.B - 1) * p.a0_sb + S) * 4);
    if (p.event_idx) {
        o_ev = add_in(p.event_idx, (size_t)(p.T - 1) * 4);
        if (p.Z) o_zj = add_in(p.z_jump, ((size_t)(p.B - 1) * p.zj_sb + (size_t)(p.E - 1) * p.zj_se + p.Z) * 4);
        if (dae && p.V) o_vj = add_in(p.v_jump, ((size_t)(p.B - 1) * p.vj_sb + (size_t)(p.E - 1) * p.vj_se + p.V) * 4);
    }
    size_t o_W[2][PSNODE_MAX_LAYERS], o_b[2][PSNODE_MAX_LAYERS];
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        const psnode_mlp& m = net ? p.ae : p.de;
        for (int l = 0; l < m.n_layers; l++) {
            o_W[net][l] = add_in(m.W[l], (size_t)m.out_dim[l] * m.in_dim[l] * 4);
            o_b[net][l] = add_in(m.b[l], (size_t)m.out_dim[l] * 4);
        }
    }
    // outputs: dense time-major on the device, copied back row by row into the caller's strided view
    const size_t xs_bytes = (size_t)p.T * p.B * p.X * 4, is_bytes = dae ? (size_t)p.T * p.B * p.I * 4 : 0;
    const size_t o_xsol = reserve(xs_bytes), o_isol = reserve(is_bytes ? is_bytes : 4);
    const size_t o_ws = reserve(0);
    // workspace size needs device-pointer-free info only
    const int64_t ws_bytes = psnode_forward_workspace(&p);
    cur += (size_t)ws_bytes;
    char* base = g_scratch.get(cur);
    if (!base) return psn_cuda_fail(cudaErrorMemoryAllocation, "psnode_forward_host scratch");
    int64_t up = 0;
    for (const Item& it : in) {
        PSN_CUDA(cudaMemcpyAsync(base + it.off, it.src, it.bytes, cudaMemcpyHostToDevice, s));
        up += (int64_t)it.bytes;
    }
    auto dev_series = [&](psnode_series& sr, size_t off, int64_t lo) { if (sr.p) sr.p = reinterpret_cast<const float*>(base + off) - lo; };
    dev_series(p.t, o_t, l_t);
    if (p.x.p && (!dae || p.teacher_x)) p.x.p = reinterpret_cast<const float*>(base + o_x) - l_x; else p.x.p = nullptr;
    if (p.Z) dev_series(p.z, o_z, l_z);
    if (dae && p.V) dev_series(p.v, o_v, l_v);
    if (dae && p.teacher_i) dev_series(p.i, o_i, l_i); else p.i.p = nullptr;
    if (dae) p.x_init = reinterpret_cast<const float*>(base + o_xinit);
    p.a0 = reinterpret_cast<const float*>(base + o_a0);
    if (p.event_idx) {
        p.event_idx = reinterpret_cast<const int32_t*>(base + o_ev);
        if (p.Z) p.z_jump = reinterpret_cast<const float*>(base + o_zj);
        if (dae && p.V) p.v_jump = reinterpret_cast<const float*>(base + o_vj);
    }
    for (int net = 0; net < (dae ? 2 : 1); net++) {
        psnode_mlp& m = net ? p.ae : p.de;
        for (int l = 0; l < m.n_layers; l++) {
            m.W[l] = reinterpret_cast<const float*>(base + o_W[net][l]);
            m.b[l] = reinterpret_cast<const float*>(base + o_b[net][l]);
        }
    }
    p.x_sol = {reinterpret_cast<float*>(base + o_xsol), (int64_t)p.B * p.X, (int64_t)p.X};
    if (dae) p.i_sol = {reinterpret_cast<float*>(base + o_isol), (int64_t)p.B * p.I, (int64_t)p.I};
    const int st = psnode_forward(&p, base + o_ws, ws_bytes, stream);
    if (st != PSNODE_OK) return st;
    int64_t down = 0;
    auto copy_back = [&](const psnode_series_out& dst, const float* src, int w) -> int {
        if (dst.sb == w && dst.st == (int64_t)p.B * w) {
            PSN_CUDA(cudaMemcpyAsync(dst.p, src, (size_t)p.T * p.B * w * 4, cudaMemcpyDeviceToHost, s));
        } else if (dst.st == w && dst.sb == (int64_t)p.T * w) {   // batch-major destination: 2-D copy per trajectory block
            for (int b = 0; b < p.B; b++)
                PSN_CUDA(cudaMemcpy2DAsync(dst.p + (int64_t)b * dst.sb, (size_t)w * 4, src + (size_t)b * w, (size_t)p.B * w * 4,
                                           (size_t)w * 4, p.T, cudaMemcpyDeviceToHost, s));
        } else {
            return PSNODE_EUNSUPPORTED;
        }
        down += (int64_t)p.T * p.B * w * 4;
        return PSNODE_OK;
    };
    int r = copy_back(hp->x_sol, p.x_sol.p, p.X);
    if (r != PSNODE_OK) return r;
    if (dae) { r = copy_back(hp->i_sol, p.i_sol.p, p.I); if (r != PSNODE_OK) return r; }
    PSN_CUDA(cudaStreamSynchronize(s));
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return PSNODE_OK;
}
