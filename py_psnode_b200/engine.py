"""Host-side glue between torch tensors and the C ABI: builds `psnode_problem`, owns the workspace, launches the
forward kernel and the reverse sweep, and exposes both as one `torch.autograd.Function`.

torch is plumbing here (device memory, streams, autograd bookkeeping); every FLOP of the integration runs in
libpsnode_b200.so.  If the library is missing or the tensors are not CUDA fp32 the call raises.
"""
import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N

_workspaces = {}      # device index -> uint8 tensor


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _workspaces.get(key)
    need = max(int(nbytes), 256)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _require_cuda_f32(name: str, ten: torch.Tensor) -> None:
    if not ten.is_cuda:
        raise RuntimeError(
            f"py_psnode_b200: `{name}` lives on {ten.device}; the fused integrator runs on CUDA (sm_100a) only and has "
            "no CPU fallback.  Move the batch and the model to a CUDA device.")
    if ten.dtype != torch.float32:
        raise TypeError(f"py_psnode_b200: `{name}` has dtype {ten.dtype}; the integrator computes in float32 like the reference")


def _series(ten: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    """(T,B,W) view with unit stride over W (copy only if the feature stride is not 1); None for zero width."""
    if ten is None or ten.shape[-1] == 0:
        return None
    _require_cuda_f32(name, ten)
    if ten.dim() != 3:
        raise ValueError(f"`{name}` must be (T,B,width), got {tuple(ten.shape)}")
    if ten.stride(2) != 1 and ten.shape[2] != 1:
        ten = ten.contiguous()
    return ten


def _rows(ten: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    """(B,W) matrix with unit stride over W."""
    if ten is None or ten.shape[-1] == 0:
        return None
    _require_cuda_f32(name, ten)
    if ten.stride(-1) != 1 and ten.shape[-1] != 1:
        ten = ten.contiguous()
    return ten


def _set_series(dst: N.Series, ten: Optional[torch.Tensor]) -> None:
    if ten is None:
        dst.p, dst.st, dst.sb = None, 0, 0
    else:
        dst.p, dst.st, dst.sb = ten.data_ptr(), ten.stride(0), ten.stride(1)


def _fill_mlp(dst: N.Mlp, params: Sequence[torch.Tensor], keep: list) -> None:
    """params = [W0, b0, W1, b1, ...] (nn.Linear layout)."""
    n = len(params) // 2
    dst.n_layers = n
    for l in range(n):
        W, b = params[2 * l], params[2 * l + 1]
        _require_cuda_f32("weight", W)
        Wc, bc = W.detach().contiguous(), b.detach().contiguous()
        keep.extend((Wc, bc))
        dst.in_dim[l], dst.out_dim[l] = Wc.shape[1], Wc.shape[0]
        dst.W[l], dst.b[l] = Wc.data_ptr(), bc.data_ptr()


@dataclass
class Config:
    """Static (non-tensor) description of one integrate_* call."""
    kind: int
    method: int
    impl: int
    X: int
    Z: int
    V: int
    I: int
    teacher_x: bool
    teacher_i: bool
    n_de: int                       # number of Linear layers of the DE net
    n_ae: int
    has_event: bool
    check_events: bool = True
    event_ref: Optional[tuple] = None   # (t_row (T,), ev_row (E,)) of the GLOBAL sample 0 in batch-sharded runs


# fixed positional layout of the tensor arguments of _Integrate.apply
_T, _X, _Zs, _Vs, _Is, _XINIT, _A0, _EVT, _ZJ, _VJ, _NFIXED = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10


def _build_problem(cfg: Config, tens: Sequence[Optional[torch.Tensor]], x_sol, i_sol, keep: list,
                   tape: Optional[torch.Tensor] = None) -> N.Problem:
    t = _series(tens[_T], "t")
    T, B = t.shape[0], t.shape[1]
    x = _series(tens[_X], "x")
    z = _series(tens[_Zs], "z")
    v = _series(tens[_Vs], "v")
    i = _series(tens[_Is], "i")
    x_init = _rows(tens[_XINIT], "x_init")
    a0 = _rows(tens[_A0], "all_initial")
    keep.extend((t, x, z, v, i, x_init, a0))
    p = N.Problem()
    p.kind, p.method, p.impl = cfg.kind, cfg.method, cfg.impl
    p.B, p.T = B, T
    p.X, p.Z, p.V, p.I = cfg.X, cfg.Z, cfg.V, cfg.I
    p.teacher_x, p.teacher_i = int(cfg.teacher_x), int(cfg.teacher_i)
    _set_series(p.t, t)
    _set_series(p.x, x if (cfg.kind == N.ODE or cfg.teacher_x) else None)
    _set_series(p.z, z)
    _set_series(p.v, v)
    _set_series(p.i, i if cfg.teacher_i else None)
    if x_init is not None:
        p.x_init, p.x_init_sb = x_init.data_ptr(), x_init.stride(0)
    S = cfg.X + cfg.Z + cfg.V + cfg.I
    if a0 is None or a0.shape[-1] != S or a0.shape[0] != B:
        raise ValueError(f"all_initial must be (B, X+Z+V+I) = ({B}, {S}), got {None if a0 is None else tuple(a0.shape)}")
    p.a0, p.a0_sb = a0.data_ptr(), a0.stride(0)
    p.E = 0
    if cfg.has_event:
        ev_t = tens[_EVT]
        _require_cuda_f32("event_t", ev_t)
        E = ev_t.shape[1]
        ev0 = ev_t[0].reshape(E)                    # event times of sample 0 (the only sample the reference inspects)
        t00 = t[:, 0, 0]
        if cfg.event_ref is not None:               # batch-sharded: the global batch's sample 0, pinned by parallel.py
            t00, ev0 = cfg.event_ref
            _require_cuda_f32("event reference t row", t00)
            _require_cuda_f32("event reference event row", ev0)
            if t00.numel() != T or ev0.numel() != E:
                raise ValueError("pinned event reference rows do not match (T,) / (E,)")
        idx = torch.empty(max(T - 1, 1), dtype=torch.int32, device=t.device)
        err = torch.zeros(1, dtype=torch.int32, device=t.device)
        keep.extend((ev0, t00, idx, err))
        N.check(N.lib().psnode_event_table(t00.data_ptr(), t00.stride(0), T, ev0.data_ptr(), ev0.stride(0), E,
                                           idx.data_ptr(), err.data_ptr(), torch.cuda.current_stream(t.device).cuda_stream),
                "psnode_event_table")
        if cfg.check_events and int(err.item()) != 0:
            raise RuntimeError("more than one event matches the same grid time (the reference raises here too)")
        p.event_idx, p.E = idx.data_ptr(), E
        zj = tens[_ZJ]
        if cfg.Z > 0:
            _require_cuda_f32("z_jump", zj)
            if zj.stride(-1) != 1 and zj.shape[-1] != 1:
                zj = zj.contiguous()
            keep.append(zj)
            p.z_jump, p.zj_sb, p.zj_se = zj.data_ptr(), zj.stride(0), zj.stride(1)
        if cfg.kind == N.DAE and cfg.V > 0:
            vj = tens[_VJ]
            _require_cuda_f32("v_jump", vj)
            if vj.stride(-1) != 1 and vj.shape[-1] != 1:
                vj = vj.contiguous()
            keep.append(vj)
            p.v_jump, p.vj_sb, p.vj_se = vj.data_ptr(), vj.stride(0), vj.stride(1)
    params = tens[_NFIXED:]
    _fill_mlp(p.de, params[:2 * cfg.n_de], keep)
    if cfg.kind == N.DAE:
        _fill_mlp(p.ae, params[2 * cfg.n_de:2 * (cfg.n_de + cfg.n_ae)], keep)
    _set_series(p.x_sol, x_sol)
    _set_series(p.i_sol, i_sol)
    if tape is not None:
        p.tape, p.tape_floats = tape.data_ptr(), tape.numel()
    return p


def _tape_budget_bytes(device: torch.device) -> int:
    """How much HBM one call may spend on the activation tape (psnode_b200.h `tape`): PSNODE_TAPE_MAX_GB, default 60 % of
    the memory that is free right now (B200: 180 GB; the cfg2 tape is 13.6 GB)."""
    env = os.environ.get("PSNODE_TAPE_MAX_GB")
    if env is not None:
        return int(float(env) * 2 ** 30)
    free, _total = torch.cuda.mem_get_info(device)
    reusable = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    return int(0.6 * (free + reusable))


_tape_pool = {}      # device index -> idle tape buffer (one per device; a second concurrent call allocates its own)


def _take_tape(device: torch.device, n_floats: int) -> torch.Tensor:
    """Lease a tape buffer: the idle pooled one if it is large enough, else a fresh allocation (13.6 GB at cfg2 -- not
    something to hand back to the caching allocator and re-request every optimiser step)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    buf = _tape_pool.pop(key, None)
    if buf is None or buf.numel() < n_floats:
        buf = None                      # drop the smaller buffer before asking for the larger one
        buf = torch.empty(n_floats, dtype=torch.float32, device=device)
    return buf


def _give_tape(buf: Optional[torch.Tensor]) -> None:
    """Return a leased tape buffer after the reverse sweep consumed it (stream order keeps the next forward behind it)."""
    if buf is None:
        return
    key = buf.device.index
    old = _tape_pool.get(key)
    if old is None or old.numel() < buf.numel():
        _tape_pool[key] = buf


def release_tape_pool() -> None:
    """Free the pooled tape buffers (e.g. before an evaluation phase that needs the memory)."""
    _tape_pool.clear()


class TapeChunks:
    """Marker returned instead of a tape when the activation tape of the whole batch would not fit in HBM."""

    def __init__(self, rows: int):
        self.rows = rows


def forward_raw(cfg: Config, tens: Sequence[Optional[torch.Tensor]], want_tape: bool = False, input_grads: bool = False,
                force_tape: bool = False):
    """Run the forward kernel; returns time-major contiguous (x_sol, i_sol) -- and, with `want_tape`, the activation tape
    the tensor-core reverse sweep consumes (None when the problem has no tape-based sweep or the tape would not fit).
    `input_grads`: the caller will ask for input-series / jump gradients, so a tape is only worth recording if the tape-based
    sweep of this problem produces them (psnode_tape_covers_input_grads).  `force_tape`: skip the HBM budget check (the
    chunked reverse sweep sized its chunks against the budget already)."""
    t = tens[_T]
    _require_cuda_f32("t", t)
    L = N.lib()
    T, B = t.shape[0], t.shape[1]
    with torch.cuda.device(t.device):
        x_sol = torch.empty((T, B, cfg.X), dtype=torch.float32, device=t.device)
        i_sol = torch.empty((T, B, cfg.I), dtype=torch.float32, device=t.device) if cfg.kind == N.DAE else None
        keep: list = []
        p = _build_problem(cfg, tens, x_sol, i_sol, keep)
        tape = None
        if want_tape and input_grads and not L.psnode_tape_covers_input_grads(C.byref(p)):
            want_tape = False
        if want_tape:
            n_tape = int(L.psnode_tape_floats(C.byref(p)))
            key = t.device.index if t.device.index is not None else torch.cuda.current_device()
            pooled = _tape_pool.get(key)
            if n_tape > 0 and (force_tape
                               or (pooled is not None and pooled.numel() >= n_tape and os.environ.get("PSNODE_TAPE_MAX_GB") is None)
                               or n_tape * 4 <= _tape_budget_bytes(t.device)):
                tape = _take_tape(t.device, n_tape)
                p.tape, p.tape_floats = tape.data_ptr(), tape.numel()
            elif n_tape > 0:
                # the whole batch's tape does not fit: the reverse sweep will re-integrate and differentiate the batch in
                # chunks of `TapeChunks.rows` trajectories (each with its own tape) instead of falling back to the generic sweep
                per_group = n_tape // ((B + 15) // 16)
                rows = 16 * int(_tape_budget_bytes(t.device) // (4 * per_group))
                if rows >= 32:
                    tape = TapeChunks(rows)
        ws = _workspace(t.device, L.psnode_forward_workspace(C.byref(p)))
        stream = torch.cuda.current_stream(t.device).cuda_stream
        N.check(L.psnode_forward(C.byref(p), ws.data_ptr(), ws.numel(), stream), "psnode_forward")
    return x_sol, i_sol, tape


def _theta_sizes(params: Sequence[torch.Tensor]) -> List[int]:
    return [int(q.numel()) for q in params]


class _Integrate(torch.autograd.Function):
    """x_sol, i_sol = integrate(cfg, t, x, z, v, i, x_init, a0, event_t, z_jump, v_jump, *weights)."""

    @staticmethod
    def forward(ctx, cfg: Config, *tens):
        needs = ctx.needs_input_grad[1:]
        # the tape-based reverse sweeps produce parameter, x0 and all_initial gradients -- and, for the latent `*_02` nets, the
        # input-series / jump gradients too; anything else is recomputed from the stored trajectory
        input_grads = any(needs[k] for k in (_Zs, _Vs, _Is, _ZJ, _VJ))
        x_sol, i_sol, tape = forward_raw(cfg, tens, want_tape=not (cfg.teacher_x or cfg.teacher_i), input_grads=input_grads)
        ctx.tape = tape
        ctx.cfg = cfg
        ctx.n_in = len(tens)
        ctx.save_for_backward(*[q for q in tens if q is not None], x_sol, *( [i_sol] if i_sol is not None else []))
        ctx.present = [q is not None for q in tens]
        if i_sol is None:
            i_sol = x_sol.new_empty(0)
            ctx.mark_non_differentiable(i_sol)
        return x_sol, i_sol

    @staticmethod
    def backward(ctx, gx, gi):
        cfg: Config = ctx.cfg
        saved = list(ctx.saved_tensors)
        tens: List[Optional[torch.Tensor]] = []
        it = iter(saved)
        for present in ctx.present:
            tens.append(next(it) if present else None)
        x_sol = next(it)
        i_sol = next(it) if cfg.kind == N.DAE else None
        tape, ctx.tape = ctx.tape, None
        needs = ctx.needs_input_grad[1:]
        if isinstance(tape, TapeChunks):
            grads = _backward_chunked(cfg, tens, gx, gi if cfg.kind == N.DAE else None, needs, tape.rows)
        else:
            grads = _backward_or_recompute(cfg, tens, x_sol, i_sol, gx, gi if cfg.kind == N.DAE else None, needs, tape)
            _give_tape(tape)
        return (None, *grads)


def _backward_chunked(cfg: Config, tens, gx, gi, needs, rows: int) -> List[Optional[torch.Tensor]]:
    """Tape-based reverse sweep for a batch whose tape does not fit: trajectories are independent, so the batch is
    re-integrated chunk by chunk (forward with tape, then the tensor-core sweep), parameter gradients are summed and the
    per-trajectory gradients concatenated."""
    import dataclasses
    t = tens[_T]
    T, B = t.shape[0], t.shape[1]
    if cfg.has_event and cfg.event_ref is None:      # every chunk must test the events of the GLOBAL sample 0
        E = tens[_EVT].shape[1]
        cfg = dataclasses.replace(cfg, event_ref=(t[:, 0, 0], tens[_EVT][0].reshape(E)))
    series = (_T, _X, _Zs, _Vs, _Is)
    rows_major = (_XINIT, _A0, _EVT, _ZJ, _VJ)
    total: List[Optional[torch.Tensor]] = [None] * len(tens)
    parts: List[list] = [[] for _ in tens]
    if gx is None:
        gx = torch.zeros((T, B, cfg.X), dtype=torch.float32, device=t.device)
    if cfg.kind == N.DAE and gi is None:
        gi = torch.zeros((T, B, cfg.I), dtype=torch.float32, device=t.device)
    for b0 in range(0, B, rows):
        b1 = min(B, b0 + rows)
        sub = list(tens)
        for k in series:
            if sub[k] is not None:
                sub[k] = sub[k][:, b0:b1]
        for k in rows_major:
            if sub[k] is not None:
                sub[k] = sub[k][b0:b1]
        # the chunk size was derived from the budget at forward time; by now x_sol, the upstream gradient and the loss
        # temporaries are allocated, so the budget is NOT re-checked here: take the tape, and if the allocator cannot
        # provide it fall back to the recomputing sweep for this chunk instead of failing the training step
        input_grads = any(needs[k] for k in (_Zs, _Vs, _Is, _ZJ, _VJ))
        try:
            xs, _is, tape = forward_raw(cfg, sub, want_tape=True, input_grads=input_grads, force_tape=True)
        except torch.cuda.OutOfMemoryError:
            release_tape_pool()
            torch.cuda.empty_cache()
            xs, _is, tape = forward_raw(cfg, sub, want_tape=False)
        if isinstance(tape, TapeChunks):
            tape = None
        g = _backward_or_recompute(cfg, sub, xs, _is, gx[:, b0:b1], gi[:, b0:b1] if gi is not None else None, needs, tape)
        _give_tape(tape)
        for k, gk in enumerate(g):
            if gk is None:
                continue
            if k >= _NFIXED:
                total[k] = gk.clone() if total[k] is None else total[k].add_(gk)
            else:
                parts[k].append(gk)
    for k in range(_NFIXED):
        if parts[k]:
            total[k] = torch.cat(parts[k], dim=1 if k in series else 0)
    return total


@dataclass
class LossSpec:
    """Masked squared-error terms fused with the integration: sum w_c * mask * (sol - target)^2 over the
    trajectory (x term) and, for a DAE, over the algebraic trajectory (i term).  Time-major (T,B,.) views."""
    target_x: torch.Tensor
    mask: torch.Tensor
    weight_x: Optional[torch.Tensor] = None
    target_i: Optional[torch.Tensor] = None
    weight_i: Optional[torch.Tensor] = None


def _set_term(term: "N.LossTerm", target, mask, weight, scale, keep) -> None:
    target, mask = _series(target, "loss target"), _series(mask, "loss mask")
    keep.extend((target, mask, weight, scale))
    _set_series(term.target, target)
    _set_series(term.mask, mask)
    term.feat_weight = weight.data_ptr() if weight is not None else None
    term.scale = scale.data_ptr()


def backward_raw(cfg: Config, tens, x_sol, i_sol, gx, gi, needs, tape=None, fuse: Optional[LossSpec] = None,
                 fuse_scale: Optional[torch.Tensor] = None) -> List[Optional[torch.Tensor]]:
    """Reverse sweep through the native library.  `needs[k]` says whether tens[k] wants a gradient; `tape` is what
    forward_raw(..., want_tape=True) recorded (None: the sweep recomputes the stages from x_sol).  With `fuse` the upstream
    gradient is the masked-MSE gradient scaled by the device scalar `fuse_scale`: the tensor-core sweeps form it on the fly
    (psnode_adjoint.fuse_x / fuse_i), the recomputing sweeps get it materialised by psnode_masked_sse_grad."""
    L = N.lib()
    t = tens[_T]
    dev = t.device
    T, B = t.shape[0], t.shape[1]
    X, Z, V, I = cfg.X, cfg.Z, cfg.V, cfg.I
    S = X + Z + V + I
    dae = cfg.kind == N.DAE
    out: List[Optional[torch.Tensor]] = [None] * len(tens)
    with torch.cuda.device(dev):
        keep: list = []
        p = _build_problem(cfg, tens, x_sol, i_sol, keep, tape)
        a = N.Adjoint()
        if fuse is not None:
            _set_term(a.fuse_x, fuse.target_x, fuse.mask, fuse.weight_x, fuse_scale, keep)
            if dae and fuse.target_i is not None:
                _set_term(a.fuse_i, fuse.target_i, fuse.mask, fuse.weight_i, fuse_scale, keep)
            elif dae:
                gi = torch.zeros_like(i_sol)
                _set_series(a.gi, gi)
        else:
            gx = torch.zeros_like(x_sol) if gx is None else _series(gx, "grad x_sol")
            _set_series(a.gx, gx)
            if dae:
                gi = torch.zeros_like(i_sol) if gi is None else _series(gi, "grad i_sol")
                _set_series(a.gi, gi)
        params = tens[_NFIXED:]
        sizes = _theta_sizes(params)
        d_theta = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        a.d_theta, a.n_theta = d_theta.data_ptr(), d_theta.numel()
        d_x0 = torch.empty((B, X), dtype=torch.float32, device=dev)
        a.d_x0, a.d_x0_sb = d_x0.data_ptr(), X
        d_a0 = None
        if needs[_A0]:
            d_a0 = torch.empty((B, S), dtype=torch.float32, device=dev)
            a.d_a0, a.d_a0_sb = d_a0.data_ptr(), S
        d_z = d_v = d_zj = d_vj = d_xt = d_it = None
        if Z > 0 and needs[_Zs]:
            d_z = torch.empty((T, B, Z), dtype=torch.float32, device=dev)
            _set_series(a.d_z, d_z)
        if dae and V > 0 and needs[_Vs]:
            d_v = torch.empty((T, B, V), dtype=torch.float32, device=dev)
            _set_series(a.d_v, d_v)
        if cfg.has_event and Z > 0 and needs[_ZJ]:
            E = tens[_ZJ].shape[1]
            d_zj = torch.empty((B, E, Z), dtype=torch.float32, device=dev)
            a.d_zjump, a.d_zj_sb, a.d_zj_se = d_zj.data_ptr(), E * Z, Z
        if cfg.has_event and dae and V > 0 and needs[_VJ]:
            E = tens[_VJ].shape[1]
            d_vj = torch.empty((B, E, V), dtype=torch.float32, device=dev)
            a.d_vjump, a.d_vj_sb, a.d_vj_se = d_vj.data_ptr(), E * V, V
        if cfg.teacher_x and needs[_X]:
            d_xt = torch.empty((T, B, X), dtype=torch.float32, device=dev)
            _set_series(a.d_xteach, d_xt)
        if cfg.teacher_i and needs[_Is]:
            d_it = torch.empty((T, B, I), dtype=torch.float32, device=dev)
            _set_series(a.d_iteach, d_it)
        if fuse is not None and not L.psnode_sweep_fuses_loss(C.byref(p), C.byref(a)):
            # recomputing sweep: materialise the masked-MSE gradient with the fused loss kernel and pass it as gx / gi
            from .losses import masked_sse_grad_into
            gx = masked_sse_grad_into(x_sol, fuse.target_x, fuse.mask, fuse.weight_x, fuse_scale)
            _set_series(a.gx, gx)
            a.fuse_x = N.LossTerm()
            if dae:
                gi = (masked_sse_grad_into(i_sol, fuse.target_i, fuse.mask, fuse.weight_i, fuse_scale) if fuse.target_i is not None
                      else torch.zeros_like(i_sol))
                _set_series(a.gi, gi)
                a.fuse_i = N.LossTerm()
        ws = _workspace(dev, L.psnode_backward_workspace(C.byref(p), C.byref(a)))
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(L.psnode_backward(C.byref(p), C.byref(a), ws.data_ptr(), ws.numel(), stream), "psnode_backward")
    # ---- route the native outputs to the autograd inputs -------------------------------------------
    if dae:
        if needs[_XINIT]:
            out[_XINIT] = d_x0
        if needs[_X] and tens[_X] is not None and tens[_X].shape[-1] != 0:
            out[_X] = d_xt if d_xt is not None else torch.zeros_like(tens[_X])
    else:
        if needs[_X]:
            g = d_xt if d_xt is not None else torch.zeros((T, B, X), dtype=torch.float32, device=dev)
            g[0] += d_x0                       # x[0] is the initial state (and x_sol[0])
            out[_X] = g
    out[_Zs], out[_Vs], out[_Is] = d_z, d_v, d_it
    out[_A0], out[_ZJ], out[_VJ] = d_a0, d_zj, d_vj
    off = 0
    for k, n in enumerate(sizes):
        if needs[_NFIXED + k]:
            out[_NFIXED + k] = d_theta[off:off + n].view_as(params[k])
        off += n
    return out


def _backward_or_recompute(cfg: Config, tens, x_sol, i_sol, gx, gi, needs, tape, **kw) -> List[Optional[torch.Tensor]]:
    """backward_raw on the tape; if the tape-based sweep's own workspace (delta records of the latent / hidden-128 paths: up to the
    size of the tape again) does not fit next to the tape, drop the tape and run the recomputing sweep instead of failing the step."""
    if tape is None:
        return backward_raw(cfg, tens, x_sol, i_sol, gx, gi, needs, None, **kw)
    try:
        return backward_raw(cfg, tens, x_sol, i_sol, gx, gi, needs, tape, **kw)
    except torch.cuda.OutOfMemoryError:
        release_tape_pool()
        torch.cuda.empty_cache()
        return backward_raw(cfg, tens, x_sol, i_sol, gx, gi, needs, None, **kw)


class _IntegrateLoss(torch.autograd.Function):
    """num, x_sol, i_sol = integrate_loss(cfg, spec, t, x, ...): the integration fused with the masked squared-error
    numerator of the scripts' loss.  Only `num` is differentiable; its backward runs the reverse sweep with the loss gradient
    formed inside the sweep, so dL/dx_sol (T,B,X) never exists (cfg2: 262 MB written and read per step otherwise)."""

    @staticmethod
    def forward(ctx, cfg: Config, spec: LossSpec, *tens):
        from .losses import masked_sse
        needs = ctx.needs_input_grad[2:]
        input_grads = any(needs[k] for k in (_Zs, _Vs, _Is, _ZJ, _VJ))
        x_sol, i_sol, tape = forward_raw(cfg, tens, want_tape=not (cfg.teacher_x or cfg.teacher_i), input_grads=input_grads)
        with torch.no_grad():
            num = masked_sse(x_sol, spec.target_x, spec.mask, spec.weight_x)
            if i_sol is not None and spec.target_i is not None:
                num = num + masked_sse(i_sol, spec.target_i, spec.mask, spec.weight_i)
        ctx.tape, ctx.cfg, ctx.spec = tape, cfg, spec
        # the trajectories are returned as non-differentiable by-products: without this autograd would hand backward() freshly
        # zero-filled (T,B,X) / (T,B,I) gradients for them (2 x 15.6 GiB at the cfg5 shard)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(*[q for q in tens if q is not None], x_sol, *([i_sol] if i_sol is not None else []))
        ctx.present = [q is not None for q in tens]
        if i_sol is None:
            i_sol = x_sol.new_empty(0)
        ctx.mark_non_differentiable(x_sol, i_sol)
        return num, x_sol, i_sol

    @staticmethod
    def backward(ctx, gnum, _gx, _gi):
        cfg: Config = ctx.cfg
        saved = list(ctx.saved_tensors)
        it = iter(saved)
        tens = [next(it) if present else None for present in ctx.present]
        x_sol = next(it)
        i_sol = next(it) if cfg.kind == N.DAE else None
        tape, ctx.tape = ctx.tape, None
        needs = ctx.needs_input_grad[2:]
        if gnum is None:
            return (None, None) + (None,) * len(ctx.present)
        scale = gnum.detach().to(torch.float32).reshape(1).contiguous()
        if isinstance(tape, TapeChunks):
            tape = None                     # chunked re-integration is not combined with loss fusion: recomputing sweep
        grads = _backward_or_recompute(cfg, tens, x_sol, i_sol, None, None, needs, tape, fuse=ctx.spec, fuse_scale=scale)
        _give_tape(tape)
        return (None, None, *grads)


def integrate_loss(cfg: Config, spec: LossSpec, tens: Sequence[Optional[torch.Tensor]]):
    """(loss numerator, x_sol, i_sol) -- see _IntegrateLoss."""
    num, x_sol, i_sol = _IntegrateLoss.apply(cfg, spec, *tens)
    return num, x_sol, (i_sol if cfg.kind == N.DAE else None)


def integrate(cfg: Config, tens: Sequence[Optional[torch.Tensor]]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Autograd-aware entry used by the solver classes."""
    needs_grad = torch.is_grad_enabled() and any(q is not None and q.requires_grad for q in tens)
    if not needs_grad:
        return forward_raw(cfg, tens)[:2]
    x_sol, i_sol = _Integrate.apply(cfg, *tens)
    return x_sol, (i_sol if cfg.kind == N.DAE else None)


# ------------------------------------------------------------------------------------------------ Init_Func + all_initial in one launch
def _row(ten: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if ten is None or ten.shape[-1] == 0:
        return None
    _require_cuda_f32(name, ten)
    if ten.dim() != 2:
        raise ValueError(f"`{name}` must be (B, width), got {tuple(ten.shape)}")
    return ten if ten.stride(1) == 1 or ten.shape[1] == 1 else ten.contiguous()


class _InitState(torch.autograd.Function):
    """x0, all_initial = init_state(z0, v0, i0, *init_params): `psnode_init_state` / `psnode_init_state_backward`."""

    @staticmethod
    def forward(ctx, z0, v0, i0, *params):
        L = N.lib()
        rows = [_row(z0, "z0"), _row(v0, "v0"), _row(i0, "i0")]
        ref = next(r for r in rows if r is not None)
        dev, B = ref.device, ref.shape[0]
        keep: list = []
        mlp = N.Mlp()
        _fill_mlp(mlp, params, keep)
        X = params[-2].shape[0]
        widths = [0 if r is None else r.shape[1] for r in rows]
        S = X + sum(widths)
        with torch.cuda.device(dev):
            x0 = torch.empty((B, X), dtype=torch.float32, device=dev)
            a0 = torch.empty((B, S), dtype=torch.float32, device=dev)
            ptr = lambda r: (None, 0) if r is None else (r.data_ptr(), r.stride(0))
            (zp, zs), (vp, vs), (ip, is_) = ptr(rows[0]), ptr(rows[1]), ptr(rows[2])
            N.check(L.psnode_init_state(C.byref(mlp), zp, zs, vp, vs, ip, is_, B, widths[0], widths[1], widths[2], x0.data_ptr(), X, a0.data_ptr(), S,
                                        torch.cuda.current_stream(dev).cuda_stream), "psnode_init_state")
        ctx.save_for_backward(*[r for r in rows if r is not None], *params)
        ctx.present = [r is not None for r in rows]
        ctx.n_params = len(params)
        return x0, a0

    @staticmethod
    def backward(ctx, gx0, ga0):
        L = N.lib()
        saved = list(ctx.saved_tensors)
        it = iter(saved)
        rows = [next(it) if pr else None for pr in ctx.present]
        params = [next(it) for _ in range(ctx.n_params)]
        ref = next(r for r in rows if r is not None)
        dev, B = ref.device, ref.shape[0]
        keep: list = []
        mlp = N.Mlp()
        _fill_mlp(mlp, params, keep)
        widths = [0 if r is None else r.shape[1] for r in rows]
        gx0 = None if gx0 is None else gx0.contiguous()
        ga0 = None if ga0 is None else ga0.contiguous()
        if gx0 is None and ga0 is None:
            return (None,) * (3 + ctx.n_params)
        with torch.cuda.device(dev):
            sizes = _theta_sizes(params)
            d_theta = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
            d_rows = [None if r is None else torch.empty((B, r.shape[1]), dtype=torch.float32, device=dev) for r in rows]
            ws = _workspace(dev, L.psnode_init_state_backward_workspace(C.byref(mlp), B))
            ptr = lambda r: (None, 0) if r is None else (r.data_ptr(), r.stride(0))
            (zp, zs), (vp, vs), (ip, is_) = ptr(rows[0]), ptr(rows[1]), ptr(rows[2])
            (dzp, dzs), (dvp, dvs), (dip, dis) = ptr(d_rows[0]), ptr(d_rows[1]), ptr(d_rows[2])
            gxp, gxs = ptr(gx0)
            gap, gas = ptr(ga0)
            N.check(L.psnode_init_state_backward(C.byref(mlp), zp, zs, vp, vs, ip, is_, B, widths[0], widths[1], widths[2], gxp, gxs, gap, gas,
                                                 d_theta.data_ptr(), dzp, dzs, dvp, dvs, dip, dis, ws.data_ptr(), ws.numel(),
                                                 torch.cuda.current_stream(dev).cuda_stream), "psnode_init_state_backward")
        grads, off = [], 0
        for k, n in enumerate(sizes):
            grads.append(d_theta[off:off + n].view_as(params[k]) if ctx.needs_input_grad[3 + k] else None)
            off += n
        return (*[g if ctx.needs_input_grad[k] else None for k, g in enumerate(d_rows)], *grads)


def init_state(params: Sequence[torch.Tensor], z0, v0, i0):
    return _InitState.apply(z0, v0, i0, *params)


# ------------------------------------------------------------------------------------------------ encoded entry
def forward_encoded(cfg: Config, t, x0, a0, z_raw, v_raw, event_t, zj_raw, vj_raw, de_params, ae_params, z_enc, v_enc, x_dec, i_dec,
                    chunk_rows: int = 0):
    """`psnode_forward_encoded`: raw (T,B,<=8) input series in, decoded (T,B,x_dim) trajectories out; the
    encoders run inside the hoisted projection GEMMs, the integration in time chunks, the decoders before the store -- no
    (T,B,H) tensor exists.  `x0` / `a0` are the LATENT initial state (B,H) and all_initial (B,S).  Forward / evaluation only."""
    L = N.lib()
    _require_cuda_f32("t", t)
    dev = t.device
    T, B = t.shape[0], t.shape[1]
    dae = cfg.kind == N.DAE
    keep: list = []
    with torch.cuda.device(dev):
        p = N.Problem()
        p.kind, p.method, p.impl = cfg.kind, cfg.method, N.IMPL_LAYER
        p.B, p.T = B, T
        p.X, p.Z, p.V, p.I = cfg.X, cfg.Z, cfg.V, cfg.I
        ts = _series(t, "t")
        _set_series(p.t, ts)
        x0 = _rows(x0, "latent initial state")
        a0 = _rows(a0, "all_initial")
        keep.extend((ts, x0, a0))
        if dae:
            p.x_init, p.x_init_sb = x0.data_ptr(), x0.stride(0)
        else:
            p.x.p, p.x.st, p.x.sb = x0.data_ptr(), 0, x0.stride(0)
        p.a0, p.a0_sb = a0.data_ptr(), a0.stride(0)
        c = N.Codec()
        zr = _series(z_raw, "z_raw")
        keep.append(zr)
        _set_series(c.z_raw, zr)
        c.ZR = zr.shape[2]
        if dae:
            vr = _series(v_raw, "v_raw")
            keep.append(vr)
            _set_series(c.v_raw, vr)
            c.VR = vr.shape[2]
        p.E = 0
        if event_t is not None:
            _require_cuda_f32("event_t", event_t)
            E = event_t.shape[1]
            ev0, t00 = event_t[0].reshape(E), ts[:, 0, 0]
            if cfg.event_ref is not None:
                t00, ev0 = cfg.event_ref
            idx = torch.empty(max(T - 1, 1), dtype=torch.int32, device=dev)
            err = torch.zeros(1, dtype=torch.int32, device=dev)
            keep.extend((ev0, t00, idx, err))
            N.check(L.psnode_event_table(t00.data_ptr(), t00.stride(0), T, ev0.data_ptr(), ev0.stride(0), E, idx.data_ptr(), err.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "psnode_event_table")
            if cfg.check_events and int(err.item()) != 0:
                raise RuntimeError("more than one event matches the same grid time (the reference raises here too)")
            p.event_idx, p.E = idx.data_ptr(), E
            zj = zj_raw if zj_raw.stride(-1) == 1 or zj_raw.shape[-1] == 1 else zj_raw.contiguous()
            _require_cuda_f32("z_jump", zj)
            keep.append(zj)
            c.zj_raw, c.zjr_sb, c.zjr_se = zj.data_ptr(), zj.stride(0), zj.stride(1)
            if dae:
                vj = vj_raw if vj_raw.stride(-1) == 1 or vj_raw.shape[-1] == 1 else vj_raw.contiguous()
                _require_cuda_f32("v_jump", vj)
                keep.append(vj)
                c.vj_raw, c.vjr_sb, c.vjr_se = vj.data_ptr(), vj.stride(0), vj.stride(1)
        _fill_mlp(p.de, de_params, keep)
        _fill_mlp(c.z_enc, z_enc, keep)
        _fill_mlp(c.x_dec, x_dec, keep)
        c.XR = x_dec[2].shape[0]
        x_out = torch.empty((T, B, c.XR), dtype=torch.float32, device=dev)
        _set_series(c.x_out, x_out)
        i_out = None
        if dae:
            _fill_mlp(p.ae, ae_params, keep)
            _fill_mlp(c.v_enc, v_enc, keep)
            _fill_mlp(c.i_dec, i_dec, keep)
            c.IR = i_dec[2].shape[0]
            i_out = torch.empty((T, B, c.IR), dtype=torch.float32, device=dev)
            _set_series(c.i_out, i_out)
        c.chunk_rows = int(chunk_rows)
        nbytes = L.psnode_forward_encoded_workspace(C.byref(p), C.byref(c))
        if nbytes <= 0:
            raise RuntimeError("psnode_forward_encoded: unsupported problem (latent width must be 128 or 256, 2-layer nets, raw inputs <= 8 "
                               "wide, decoded outputs <= 128 wide)")
        ws = _workspace(dev, nbytes)
        N.check(L.psnode_forward_encoded(C.byref(p), C.byref(c), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream),
                "psnode_forward_encoded")
    return x_out, i_out


# ------------------------------------------------------------------------------------------------ host-buffer entry
def _host_series(ten: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if ten is None or ten.shape[-1] == 0:
        return None
    if ten.is_cuda or ten.dtype != torch.float32:
        raise TypeError(f"forward_host: `{name}` must be a CPU float32 tensor")
    if ten.dim() == 3 and ten.stride(2) != 1 and ten.shape[2] != 1:
        ten = ten.contiguous()
    return ten


def forward_host(cfg: Config, tens: Sequence[Optional[torch.Tensor]], out_x: Optional[torch.Tensor] = None,
                 out_i: Optional[torch.Tensor] = None, device: Optional[torch.device] = None):
    """`psnode_forward_host`: every tensor (series, all_initial, jumps, weights, outputs) lives in HOST memory.

    Pinned tensors (`.pin_memory()`) of >= 1 MB are read / written in place by the kernel over PCIe (zero copy, fused with
    the integration); pageable ones are staged with cudaMemcpyAsync.  Returns (x_sol, i_sol, h2d_bytes, d2h_bytes) with
    time-major (T,B,.) CPU outputs (pinned unless the caller passed its own)."""
    L = N.lib()
    t = _host_series(tens[_T], "t")
    T, B = t.shape[0], t.shape[1]
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    keep: list = []
    host = [t] + [_host_series(tens[k], n) for k, n in ((_X, "x"), (_Zs, "z"), (_Vs, "v"), (_Is, "i"))]
    x_init = tens[_XINIT]
    a0 = tens[_A0]
    for name, ten in (("x_init", x_init), ("all_initial", a0)):
        if ten is not None and (ten.is_cuda or ten.dtype != torch.float32):
            raise TypeError(f"forward_host: `{name}` must be a CPU float32 tensor")
    if out_x is None:
        out_x = torch.empty((T, B, cfg.X), dtype=torch.float32).pin_memory()
    if cfg.kind == N.DAE and out_i is None:
        out_i = torch.empty((T, B, cfg.I), dtype=torch.float32).pin_memory()
    p = N.Problem()
    p.kind, p.method, p.impl = cfg.kind, cfg.method, cfg.impl
    p.B, p.T = B, T
    p.X, p.Z, p.V, p.I = cfg.X, cfg.Z, cfg.V, cfg.I
    p.teacher_x, p.teacher_i = int(cfg.teacher_x), int(cfg.teacher_i)
    _set_series(p.t, host[0])
    _set_series(p.x, host[1] if (cfg.kind == N.ODE or cfg.teacher_x) else None)
    _set_series(p.z, host[2])
    _set_series(p.v, host[3])
    _set_series(p.i, host[4] if cfg.teacher_i else None)
    if x_init is not None and x_init.shape[-1] != 0:
        x_init = x_init if x_init.stride(-1) == 1 else x_init.contiguous()
        p.x_init, p.x_init_sb = x_init.data_ptr(), x_init.stride(0)
    a0 = a0 if a0.stride(-1) == 1 else a0.contiguous()
    p.a0, p.a0_sb = a0.data_ptr(), a0.stride(0)
    keep.extend(host + [x_init, a0])
    p.E = 0
    if cfg.has_event:
        ev_t, zj, vj = tens[_EVT], tens[_ZJ], tens[_VJ]
        E = ev_t.shape[1]
        t00, ev0 = (cfg.event_ref if cfg.event_ref is not None else (t[:, 0, 0], ev_t[0].reshape(E)))
        hit = t00[:-1].reshape(-1, 1) == ev0.reshape(1, -1)                       # (T-1, E), exact float equality
        if cfg.check_events and bool((hit.sum(dim=1) > 1).any()):
            raise RuntimeError("more than one event matches the same grid time (the reference raises here too)")
        idx = torch.where(hit.any(dim=1), hit.float().argmax(dim=1), torch.full((max(T - 1, 0),), -1)).to(torch.int32).contiguous()
        if idx.numel() == 0:
            idx = torch.full((1,), -1, dtype=torch.int32)
        keep.append(idx)
        p.event_idx, p.E = idx.data_ptr(), E
        if cfg.Z > 0:
            zj = zj if zj.stride(-1) == 1 else zj.contiguous()
            keep.append(zj)
            p.z_jump, p.zj_sb, p.zj_se = zj.data_ptr(), zj.stride(0), zj.stride(1)
        if cfg.kind == N.DAE and cfg.V > 0:
            vj = vj if vj.stride(-1) == 1 else vj.contiguous()
            keep.append(vj)
            p.v_jump, p.vj_sb, p.vj_se = vj.data_ptr(), vj.stride(0), vj.stride(1)
    params = tens[_NFIXED:]

    def fill(dst, plist):
        dst.n_layers = len(plist) // 2
        for l in range(dst.n_layers):
            W, b = plist[2 * l].detach(), plist[2 * l + 1].detach()
            if W.is_cuda:
                raise TypeError("forward_host: weights must be CPU tensors")
            W, b = W.contiguous(), b.contiguous()
            keep.extend((W, b))
            dst.in_dim[l], dst.out_dim[l] = W.shape[1], W.shape[0]
            dst.W[l], dst.b[l] = W.data_ptr(), b.data_ptr()
    fill(p.de, params[:2 * cfg.n_de])
    if cfg.kind == N.DAE:
        fill(p.ae, params[2 * cfg.n_de:2 * (cfg.n_de + cfg.n_ae)])
    _set_series(p.x_sol, out_x)
    _set_series(p.i_sol, out_i)
    up, down = C.c_int64(0), C.c_int64(0)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(L.psnode_forward_host(C.byref(p), stream, C.byref(up), C.byref(down)), "psnode_forward_host")
    return out_x, out_i, int(up.value), int(down.value)
