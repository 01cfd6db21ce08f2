"""Recognise the reference's RHS / algebraic modules and event callbacks so their arithmetic can run in the kernel.

The reference's solver calls duck-typed Python objects every step:
    x_func(t0=, xt=, zt=[, vt=, it=], all_initial=)        my_fixed_grid.py:16-17
    i_func(xt=, zt=, vt=, all_initial=)                    my_solvers.py:95,110,121
    event_fn(t0) / jump_change_fn(t0, z0[, v0])            my_solvers.py:70-71, 108-109
The kernel cannot call Python, so the host side extracts *what those objects compute*:
  * an `nn.Sequential` of alternating `nn.Linear` / `nn.ELU(alpha=1)` stored as `.x_dot` (DE_Func) or
    `.i_calculator` (AE_Func) -- every script-local class of the reference has this shape
    (neural_00_ODE_01_no_encode.py:61-64, neural_00_ODE_02_direct_encode.py:52-53,
     neural_01_DAE_01_no_encode.py:64-67/77-80, neural_01_DAE_02_direct_encode.py:73-80/90-97);
  * that the module's forward really is  mlp(cat(a0, s - a0, s))  /  mlp(cat(a0, x, z, v)): checked ONCE per
    module instance by evaluating the module itself on a tiny random probe (`verify_*`), so a look-alike module
    with different semantics is rejected instead of silently mis-integrated;
  * the event tensors held by the bound `ODE_Event` / `DAE_Event` object.
Anything else raises UnsupportedModuleError -- loudly; there is no silent fallback.
"""
import weakref
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn


class UnsupportedModuleError(TypeError):
    """x_func / i_func / event callbacks are not of the shape the fused integrator understands."""


def linear_chain(module: nn.Module, preferred_attr: str) -> List[nn.Linear]:
    """Return the Linear layers of the module's single Linear/ELU `nn.Sequential`."""
    seq = getattr(module, preferred_attr, None)
    if not isinstance(seq, nn.Sequential):
        seqs = [m for m in module.children() if isinstance(m, nn.Sequential)]
        others = [m for m in module.children() if not isinstance(m, nn.Sequential)]
        if len(seqs) != 1 or others:
            raise UnsupportedModuleError(
                f"{type(module).__name__}: expected one nn.Sequential of Linear/ELU layers (attribute "
                f"`{preferred_attr}`), found children {[type(m).__name__ for m in module.children()]}")
        seq = seqs[0]
    mods = list(seq)
    if len(mods) % 2 != 1:
        raise UnsupportedModuleError(f"{type(module).__name__}: Sequential must be Linear,(ELU,Linear)*")
    layers = []
    for k, m in enumerate(mods):
        if k % 2 == 0:
            if not isinstance(m, nn.Linear) or m.bias is None:
                raise UnsupportedModuleError(f"{type(module).__name__}: layer {k} is {type(m).__name__}, need Linear with bias")
            layers.append(m)
        else:
            if not isinstance(m, nn.ELU) or float(m.alpha) != 1.0:
                raise UnsupportedModuleError(f"{type(module).__name__}: layer {k} is {type(m).__name__}, need ELU(alpha=1)")
    for a, b in zip(layers[:-1], layers[1:]):
        if a.out_features != b.in_features:
            raise UnsupportedModuleError(f"{type(module).__name__}: inconsistent layer widths")
    if len(layers) > 8:
        raise UnsupportedModuleError(f"{type(module).__name__}: more than 8 Linear layers")
    return layers


def match_init(module: nn.Module, Z: int, V: int, I: int) -> List[nn.Linear]:
    """Linear layers of an `Init_Func` (neural_01_DAE_01_no_encode.py:50-58): forward(z0, v0, i0) = mlp(cat(z0, v0, i0)), verified once per
    module instance on a random probe."""
    layers = linear_chain(module, "init_fun")
    if layers[0].in_features != Z + V + I:
        raise UnsupportedModuleError(f"{type(module).__name__}: first Linear takes {layers[0].in_features} inputs, expected Z+V+I = {Z + V + I}")

    def probe():
        p = layers[0].weight
        g = torch.Generator(device="cpu").manual_seed(0)
        mk = lambda w: torch.randn(3, w, generator=g).to(device=p.device, dtype=p.dtype)
        z0, v0, i0 = mk(Z), mk(V), mk(I)
        return module(z0, v0, i0), _mlp(layers, torch.cat((z0, v0, i0), dim=-1))
    _probe_ok(module, ("init", Z, V, I), probe)
    return layers


def codec_chain(seq: nn.Module, what: str) -> List[nn.Linear]:
    """The two Linear layers of an encoder / decoder of the `*_02_direct_encode` models: nn.Sequential(Linear, ELU, Linear)
    (neural_00_ODE_02_direct_encode.py:63-68, neural_01_DAE_02_direct_encode.py:107-119)."""
    if not isinstance(seq, nn.Sequential):
        raise UnsupportedModuleError(f"{what}: expected nn.Sequential(Linear, ELU, Linear), got {type(seq).__name__}")
    holder = nn.Module()
    holder.seq = seq
    layers = linear_chain(holder, "seq")
    if len(layers) != 2:
        raise UnsupportedModuleError(f"{what}: expected exactly two Linear layers, found {len(layers)}")
    return layers


def _mlp(layers: Sequence[nn.Linear], u: torch.Tensor) -> torch.Tensor:
    for k, lin in enumerate(layers):
        u = torch.nn.functional.linear(u, lin.weight, lin.bias)
        if k != len(layers) - 1:
            u = torch.nn.functional.elu(u)
    return u


_verified = weakref.WeakKeyDictionary()   # module -> set of signatures already probed


def _probe_ok(module, sig, fn) -> None:
    seen = _verified.setdefault(module, set())
    if sig in seen:
        return
    with torch.no_grad():
        got, want = fn()
    if got.shape != want.shape or not torch.allclose(got, want, rtol=1e-4, atol=1e-5):
        raise UnsupportedModuleError(
            f"{type(module).__name__}.forward does not compute the Linear/ELU chain on the expected concatenation "
            f"(signature {sig}); the fused integrator cannot represent it")
    seen.add(sig)


def match_de(module: nn.Module, X: int, Z: int, V: int = 0, I: int = 0, dae: bool = False) -> List[nn.Linear]:
    """DE_Func: dx/dt = mlp(cat(a0, s - a0, s)), s = cat(x, z[, v, i])."""
    layers = linear_chain(module, "x_dot")
    S = X + Z + V + I
    if layers[0].in_features != 3 * S or layers[-1].out_features != X:
        raise UnsupportedModuleError(
            f"{type(module).__name__}: first/last widths {layers[0].in_features}/{layers[-1].out_features} "
            f"do not match 3*(X+Z+V+I)={3 * S} / X={X}")
    p = layers[0].weight

    def probe():
        g = torch.Generator(device="cpu").manual_seed(1234)
        mk = lambda w: torch.randn(3, w, generator=g).to(device=p.device, dtype=p.dtype)
        a0, x, z = mk(S), mk(X), mk(Z)
        t0 = torch.zeros(3, 1, device=p.device, dtype=p.dtype)
        if dae:
            v, i = mk(V), mk(I)
            got = module(t0=t0, xt=x, zt=z, vt=v, it=i, all_initial=a0)
            s = torch.cat((x, z, v, i), dim=-1)
        else:
            got = module(t0=t0, xt=x, zt=z, all_initial=a0)
            s = torch.cat((x, z), dim=-1)
        return got, _mlp(layers, torch.cat((a0, s - a0, s), dim=-1))

    _probe_ok(module, ("de", X, Z, V, I, dae), probe)
    return layers


def match_ae(module: nn.Module, X: int, Z: int, V: int, I: int) -> List[nn.Linear]:
    """AE_Func: i = mlp(cat(a0, x, z, v))."""
    layers = linear_chain(module, "i_calculator")
    S = X + Z + V + I
    if layers[0].in_features != S + X + Z + V or layers[-1].out_features != I:
        raise UnsupportedModuleError(
            f"{type(module).__name__}: first/last widths {layers[0].in_features}/{layers[-1].out_features} "
            f"do not match S+X+Z+V={S + X + Z + V} / I={I}")
    p = layers[0].weight

    def probe():
        g = torch.Generator(device="cpu").manual_seed(4321)
        mk = lambda w: torch.randn(3, w, generator=g).to(device=p.device, dtype=p.dtype)
        a0, x, z, v = mk(S), mk(X), mk(Z), mk(V)
        got = module(xt=x, zt=z, vt=v, all_initial=a0)
        return got, _mlp(layers, torch.cat((a0, x, z, v), dim=-1))

    _probe_ok(module, ("ae", X, Z, V, I), probe)
    return layers


def match_event(event_fn, jump_change_fn, dae: bool) -> Optional[Tuple[torch.Tensor, ...]]:
    """Return (event_t, z_jump[, v_jump]) held by the bound event object, or None when no event can ever fire."""
    if event_fn is None:
        return None
    owner = getattr(event_fn, "__self__", None)
    name = getattr(event_fn, "__name__", "")
    if owner is None or name != "event_fn" or not hasattr(owner, "event_t") or not hasattr(owner, "z_jump"):
        raise UnsupportedModuleError(
            "event_fn must be the bound `event_fn` of an ODE_Event / DAE_Event object (the kernel folds its table "
            "lookup into input staging and cannot call an arbitrary Python predicate every step)")
    if jump_change_fn is None or getattr(jump_change_fn, "__self__", None) is not owner \
            or getattr(jump_change_fn, "__name__", "") != "jump_change_fn":
        raise UnsupportedModuleError("jump_change_fn must be the bound `jump_change_fn` of the same event object as event_fn")
    if owner.event_t is None:
        return None            # set_event() never called: event_fn() is constantly False (neural_base.py:53)
    if owner.event_t.dim() >= 2 and owner.event_t.shape[1] == 0:
        return None            # a dataset without events: `t0[0] in event_t[0]` is False for an empty table (neural_base.py:54)
    if dae:
        if not hasattr(owner, "v_jump"):
            raise UnsupportedModuleError("DAE integration needs a DAE_Event (z_jump and v_jump)")
        return owner.event_t, owner.z_jump, owner.v_jump
    return owner.event_t, owner.z_jump


def event_reference(event_fn):
    """(t_row, ev_row) pinned on the event object by `parallel.pin_event_reference` (batch-sharded runs: every rank must
    test the GLOBAL sample 0, because the reference's predicate looks at sample 0 of the whole batch, neural_base.py:54),
    or None -> use sample 0 of the tensors passed to the call."""
    owner = getattr(event_fn, "__self__", None)
    return getattr(owner, "_psn_event_ref", None) if owner is not None else None
