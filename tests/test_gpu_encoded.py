"""Encoder / decoder fusion (C ABI `psnode_forward_encoded`, `integrate_ODE_encoded` / `integrate_DAE_encoded`):
the `ODE_Model.forward` / `DAE_Model.forward` pipelines of the `*_02_direct_encode` scripts (neural_00_ODE_02_direct_encode.py:75-89,
neural_01_DAE_02_direct_encode.py:126-153) in one call from the RAW input series -- encoders inside the hoisted projection GEMMs
(generated shared-memory operand), integration in time chunks, decoders before the store -- against
  * goldens the UNMODIFIED reference models produced on the CPU (the scripts' own classes, imported from oracle/_ref, hold the weights),
  * the unfused pipeline of this repo (torch encoders -> integrate_* -> torch decoders) at ragged batches, several time chunks,
    events on and off chunk boundaries, batch-major storage, H = 128 and 256."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import ATOL, GOLDEN_DIR, RTOL, tol_report

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
REF_STUBS = os.path.join(ROOT, "oracle", "_ref", "stubs")
DEV = "cuda:0"


def _script(modname):
    if not os.path.isfile(os.path.join(REF_SRC, modname + ".py")):
        pytest.skip("oracle/_ref is absent (python oracle/make_ref.py vendors the reference in the build container)")
    import neural_dae                                   # the repo's shim must be the one the scripts bind to
    assert os.path.abspath(neural_dae.__file__).startswith(ROOT) and "_ref" not in neural_dae.__file__
    for p in (REF_SRC, REF_STUBS):
        if p not in sys.path:
            sys.path.append(p)
    return importlib.import_module(modname)


def _load(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    kw = {str(k): int(v) for k, v in zip(g["kw_keys"], g["kw_vals"])}
    state = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")}
    d = {k[3:]: torch.from_numpy(v).to(DEV) for k, v in g.items() if k.startswith("in_")}
    return g, kw, state, d


@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_encoded_ode_model_vs_reference_golden(native_lib, solver):
    import neural_dae
    from py_psnode_b200 import _native
    mod = _script("neural_00_ODE_02_direct_encode")
    g, kw, state, d = _load("script_ode02")
    model = mod.ODE_Model(**kw)
    model.load_state_dict(state)
    model = model.to(DEV).eval()
    S = {"euler": neural_dae.Euler, "rk4": neural_dae.RK4}[solver]
    with torch.no_grad():
        x0 = model.x_encoder(d["x"][:, 0])
        a0 = torch.cat((x0, model.z_encoder(d["z"][:, 0])), dim=-1)
        got = S().integrate_ODE_encoded(x_func=model.de_func, t=d["t"].permute(1, 0, 2), x0=x0, z=d["z"].permute(1, 0, 2), all_initial=a0,
                                        z_encoder=model.z_encoder, x_decoder=model.x_decoder, event_t=d["event_t"], z_jump=d["z_jump"],
                                        chunk_rows=4)
    assert _native.last_kernel().startswith("psn_lg_gemm_kernel<dec2>"), _native.last_kernel()
    got = got.permute(1, 0, 2).cpu()
    want, want64 = torch.from_numpy(g[f"{solver}_pred0"]), torch.from_numpy(g[f"{solver}_pred64_0"])
    assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), tol_report(got, want, want64)


@pytest.mark.parametrize("solver", ["euler", "rk4"])
def test_encoded_dae_model_vs_reference_golden(native_lib, solver):
    import neural_dae
    mod = _script("neural_01_DAE_02_direct_encode")
    g, kw, state, d = _load("script_dae02_h128")
    model = mod.DAE_Model(**kw)
    model.load_state_dict(state)
    model = model.to(DEV).eval()
    S = {"euler": neural_dae.Euler, "rk4": neural_dae.RK4}[solver]
    with torch.no_grad():
        z0, v0, i0 = d["z"][:, 0], d["v"][:, 0], d["i"][:, 0]
        x0 = model.init_func(z0, v0, i0)
        Xh0 = model.x_encoder(x0)
        a0 = torch.cat((Xh0, model.z_encoder(z0), model.v_encoder(v0), model.i_encoder(i0)), dim=-1)
        xp, ip = S().integrate_DAE_encoded(x_init=Xh0, x_func=model.de_func, i_func=model.ae_func, t=d["t"].permute(1, 0, 2),
                                           z=d["z"].permute(1, 0, 2), v=d["v"].permute(1, 0, 2), all_initial=a0, z_encoder=model.z_encoder,
                                           v_encoder=model.v_encoder, x_decoder=model.x_decoder, i_decoder=model.i_decoder,
                                           event_t=d["event_t"], z_jump=d["z_jump"], v_jump=d["v_jump"], chunk_rows=3)
        xp[0] = x0                                      # DAE_Model.forward: x_pred[0] = x0 (neural_01_DAE_02_direct_encode.py:150)
    for k, got in enumerate((xp, ip)):
        got = got.permute(1, 0, 2).cpu()
        want, want64 = torch.from_numpy(g[f"{solver}_pred{k}"]), torch.from_numpy(g[f"{solver}_pred64_{k}"])
        assert torch.allclose(got, want, rtol=RTOL, atol=ATOL), f"output {k}: " + tol_report(got, want, want64)


def _codec(i, h, o):
    return nn.Sequential(nn.Linear(i, h), nn.ELU(), nn.Linear(h, o))


@pytest.mark.parametrize("H,solver,B,N,chunk,events", [(256, "rk4", 150, 30, 7, 2), (128, "midpoint", 40, 12, 5, 1), (128, "euler", 300, 9, 0, 0)])
def test_encoded_dae_vs_unfused_pipeline(native_lib, H, solver, B, N, chunk, events):
    from py_psnode_b200 import AE_Func, DAE_Event, DE_Func, Euler, Midpoint, RK4, _native
    torch.manual_seed(90 + H + N)
    T = N + 1
    XR, ZR, VR, IR = 5, 1, 2, 3
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, v_dim=H, i_dim=H, depth=2).to(DEV)
    ae = AE_Func(x_dim=H, v_dim=H, i_dim=H, hidden_dim=H, z_dim=H, depth=2).to(DEV)
    z_enc, v_enc, x_dec, i_dec = _codec(ZR, H, H).to(DEV), _codec(VR, H, H).to(DEV), _codec(H, H, XR).to(DEV), _codec(H, H, IR).to(DEV)
    t = (torch.arange(T, dtype=torch.float32, device=DEV) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    # raw series in the scripts' batch-major storage, passed as time-major views
    z = (torch.randn(B, T, ZR, device=DEV)).permute(1, 0, 2)
    v = (torch.randn(B, T, VR, device=DEV)).permute(1, 0, 2)
    x_init = torch.randn(B, H, device=DEV) * 0.05
    i0 = torch.randn(B, H, device=DEV) * 0.05
    S = {"euler": Euler, "midpoint": Midpoint, "rk4": RK4}[solver]
    event_t = zj = vj = None
    kw = {}
    with torch.no_grad():
        Zh, Vh = z_enc(z), v_enc(v)
        a0 = torch.cat((x_init, Zh[0], Vh[0], i0), dim=-1)
        if events:
            steps = [chunk if chunk else N // 3, (2 * N) // 3][:events]      # the first event sits on a chunk boundary
            event_t = torch.stack([t[s, :, 0] for s in steps], dim=1).view(B, events, 1).clone()
            zj, vj = torch.randn(B, events, ZR, device=DEV), torch.randn(B, events, VR, device=DEV)
            ev = DAE_Event()
            ev.set_event(t=event_t, z=z_enc(zj), v=v_enc(vj))
            kw = dict(event_fn=ev.event_fn, jump_change_fn=ev.jump_change_fn)
        xs, is_ = S(impl="layer").integrate_DAE(x_init=x_init, x_func=de, i_func=ae, t=t, x=x_init.unsqueeze(0).expand(T, B, H), z=Zh, v=Vh,
                                               i=i0.unsqueeze(0).expand(T, B, H), all_initial=a0, **kw)
        want_x, want_i = x_dec(xs), i_dec(is_)
        got_x, got_i = S().integrate_DAE_encoded(x_init=x_init, x_func=de, i_func=ae, t=t, z=z, v=v, all_initial=a0, z_encoder=z_enc,
                                                 v_encoder=v_enc, x_decoder=x_dec, i_decoder=i_dec, event_t=event_t, z_jump=zj, v_jump=vj,
                                                 chunk_rows=chunk)
        again_x, _ = S().integrate_DAE_encoded(x_init=x_init, x_func=de, i_func=ae, t=t, z=z, v=v, all_initial=a0, z_encoder=z_enc,
                                               v_encoder=v_enc, x_decoder=x_dec, i_decoder=i_dec, event_t=event_t, z_jump=zj, v_jump=vj,
                                               chunk_rows=chunk)
    assert _native.last_kernel().startswith("psn_lg_gemm_kernel<dec2>"), _native.last_kernel()
    assert got_x.shape == (T, B, XR) and got_i.shape == (T, B, IR)
    assert torch.allclose(got_x, want_x, rtol=RTOL, atol=2 * ATOL), "x: " + tol_report(got_x.cpu(), want_x.cpu())
    assert torch.allclose(got_i, want_i, rtol=RTOL, atol=2 * ATOL), "i: " + tol_report(got_i.cpu(), want_i.cpu())
    assert torch.equal(again_x, got_x), "deterministic"


@pytest.mark.parametrize("H,B", [(256, 130), (128, 130), (128, 1000)])
def test_encoded_ode_vs_unfused_pipeline(native_lib, H, B):
    """H = 256: per-layer GEMM kernel; H = 128: the wide kernels (projection tiles generated from the raw series, TMEM-resident time loop)."""
    from py_psnode_b200 import DE_Func, ODE_Event, RK4
    torch.manual_seed(95)
    N, XR, ZR = 20, 8, 2
    T = N + 1
    de = DE_Func(x_dim=H, z_dim=H, hidden_dim=H, depth=2).to(DEV)
    z_enc, x_dec = _codec(ZR, H, H).to(DEV), _codec(H, H, XR).to(DEV)
    t = (torch.arange(T, dtype=torch.float32, device=DEV) * 0.01).view(T, 1, 1).repeat(1, B, 1)
    z = torch.randn(T, B, ZR, device=DEV)
    x0 = torch.randn(B, H, device=DEV) * 0.05
    event_t = t[6].view(B, 1, 1).clone()
    zj = torch.randn(B, 1, ZR, device=DEV)
    with torch.no_grad():
        Zh = z_enc(z)
        a0 = torch.cat((x0, Zh[0]), dim=-1)
        ev = ODE_Event()
        ev.set_event(t=event_t, z=z_enc(zj))
        xs = RK4().integrate_ODE(x_func=de, t=t, x=x0.unsqueeze(0).expand(T, B, H), z=Zh, all_initial=a0, event_fn=ev.event_fn,
                                 jump_change_fn=ev.jump_change_fn)
        want = x_dec(xs)
        got = RK4().integrate_ODE_encoded(x_func=de, t=t, x0=x0, z=z, all_initial=a0, z_encoder=z_enc, x_decoder=x_dec, event_t=event_t,
                                          z_jump=zj, chunk_rows=6)
    assert torch.allclose(got, want, rtol=RTOL, atol=2 * ATOL), tol_report(got.cpu(), want.cpu())
