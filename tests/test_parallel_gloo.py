"""Host-side logic of the batch-sharded data-parallel path (py_psnode_b200/parallel.py) on CPU with the gloo backend,
world_size 2: contiguous sharding, the pinned global event reference, and the single flat gradient all-reduce with the
batch-global mask normalisation must reproduce the single-process gradient of the reference's masked-MSE loss.
The integration itself has no CPU path, so the oracle (test infrastructure) stands in for the forward/backward here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from py_psnode_b200 import parallel


def test_shard_bounds_cover_batch():
    for n in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _problem(seed=0, B=6, N=12, X=4, Z=2, H=16):
    g = torch.Generator().manual_seed(seed)
    T = N + 1
    t = (torch.arange(T, dtype=torch.float32) * 0.01).view(1, T, 1).repeat(B, 1, 1)
    x = torch.randn(B, T, X, generator=g) * 0.1
    z = torch.randn(B, T, Z, generator=g) * 0.1
    # per-sample event times DIFFER: the reference still fires everything on sample 0's time (neural_base.py:54)
    ev_idx = torch.tensor([N // 2, 3, 4, 5, 6, 7][:B])
    event_t = t[torch.arange(B), ev_idx].view(B, 1, 1).clone()
    z_jump = torch.randn(B, 1, Z, generator=g) * 0.1
    mask = (torch.rand(B, T, 1, generator=g) > 0.2).float().expand(B, T, X).contiguous()
    S = X + Z
    dims = [3 * S, H, H, X]
    params = []
    for a, b in zip(dims[:-1], dims[1:]):
        params.append((torch.randn(b, a, generator=g) * 0.2, torch.randn(b, generator=g) * 0.1))
    return dict(t=t, x=x, z=z, event_t=event_t, z_jump=z_jump, mask=mask, params=params)


def _loss_parts(pb, params, lo, hi, t_row, ev_row):
    from oracle import psnode_oracle as O
    t, x, z = (pb[k][lo:hi].permute(1, 0, 2) for k in ("t", "x", "z"))
    a0 = torch.cat((x[0], z[0]), dim=-1)
    # the oracle reads sample 0 of the tensors it is given: hand it the pinned global rows through a 1-sample prefix
    ev = ev_row.view(1, -1, 1).expand(hi - lo, -1, 1)
    tt = t.clone()
    tt[:, 0, 0] = t_row
    sol = O.integrate_ode("rk4", params, tt, x, z, a0, ev, pb["z_jump"][lo:hi])
    return parallel.masked_mse_sum(sol.permute(1, 0, 2), pb["x"][lo:hi], pb["mask"][lo:hi])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        pb = _problem()
        B = pb["t"].shape[0]
        lo, hi = parallel.shard_bounds(B, rank, world)
        leaves = [torch.nn.Parameter(q.clone()) for wb in pb["params"] for q in wb]
        params = [(leaves[2 * k], leaves[2 * k + 1]) for k in range(len(pb["params"]))]

        class Ev:            # stands in for ODE_Event: pin_event_reference only stores attributes
            pass
        ev = Ev()
        t_l, ev_l = parallel.shard_batch([pb["t"], pb["event_t"]], rank, world)
        parallel.pin_event_reference(ev, t_l, ev_l)
        t_row, ev_row = ev._psn_event_ref
        bucket = parallel.GradBucket(leaves, n_extras=2)
        loss = parallel.sharded_training_step(lambda: None, leaves, bucket,
                                              lambda _: _loss_parts(pb, params, lo, hi, t_row, ev_row))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss, t_row=t_row.numpy(), ev_row=ev_row.numpy(),
                 **{f"g{k}": p.grad.numpy() for k, p in enumerate(leaves)})
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gradients_equal_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process reference: whole batch, loss normalised by the global mask sum
    torch.set_num_threads(1)
    pb = _problem()
    leaves = [q.clone().requires_grad_(True) for wb in pb["params"] for q in wb]
    params = [(leaves[2 * k], leaves[2 * k + 1]) for k in range(len(pb["params"]))]
    B = pb["t"].shape[0]
    num, den = _loss_parts(pb, params, 0, B, pb["t"][0, :, 0], pb["event_t"][0, :, 0])
    loss = num / den
    loss.backward()
    r0, r1 = (np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world))
    np.testing.assert_array_equal(r0["t_row"], pb["t"][0, :, 0].numpy())       # every rank tests GLOBAL sample 0
    np.testing.assert_array_equal(r1["t_row"], pb["t"][0, :, 0].numpy())
    np.testing.assert_array_equal(r1["ev_row"], pb["event_t"][0, :, 0].numpy())
    np.testing.assert_allclose(float(r0["loss"]), loss.item(), rtol=1e-5)
    np.testing.assert_allclose(float(r1["loss"]), loss.item(), rtol=1e-5)
    for k, p in enumerate(leaves):
        np.testing.assert_array_equal(r0[f"g{k}"], r1[f"g{k}"])                  # ranks agree bit for bit after the all-reduce
        np.testing.assert_allclose(r0[f"g{k}"], p.grad.numpy(), rtol=2e-4, atol=1e-7)


def test_bucket_single_process_normalises():
    lin = torch.nn.Linear(3, 2)
    (lin(torch.ones(4, 3)).sum() * 2.0).backward()
    want = [p.grad.clone() / 8.0 for p in lin.parameters()]
    bucket = parallel.GradBucket(lin.parameters(), n_extras=1)
    red = bucket.allreduce_(extras=(8.0,), normalise_by_extra=0)
    assert float(red[0]) == 8.0
    for p, w in zip(lin.parameters(), want):
        torch.testing.assert_close(p.grad, w)
    assert bucket.nbytes == (3 * 2 + 2 + 1) * 4
