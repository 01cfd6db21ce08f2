"""DeviceBatchPipeline on the CPU device: it must hand the training loop exactly what the reference's
`DataLoader(dataset, batch_size)` + per-field `.to(device)` hands it (neural_00_ODE_01_no_encode.py:326-347) -- same tuple order,
same shapes, every sample exactly once per epoch."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader


def _write_npz(path, n, T, dae):
    rng = np.random.default_rng(5)
    f = dict(t=np.tile((np.arange(T, dtype=np.float32) * 0.01).reshape(1, T, 1), (n, 1, 1)),
             x=rng.standard_normal((n, T, 3)).astype(np.float32), z=rng.standard_normal((n, T, 2)).astype(np.float32),
             event_t=np.full((n, 1, 1), 0.05, dtype=np.float32), z_jump=rng.standard_normal((n, 1, 2)).astype(np.float32),
             mask=np.ones((n, T, 1), dtype=np.float32), name=np.array("synthetic"))
    if dae:
        f.update(v=rng.standard_normal((n, T, 2)).astype(np.float32), i=rng.standard_normal((n, T, 1)).astype(np.float32),
                 v_jump=rng.standard_normal((n, 1, 2)).astype(np.float32))
    np.savez(path, **f)


@pytest.mark.parametrize("dae", [False, True])
def test_pipeline_matches_dataloader_order_and_shapes(tmp_path, dae):
    from py_psnode_b200 import DAE_Curves_Sample, ODE_Curves_Sample
    from py_psnode_b200.pipeline import DeviceBatchPipeline
    p = tmp_path / "set.npz"
    _write_npz(p, n=23, T=9, dae=dae)
    ds = (DAE_Curves_Sample if dae else ODE_Curves_Sample)(str(p), device="cpu", cut_length=7)
    pipe = DeviceBatchPipeline(ds, batch_size=5, device="cpu", shuffle=False)
    ref = list(DataLoader(ds, batch_size=5, shuffle=False))
    got = list(pipe)
    assert len(got) == len(ref) == len(pipe) == 5
    for a, b in zip(got, ref):
        assert len(a) == len(b) == (9 if dae else 6)
        for u, v in zip(a, b):
            assert u.shape == v.shape and torch.equal(u, v)


def test_pipeline_shuffle_covers_every_sample_once_and_drop_last(tmp_path):
    from py_psnode_b200 import ODE_Curves_Sample
    from py_psnode_b200.pipeline import DeviceBatchPipeline
    p = tmp_path / "set.npz"
    _write_npz(p, n=23, T=6, dae=False)
    ds = ODE_Curves_Sample(str(p), device="cpu")
    pipe = DeviceBatchPipeline(ds, batch_size=4, device="cpu", shuffle=True, seed=3)
    key = lambda x: tuple(np.round(x[:, 0, 0].numpy(), 6))
    seen = []
    for batch in pipe:
        seen.extend(key(batch[1]))
    assert sorted(seen) == sorted(key(ds.x))
    first_epoch = seen
    second = [v for batch in pipe for v in key(batch[1])]
    assert sorted(second) == sorted(first_epoch) and second != first_epoch        # reshuffled every epoch
    pipe2 = DeviceBatchPipeline(ds, batch_size=4, device="cpu", shuffle=False, drop_last=True)
    assert len(pipe2) == 5 and sum(b[0].shape[0] for b in pipe2) == 20
