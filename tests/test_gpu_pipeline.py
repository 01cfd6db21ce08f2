"""DeviceBatchPipeline on the GPU: pinned arena, host gather into pinned staging, async H2D on a side stream, slots recycled
behind the consumer's work -- the batches must be exactly the dataset rows even when the consumer keeps the GPU busy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_batches_are_exact_under_load(tmp_path):
    from py_psnode_b200 import DAE_Curves_Sample
    from py_psnode_b200.pipeline import DeviceBatchPipeline
    from test_pipeline_cpu import _write_npz
    p = tmp_path / "set.npz"
    _write_npz(p, n=203, T=40, dae=True)
    ds = DAE_Curves_Sample(str(p), device="cuda:0")
    pipe = DeviceBatchPipeline(ds, batch_size=16, device="cuda:0", shuffle=False)
    assert all(h.is_pinned() for h in pipe.host)
    busy = torch.randn(2048, 2048, device="cuda:0")
    for epoch in range(2):
        k = 0
        for batch in pipe:
            nb = batch[0].shape[0]
            held = [b.clone() for b in batch]                  # consumer work on the batch ...
            for _ in range(3):
                busy = busy @ busy * 1e-3                      # ... and unrelated GPU load while the next batch is staged
            for name, dv in zip(pipe.fields, held):
                want = getattr(ds, name)[k:k + nb]
                assert dv.is_cuda and torch.equal(dv.cpu(), want), (epoch, k, name)
            k += nb
        assert k == 203
