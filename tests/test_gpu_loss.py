"""Fused masked squared-error loss (csrc/psnode_loss.cu) against the scripts' formula
`torch.sum(Loss_func(x_pred, x, reduction='none') * mask)` (neural_00_ODE_01_no_encode.py:353-354) and its autograd
gradient, on the layouts the scripts produce: time-major solver output, batch-major permuted views, narrow states."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(pred, target, mask, w=None):
    se = torch.nn.functional.mse_loss(pred, target, reduction="none") * mask
    if w is not None:
        se = se * w
    return se.sum()


@pytest.mark.parametrize("layout", ["time_major", "batch_major_view", "offset_slice"])
@pytest.mark.parametrize("X,weighted", [(16, False), (16, True), (6, False), (5, True), (4, False)])
def test_masked_sse_value_and_gradient(native_lib, layout, X, weighted):
    from py_psnode_b200 import _native
    from py_psnode_b200.losses import masked_sse
    torch.manual_seed(7)
    dev = "cuda:0"
    T, B = 37, 301
    if layout == "time_major":
        pred = torch.randn(T, B, X, device=dev)
        target = torch.randn(T, B, X, device=dev)
        mask = (torch.rand(T, B, 1, device=dev) > 0.3).float()
    elif layout == "batch_major_view":           # what the scripts hold: (B,T,.) storage, compared after permute
        pred = torch.randn(B, T, X, device=dev).permute(1, 0, 2)
        target = torch.randn(B, T, X, device=dev).permute(1, 0, 2)
        mask = (torch.rand(B, T, 1, device=dev) > 0.3).float().permute(1, 0, 2)
    else:                                        # rows that are not 16-byte aligned
        pred = torch.randn(T, B, X + 3, device=dev)[..., 1:1 + X]
        target = torch.randn(T, B, X + 1, device=dev)[..., 1:]
        mask = (torch.rand(T, B, 1, device=dev) > 0.3).float()
    w = (torch.rand(X, device=dev) * 9 + 1) if weighted else None
    p1 = pred.detach().clone().requires_grad_(True)
    want = _reference(p1, target, mask, w)
    (want * 0.37).backward()
    p2 = pred.detach().requires_grad_(True)      # keeps the strided layout for the fused op
    n0 = _native.launch_count()
    got = masked_sse(p2, target, mask, w)
    assert _native.launch_count() - n0 == 2      # partial sums + final sum
    (got * 0.37).backward()
    assert torch.allclose(got, want, rtol=2e-6, atol=0), (float(got), float(want))
    assert torch.allclose(p2.grad, p1.grad, rtol=1e-6, atol=1e-7), float((p2.grad - p1.grad).abs().max())
    # deterministic reduction order
    again = masked_sse(pred, target, mask, w)
    assert torch.equal(again, got.detach())


def test_masked_mse_sum_uses_the_fused_kernel_and_matches_the_script_loss(native_lib):
    """parallel.masked_mse_sum on CUDA tensors = numerator / denominator of the ODE script's loss (:353-355)."""
    from py_psnode_b200 import _native, parallel
    torch.manual_seed(9)
    dev = "cuda:0"
    B, T, X = 64, 50, 16
    x = torch.randn(B, T, X, device=dev)
    x_pred = (x + 0.1 * torch.randn(B, T, X, device=dev)).requires_grad_(True)
    mask = (torch.rand(B, T, 1, device=dev) > 0.2).float()
    x_loss = torch.sum(torch.sum(torch.nn.functional.mse_loss(x_pred, x, reduction="none") * mask, dim=1), dim=0) / torch.sum(mask)
    want = torch.sum(x_loss)
    gwant, = torch.autograd.grad(want, x_pred)
    n0 = _native.launch_count()
    num, den = parallel.masked_mse_sum(x_pred, x, mask)
    assert _native.launch_count() - n0 == 2
    got = num / den
    ggot, = torch.autograd.grad(got, x_pred)
    assert torch.allclose(got, want, rtol=2e-6), (float(got), float(want))
    assert torch.allclose(ggot, gwant, rtol=1e-6, atol=1e-9)


def test_masked_sse_rejects_host_tensors(native_lib):
    from py_psnode_b200.losses import masked_sse
    with pytest.raises(TypeError):
        masked_sse(torch.zeros(3, 4, 2), torch.zeros(3, 4, 2), torch.ones(3, 4, 1))
