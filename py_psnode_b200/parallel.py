"""Batch-sharded data parallelism for the integration path: one process per GPU, trajectories are
independent units, so the batch is split contiguously, every rank runs the fused forward / reverse sweep on its shard,
and ONE all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of a single flat fp32 bucket carries the parameter
gradients together with the loss-normalisation scalars.  The reference has no multi-GPU code at all; what has to be
preserved from it is the arithmetic of its single-process training step:

  * the loss is normalised by the BATCH-GLOBAL sum(mask) (neural_00_ODE_01_no_encode.py:353-355,
    neural_01_DAE_01_no_encode.py:414-418): ranks back-propagate the UN-normalised local sum and divide the reduced
    gradients by the reduced mask sum (`GradBucket.allreduce_`, extras);
  * the event predicate looks at sample 0 of the whole batch (neural_base.py:54): rank 0's sample 0 is broadcast and
    pinned on the event object (`pin_event_reference`), otherwise shards could disagree about when events fire.

Nothing here touches the data path: there is no collective inside the integration.
"""
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of a batch of n trajectories owned by `rank`; sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: int, world: int, dim: int = 0):
    """Slice every tensor along its batch dimension (batch-major (B, ...) storage as the reference's DataLoader yields)."""
    out = []
    for ten in tensors:
        if ten is None:
            out.append(None)
            continue
        lo, hi = shard_bounds(ten.shape[dim], rank, world)
        out.append(ten.narrow(dim, lo, hi - lo))
    return out


def pin_event_reference(event, t_batch_major: torch.Tensor, event_t: torch.Tensor, group=None, src: int = 0) -> None:
    """Make every rank evaluate the event predicate on the GLOBAL batch's sample 0.

    t_batch_major: this rank's (B_local, T, 1) time tensor; event_t: this rank's (B_local, E, 1).  Rank `src` (which owns
    global sample 0 under contiguous sharding) broadcasts its rows t[0,:,0] and event_t[0,:,0]; they are stored on the
    event object and used for the per-step event table instead of the local sample 0."""
    t_row = t_batch_major[0, :, 0].detach().clone().contiguous()
    ev_row = event_t[0, :, 0].detach().clone().contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t_row, src=src, group=group)
        dist.broadcast(ev_row, src=src, group=group)
    event._psn_event_ref = (t_row, ev_row)


def clear_event_reference(event) -> None:
    if hasattr(event, "_psn_event_ref"):
        del event._psn_event_ref


class GradBucket:
    """One flat fp32 buffer holding every parameter gradient plus a few extra scalars; a single all-reduce per optimiser
    step (0.05 MB for the cfg2 net, 0.7 MB cfg4, 7.5 MB cfg5 -- latency-, not bandwidth-bound on NVLink)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_extras: int = 1):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.n_extras = n_extras
        self.flat = torch.zeros(sum(self.sizes) + n_extras, dtype=torch.float32, device=dev)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def allreduce_(self, extras: Sequence[float] = (), group=None, normalise_by_extra: Optional[int] = 0) -> torch.Tensor:
        """Pack grads (+extras) -> one SUM all-reduce -> unpack into p.grad.  If `normalise_by_extra` is an index, every
        gradient is divided by that reduced extra (the batch-global mask sum).  Returns the reduced extras tensor."""
        if len(extras) != self.n_extras:
            raise ValueError(f"expected {self.n_extras} extras, got {len(extras)}")
        off = 0
        for p, n in zip(self.params, self.sizes):
            seg = self.flat[off:off + n]
            if p.grad is None:
                seg.zero_()
            else:
                seg.copy_(p.grad.reshape(-1))
            off += n
        for k, e in enumerate(extras):
            if torch.is_tensor(e):
                self.flat[off + k] = e.detach().to(self.flat.dtype).reshape(())
            else:
                self.flat[off + k] = float(e)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        red = self.flat[off:off + self.n_extras].clone()
        scale = None
        if normalise_by_extra is not None:
            scale = 1.0 / red[normalise_by_extra]
        off = 0
        for p, n in zip(self.params, self.sizes):
            g = self.flat[off:off + n].view_as(p)
            if scale is not None:
                g = g * scale
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return red


def masked_mse_sum(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(sum(mse * mask), sum(mask)) of a shard: numerator and denominator of the reference's masked loss
    (neural_00_ODE_01_no_encode.py:353-355) kept apart so the division can use the all-reduced denominator."""
    if pred.is_cuda and (mask.shape[-1] == 1 or mask.stride(-1) == 0):      # one fused pass (csrc/psnode_loss.cu)
        from .losses import masked_sse
        return masked_sse(pred, target, mask), mask.sum()
    # host tensors (the gloo tests of the sharding logic) and per-feature masks: plain torch
    se = torch.nn.functional.mse_loss(pred, target, reduction="none") * mask
    return se.sum(), mask.sum()


def sharded_training_step(model_forward, params, bucket: GradBucket, numerator_and_mask_sum, group=None) -> float:
    """One data-parallel optimiser-step's worth of gradients with the reference's global-mask normalisation.

    model_forward() -> prediction(s) on this rank's shard; numerator_and_mask_sum(pred) -> (num, den): the UN-normalised
    masked loss sum of the shard and its sum(mask) (for the DAE script: num = sum(mse_x*mask) + 9*sum(mse_x[...,1:2]*mask)
    + sum(mse_i*mask), den = sum(mask), neural_01_DAE_01_no_encode.py:414-418).
    On return every p.grad holds d(num_global / den_global)/dp, identical on every rank and equal (up to fp32 summation
    order) to the single-process gradient of the reference's loss; returns the global loss value."""
    for p in params:
        p.grad = None
    num, den = numerator_and_mask_sum(model_forward())
    num.backward()
    red = bucket.allreduce_(extras=(den, num.detach()), group=group, normalise_by_extra=0)
    return (red[1] / red[0]).item()
