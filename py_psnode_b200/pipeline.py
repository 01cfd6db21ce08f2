"""Dataset -> pinned host memory -> device batches, double-buffered on a copy stream (SURVEY 8f "next-4").

The reference feeds its training loop with `DataLoader(dataset, batch_size, shuffle=True)` and then moves every field of every
batch with a blocking `d.to(device)` from PAGEABLE memory (neural_00_ODE_01_no_encode.py:326-347,
neural_01_DAE_01_no_encode.py:382-408): per batch one gather on the host (collate), one staging copy inside the driver and one
synchronous H2D copy per field, all on the training stream.  With the integration itself down to milliseconds those copies
are what the GPU waits for.  `DeviceBatchPipeline` keeps the `.npz`-backed sets (`ODE_Curves_Sample` / `DAE_Curves_Sample`,
neural_base.py:10-40,136-166) in pinned host memory, gathers batch k+1 into a pinned staging slot while batch k trains,
and copies it with `non_blocking=True` on a side stream; the consumer's stream only waits on the slot's event.

It yields exactly what the reference's loop unpacks -- a tuple of batch-major tensors in the dataset's `__getitem__` order
(`t, x, z, event_t, z_jump, mask` / `t, x, z, v, i, event_t, z_jump, v_jump, mask`) -- so a script changes one line:

    for data_batch in DeviceBatchPipeline(training_dataset, args.batch, device, shuffle=True):
        t, x, z, event_t, z_jump, mask = data_batch          # already on `device`

On a CPU device it degrades to plain indexing (used by the CPU tests of ordering / coverage)."""
from typing import Iterator, List, Optional, Sequence, Tuple

import torch


class DeviceBatchPipeline:
    def __init__(self, dataset, batch_size: int, device, shuffle: bool = True, drop_last: bool = False,
                 seed: Optional[int] = None, depth: int = 2, fields: Optional[Sequence[str]] = None):
        if batch_size < 1:
            raise ValueError("batch_size must be >= 1")
        self.device = torch.device(device)
        self.batch_size = int(batch_size)
        self.shuffle, self.drop_last = bool(shuffle), bool(drop_last)
        self.depth = max(int(depth), 2)
        names = list(fields) if fields is not None else [*dataset.series, *dataset.per_sample, "mask"]
        self.fields = names
        self._cuda = self.device.type == "cuda"
        host: List[torch.Tensor] = []
        for n in names:
            ten = getattr(dataset, n)
            if not torch.is_tensor(ten):
                raise TypeError(f"dataset.{n} is not a tensor")
            ten = ten.contiguous()
            if self._cuda and not ten.is_pinned():
                ten = ten.pin_memory()                   # once per dataset: the arena every batch is gathered from
            host.append(ten)
        self.host = host
        self.n = host[0].shape[0]
        if any(h.shape[0] != self.n for h in host):
            raise ValueError("dataset fields disagree on the number of samples")
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(int(seed))
        self._slots = None
        self.bytes_per_sample = sum(h[0].numel() * h.element_size() for h in host)

    def __len__(self) -> int:
        full, rem = divmod(self.n, self.batch_size)
        return full if (self.drop_last or rem == 0) else full + 1

    # ------------------------------------------------------------------------------------------------------------------
    def _make_slots(self):
        bs = self.batch_size
        slots = []
        for _ in range(self.depth):
            stage = [torch.empty((bs, *h.shape[1:]), dtype=h.dtype).pin_memory() for h in self.host]
            dev = [torch.empty((bs, *h.shape[1:]), dtype=h.dtype, device=self.device) for h in self.host]
            slots.append({"stage": stage, "dev": dev, "ready": torch.cuda.Event(), "consumed": None})
        return slots

    def _batches(self) -> List[torch.Tensor]:
        order = torch.randperm(self.n, generator=self._gen) if self.shuffle else torch.arange(self.n)
        out = [order[k:k + self.batch_size] for k in range(0, self.n, self.batch_size)]
        if self.drop_last and out and out[-1].numel() < self.batch_size:
            out.pop()
        return out

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, ...]]:
        batches = self._batches()
        if not self._cuda:
            for idx in batches:
                yield tuple(h.index_select(0, idx) for h in self.host)
            return
        if self._slots is None:
            self._slots = self._make_slots()
        copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)

        def stage(k: int):
            slot = self._slots[k % self.depth]
            idx = batches[k]
            nb = idx.numel()
            if slot["consumed"] is not None:
                copy_stream.wait_event(slot["consumed"])          # the training stream is done with this slot's device buffers
                slot["ready"].synchronize()                       # and the previous H2D out of its staging buffers has completed
            for h, st in zip(self.host, slot["stage"]):
                torch.index_select(h, 0, idx, out=st[:nb])        # host gather into pinned staging (overlaps the GPU's work)
            with torch.cuda.stream(copy_stream):
                for st, dv in zip(slot["stage"], slot["dev"]):
                    dv[:nb].copy_(st[:nb], non_blocking=True)
                slot["ready"].record(copy_stream)
            return nb

        sizes = {}
        for k in range(min(self.depth - 1, len(batches))):
            sizes[k] = stage(k)
        for k in range(len(batches)):
            nxt = k + self.depth - 1
            if nxt < len(batches):
                sizes[nxt] = stage(nxt)
            slot = self._slots[k % self.depth]
            main.wait_event(slot["ready"])
            nb = sizes.pop(k)
            yield tuple(dv[:nb] for dv in slot["dev"])
            ev = torch.cuda.Event()
            ev.record(main)                                        # everything the consumer launched on this batch so far
            slot["consumed"] = ev
