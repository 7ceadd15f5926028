"""Mesh-level data parallelism for GraphSAGE training (one process per GPU, torchrun).

The reference trains in a single process (``scripts/train_gnn.py:112-284``); its batches are
disjoint unions of whole meshes (``:150``), so the path shards by mesh with no data-path
exchange.  The only collective is the gradient all-reduce: <= 2.2 MB of fp32, i.e.
latency-bound over NVLink/NVSwitch.  Design:

* all gradients live in ONE flat fp32 buffer (``p.grad`` are views), split into a few buckets in
  backward order (decoder, convs L-1..0, encoder);
* a post-accumulate hook per parameter launches the bucket's asynchronous NCCL all-reduce as soon
  as its last gradient lands, so it overlaps the rest of the backward;
* ``finish()`` waits for the buckets before the optimizer step;
* ``scale_loss`` reproduces the single-process loss of the union batch exactly: the reference's
  ``MaskedMSELoss`` is a mean over the masked nodes of the WHOLE batch (``model.py:151``), so each
  rank back-propagates its SUM of squared errors (``loss_r * count_r``); the rank's masked count rides in
  one extra element of the FIRST bucket that is reduced (no collective of its own), and ``finish()``
  divides the summed gradients by the summed count.
Works with any backend (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


class MeshDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, num_buckets: int = 3, overlap: bool = True):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        dev, dt = params[0].device, params[0].dtype
        # backward produces gradients roughly in reverse registration order
        order = list(reversed(params))
        sizes = [p.numel() for p in order]
        total = sum(sizes)
        # one extra element right behind the first parameter in backward order: this rank's masked count (see scale_loss)
        self.flat = torch.zeros(total + 1, dtype=dt, device=dev)
        per = (total + num_buckets - 1) // max(num_buckets, 1)
        self.buckets: list[tuple[int, int]] = []
        self._bucket_of: dict[int, int] = {}
        self._pending: list[int] = []
        off, start, b = 0, 0, 0
        self._count_slot = None
        for p, n in zip(order, sizes):
            p.grad = self.flat[off:off + n].view_as(p)
            self._bucket_of[id(p)] = b
            off += n
            if self._count_slot is None:
                self._count_slot = off
                off += 1
            if off - start >= per or off == total + 1:
                self.buckets.append((start, off))
                start, b = off, b + 1
        self._remaining = [0] * len(self.buckets)
        self._count = [0] * len(self.buckets)
        for p in order:
            self._count[self._bucket_of[id(p)]] += 1
        self._handles: list = []
        self._scaled = False
        self._reset()
        if self.world > 1:
            # identical initial weights on every rank
            for p in module.parameters():
                dist.broadcast(p.data, src=0, group=self.pg)
            for p in order:
                p.register_post_accumulate_grad_hook(self._hook)

    def _reset(self):
        self._remaining = list(self._count)
        self._handles = []

    def _hook(self, p):
        b = self._bucket_of[id(p)]
        self._remaining[b] -= 1
        if self._remaining[b] == 0 and self.overlap:
            self._launch(b)

    def _launch(self, b):
        s, e = self.buckets[b]
        self._handles.append(dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False):  # keep the flat views alive
        self.flat.zero_()

    def scale_loss(self, loss: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
        """``loss_r`` (a mean over ``count_r`` masked elements) -> the quantity to call ``backward()`` on.

        Returns ``loss_r * count_r`` (this rank's SUM of squared errors) and starts an asynchronous all-reduce of the
        count; ``finish()`` divides the summed gradients by the summed count, which is exactly the gradient of the
        mean over the union batch (``model.py:151``).  Nothing waits between forward and backward: the count travels
        while the backward runs (a blocking all-reduce here was a synchronisation point of all ranks mid-step)."""
        if self.world == 1:
            return loss
        c = count.detach().float().reshape(1)
        self.flat[self._count_slot:self._count_slot + 1].copy_(c)  # travels with the first gradient bucket
        self._scaled = True
        return loss * c.reshape(())

    def finish(self):
        """Call after ``backward()`` and before ``optimizer.step()``."""
        if self.world > 1:
            if not self.overlap:
                for b in range(len(self.buckets)):
                    self._launch(b)
            else:
                for b, r in enumerate(self._remaining):
                    if r > 0:  # parameters that received no gradient this step
                        self._launch(b)
            for h in self._handles:
                h.wait()
            if self._scaled:  # gradients of per-rank SUMS -> gradient of the union-batch mean
                self.flat.div_(self.flat[self._count_slot:self._count_slot + 1].clamp_min(1.0))
                self._scaled = False
            else:  # scale_loss was not used: plain average of the ranks' gradients
                self.flat.div_(self.world)
        self._reset()
