"""Data-parallel training of the drop-in Unet: one process per GPU, per-rank batches, NCCL all-reduce (mean) of
gradient buckets over NVLink, launched while the rest of backward is still running.

The reference has no distributed code at all; this is the new functionality BASELINE.json's configs[3]
asks for.  Semantics follow "every GPU behaves exactly like the reference at its own batch": BatchNorm uses per-replica
batch statistics (no SyncBN), Dropout2d draws come from a per-rank generator, gradients are averaged over ranks, BN
running buffers of rank 0 are the ones broadcast / saved (as torch DDP does).

Bucket plan, in the order backward produces the gradients:
    {conv10, conv9, conv8} . {conv7} . {conv6 1x1 + BNs + biases} . {conv6.conv.0.weight (37.7 MB alone)} . {conv5}
    . {conv4, conv3, conv2}
`nsm_train._backward` hands every block's gradients to `reduce_ready` as soon as they exist; a bucket whose members
are complete is flattened and all-reduced asynchronously (NCCL runs on its own stream), `flush` (end of backward) makes
the compute stream wait for the collectives and scatters the averaged results back (NCCL averages inside the collective).  Works with any
torch.distributed backend (gloo on CPU in the tests).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist


def bucket_plan(names: List[str]) -> List[List[str]]:
    """Partition of the 66 parameter names into the six buckets above (every name exactly once)."""
    def blk(n):
        return n.split(".")[0]
    groups = [
        [n for n in names if blk(n) in ("conv10", "conv9", "conv8")],
        [n for n in names if blk(n) == "conv7"],
        [n for n in names if blk(n) == "conv6" and n != "conv6.conv.0.weight"],
        [n for n in names if n == "conv6.conv.0.weight"],
        [n for n in names if blk(n) == "conv5"],
        [n for n in names if blk(n) in ("conv4", "conv3", "conv2")],
    ]
    flat = [n for g in groups for n in g]
    assert sorted(flat) == sorted(names), "bucket plan must cover every parameter exactly once"
    return [g for g in groups if g]


class GradSync:
    def __init__(self, model, group=None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.names = [n for n, _ in model.named_parameters()]
        self.buckets = bucket_plan(self.names)
        self._bucket_of = {n: i for i, b in enumerate(self.buckets) for n in b}
        self._pending: List[Dict[str, torch.Tensor]] = [dict() for _ in self.buckets]
        self._inflight = []      # (bucket index, flat tensor, [tensors], work handle)
        self.synced_in_backward = False
        # NCCL averages inside the collective (ReduceOp.AVG): no arithmetic left on the PyTorch side, only the bucket
        # gather / scatter copies.  gloo (CPU tests) has no AVG: sum, then scale.
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        model._grad_sync = self  # picked up by nsm_train._backward

    # ---- start-up consistency ---------------------------------------------------------------------------------
    @torch.no_grad()
    def broadcast_state(self, src=0):
        """Rank `src`'s parameters and buffers become everybody's (what DDP does at construction)."""
        if self.world == 1:
            return
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t, src, group=self.group)     # on the tensor itself: bumps Tensor._version
        # packed-operand caches are keyed on (data_ptr, _version); drop them anyway so that no rank can keep operands
        # packed from its pre-broadcast parameters
        getattr(self.model, "_packed", {}).clear()
        self.model.__dict__.pop("_train_packed", None)

    # ---- overlapped path: called from inside backward ------------------------------------------------------------
    def reduce_ready(self, grads: Dict[str, torch.Tensor]):
        if self.world == 1:
            return
        touched = set()
        for n, g in grads.items():
            if n in self._bucket_of:
                b = self._bucket_of[n]
                self._pending[b][n] = g
                touched.add(b)
        for b in sorted(touched):
            if len(self._pending[b]) == len(self.buckets[b]):
                self._launch(b)

    def _launch(self, b):
        tensors = [self._pending[b][n] for n in self.buckets[b]]
        flat = torch.cat([t.reshape(-1) for t in tensors])
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        self._inflight.append((b, flat, tensors, work))
        self._pending[b] = dict()

    def flush(self):
        """End of backward: order the compute stream after the collectives, average, scatter back in place."""
        for b, flat, tensors, work in self._inflight:
            work.wait()                     # NCCL: stream-level wait, the host does not block
            if not self._avg:
                flat.mul_(1.0 / self.world)
            off = 0
            for t in tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        self.synced_in_backward = bool(self._inflight)
        self._inflight = []
        assert all(not p for p in self._pending), "gradient bucket incomplete at the end of backward"

    # ---- post-backward path (used when gradients did not come through nsm_train, e.g. foreign parameters) --------
    def finish(self):
        if self.world == 1 or self.synced_in_backward:
            self.synced_in_backward = False
            return
        params = dict(self.model.named_parameters())
        for names in self.buckets:
            ts = [params[n].grad for n in names if params[n].grad is not None]
            if not ts:
                continue
            flat = torch.cat([t.reshape(-1) for t in ts])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group)
            if not self._avg:
                flat.mul_(1.0 / self.world)
            off = 0
            for t in ts:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n


def shard_frames(num_frames: int, rank: int, world: int):
    """Inference sharding (BASELINE configs[4]): contiguous frame ranges per rank, no collective."""
    base, extra = divmod(num_frames, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
