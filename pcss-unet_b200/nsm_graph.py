"""CUDA-graph capture of the B200 hot path (whole training step, eval forward).

The stage sequence of a training step is ~260 kernel launches driven from Python (nsm_train.py).  Every launch is cheap
on the device, so at batch 32 x 512^2 the host side -- ctypes calls, tensor-map encoding, allocator traffic -- decides
whether the GPU ever waits.  ``GraphedTrainStep`` runs the unchanged step once under ``torch.cuda.graph`` (after warm-up
steps that settle every lazily created buffer) and then replays it: one ``cudaGraphLaunch`` per step, kernels back to back,
no Python between them.  Everything the step does is capturable by construction: TMA descriptors and multi-tensor pointer
tables travel as kernel parameters, the Dropout2d draws use torch's graph-safe Philox offsets, the AdamW step counter
lives on the device (nsm_adamw_clip_step), the [0,1] range check of the loss is read back lazily instead of with
``.item()`` inside the step, NCCL gradient buckets (parallel.GradSync) are captured like any other stream work.

    step = GraphedTrainStep(model, criterion, optimizer, x_example, t_example)   # captures
    loss = step(x, t)            # copies into the static buffers, replays; `loss` is a device scalar (static tensor)
    step.check()                 # optional: raises if any captured range check fired since the last call
"""
from __future__ import annotations

import torch

import nsm


class GraphedTrainStep:
    def __init__(self, model, criterion, optimizer, x, t, sync=None, warmup=3, pool=None):
        nsm.require_device(x)
        self.model, self.criterion, self.optimizer, self.sync = model, criterion, optimizer, sync
        self.x = x.detach().clone()
        self.t = t.detach().clone()
        self.graph = torch.cuda.CUDAGraph()
        self._takes_model = hasattr(criterion, "perturbation_loss")      # main.py:215,265 dispatch
        crit_flags = [m for m in criterion.modules() if hasattr(m, "lazy_range_check")]
        for m in crit_flags:
            m.lazy_range_check = True
        self._flags = crit_flags
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # warm-up off the default stream, as torch's capture recipe asks
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.check()
        with torch.cuda.graph(self.graph, pool=pool):
            self.loss = self._step()
        self.launches_per_replay = None

    def _step(self):
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model(self.x)
        if self._takes_model:
            loss, _ = self.criterion(self.model, out, self.t, self.x)
        else:
            loss = self.criterion(out, self.t, self.x)
        loss.backward()
        if self.sync is not None:
            self.sync.finish()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, x=None, t=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss

    def check(self):
        """Host read of the deferred `0 <= output <= 1` checks (customLoss.py:131); one synchronisation."""
        for m in self._flags:
            m.raise_if_out_of_range()


class GraphedInfer:
    """Eval forward of a fixed shape as one graph launch: ``y = GraphedInfer(model, x_example)(x)``."""

    def __init__(self, model, x, warmup=2):
        nsm.require_device(x)
        assert not model.training
        self.model = model
        self.x = x.detach().clone()
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                model(self.x)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.y = model(self.x)

    def __call__(self, x=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
