"""Fused optimizer step for the B200 training path.

``FusedAdamWClip`` is a ``torch.optim.Optimizer`` with AdamW's state layout (``step``, ``exp_avg``, ``exp_avg_sq`` --
checkpoints written by ``main.py:539-544`` stay loadable) whose ``step()`` runs the trainer's gradient hygiene and update
(main.py:295-423: NaN/Inf scan, ``clip_grad_norm_``, AdamW) as two multi-tensor kernels of libnsm_b200.so with no host
synchronisation.  After ``step()`` the tensors ``last_grad_norm`` (global L2 norm before clipping) and
``last_nonfinite`` (number of NaN/Inf gradient elements; the update was skipped if non-zero) can be read lazily.
"""
from __future__ import annotations

import ctypes
import math

import torch

import nsm


class FusedAdamWClip(torch.optim.Optimizer):
    def __init__(self, params, lr=7e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3, max_norm=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_norm=max_norm)
        super().__init__(params, defaults)
        self._acc = None

    @property
    def last_grad_norm(self):
        return None if self._acc is None else torch.sqrt(self._acc[0]).to(torch.float32)

    @property
    def last_nonfinite(self):
        return None if self._acc is None else self._acc[1]

    @property
    def applied_steps(self):
        """Number of updates actually applied (device counter; non-finite steps are skipped and not counted)."""
        return None if self._acc is None else self._acc[2]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if sum(1 for g in self.param_groups if any(p.grad is not None for p in g["params"])) > 1:
            # clip_grad_norm_(model.parameters()) is ONE norm over everything (main.py:405); a per-group norm would differ
            raise nsm.NsmError("FusedAdamWClip supports one param group (the global gradient norm comes from one launch)")
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            nsm.require_device(ps[0])
            fresh = self._acc is None
            if fresh:
                self._acc = torch.zeros(8, dtype=torch.float64, device=ps[0].device)   # [0..2] + one nsm_acc slot of scratch
            loaded = 0
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                loaded = max(loaded, int(st["step"]))
                st["step"] += 1       # host-side count of step() calls (state_dict layout of torch AdamW); the step the
                #                       kernel uses is the device counter acc[2], which skips non-finite steps
            if fresh and loaded:
                self._acc[2] = float(loaded)          # resumed from a checkpoint (load_state_dict)
            step = 0                                  # device-side counter mode of nsm_adamw_clip_step
            for i in range(0, len(ps), 128):
                chunk = ps[i:i + 128]
                if len(ps) > 128:
                    raise nsm.NsmError("FusedAdamWClip: more than 128 tensors per param group is not supported "
                                       "(the global norm must come from one launch)")
                grads = [p.grad.contiguous() for p in chunk]
                n = len(chunk)
                vp = ctypes.c_void_p
                arr = lambda ts: (vp * n)(*[t.data_ptr() for t in ts])  # noqa: E731
                numel = (ctypes.c_longlong * n)(*[p.numel() for p in chunk])
                b1, b2 = group["betas"]
                nsm.check(nsm.lib().nsm_adamw_clip_step(
                    n, arr(chunk), arr(grads), arr([self.state[p]["exp_avg"] for p in chunk]),
                    arr([self.state[p]["exp_avg_sq"] for p in chunk]), numel, group["lr"], b1, b2, group["eps"],
                    group["weight_decay"], group["max_norm"] or 0.0, step, self._acc.data_ptr(), nsm.stream_ptr()),
                    "nsm_adamw_clip_step")
                # the kernel updated the parameters behind autograd's back: bump their version counters so that caches
                # keyed on Tensor._version (packed weights in Unetmodel / nsm_train) see the change
                torch._C._autograd._unsafe_set_version_counter(chunk, [p._version + 1 for p in chunk])
        return loss
