"""Drop-in for the reference ``pert_loss.py`` (perturbation / temporal-stability objective) on B200.

``PerturbationLoss(perturbation_count, alpha)(model, original_input, original_output)`` follows
pert_loss.py:61-90: per-channel unbiased std of the input (:42-45), ``count`` noisy copies
``x + randn * std_c * 0.01`` (:50-57), no-grad forwards of the SAME model in its current train/eval mode (:78-81),
mean over copies of L1(out, y_i) (:84-90).  On B200 the statistics come from ``nsm_channel_sums``, all copies are
produced by one ``nsm_perturb`` launch, and the loss value plus its gradient
``sum_i sign(out - y_i) / (count * N)`` come from one ``nsm_l1_loss_fwd_bwd`` pass.  The noise itself is drawn with
``torch.randn`` in the reference's (copy, channel) order -- RNG is plumbing, not arithmetic.

``EnhancedCustomLoss(device, alpha, perturb_weight)`` is what ``main.py --loss_type perturb`` instantiates
(main.py:937-940).  The reference's class cannot be constructed (``from customLoss import VGGLoss`` raises
ImportError, pert_loss.py:111); this one implements the documented intent with the call/return convention the
trainer expects: ``criterion(model, output, target, inputs) -> (loss, {'l1_loss','vgg_loss','perturbation_loss',
'total_loss'})`` and a ``perturbation_loss`` attribute (main.py:215,265-271).
"""
from __future__ import annotations

import logging

import torch
import torch.nn as nn

import nsm
from customLoss import L1Loss, _LazyRangeCheck, make_vgg_term


def channel_stds_unbiased(x):
    """torch.std(x[:, c]) for every channel (pert_loss.py:42-45), two streaming passes in fp64, kept on device."""
    B, C = x.shape[0], x.shape[1]
    n = x.numel() // C
    x32 = x.detach().to(torch.float32).contiguous()
    means = nsm.channel_sums(x32) / n
    ss = nsm.channel_sums(x32, means)
    return torch.sqrt(ss / max(n - 1, 1)).to(torch.float32)


class _FusedPerturbL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, alpha, weight, *perturbed):
        o32 = output.detach().to(torch.float32).contiguous()
        n = o32.numel()
        p = len(perturbed)
        acc, grad = nsm.l1_loss_fwd_bwd(o32, target, [y.to(torch.float32) for y in perturbed],
                                        coef_l1=(alpha / n if target is not None else 0.0),
                                        coef_pert=(weight / (p * n) if p else 0.0), want_grad=True)
        ctx.save_for_backward(grad)
        ctx.out_dtype = output.dtype
        ctx.n_extra = 3 + p
        l1 = (acc[0] / n).to(torch.float32)
        pert = (acc[1] / (max(p, 1) * n)).to(torch.float32)
        bad = acc[2]
        ctx.mark_non_differentiable(l1, pert, bad)
        total = (alpha * l1 if target is not None else 0.0) + weight * pert
        return total, l1, pert, bad

    @staticmethod
    def backward(ctx, g, *_):
        (grad,) = ctx.saved_tensors
        return ((grad * g).to(ctx.out_dtype),) + (None,) * ctx.n_extra


class PerturbationLoss(nn.Module):
    def __init__(self, perturbation_count=3, alpha=0.9):
        super().__init__()
        if not 1 <= perturbation_count <= 4:
            raise ValueError("perturbation_count must be in 1..4 (one fused loss launch)")
        self.perturbation_count = perturbation_count
        self.alpha = alpha
        self.loss_fn = L1Loss()
        logging.info(f"初始化扰动损失，扰动数量: {perturbation_count}")

    def perturb_input(self, x, std_factor=0.01, noise=None):
        """Returns the list of perturbed inputs.  ``noise`` ([count, C, B, 1, H, W]) replays given draws."""
        nsm.require_device(x)
        B, C, H, W = x.shape
        x32 = x.detach().to(torch.float32).contiguous()
        stds = channel_stds_unbiased(x32)
        if noise is None:
            noise = torch.empty(self.perturbation_count, C, B, 1, H, W, dtype=torch.float32, device=x.device)
            for i in range(self.perturbation_count):          # same draw order as pert_loss.py:50-55
                for c in range(C):
                    torch.randn(B, 1, H, W, out=noise[i, c], dtype=torch.float32, device=x.device)
        out = nsm.perturb(x32, noise, stds, std_factor)
        return [out[i] for i in range(out.shape[0])]

    def perturbed_outputs(self, model, original_input, noise=None):
        with torch.no_grad():
            return [model(p).detach() for p in self.perturb_input(original_input, noise=noise)]

    def forward(self, model, original_input, original_output, noise=None):
        ys = self.perturbed_outputs(model, original_input, noise)
        total, _, _, _ = _FusedPerturbL1.apply(original_output, None, 0.0, 1.0, *ys)
        return total


class EnhancedCustomLoss(nn.Module, _LazyRangeCheck):
    def __init__(self, device, alpha=0.9, perturb_weight=0.5, vgg_loss="auto"):
        super().__init__()
        self.alpha = alpha
        self.perturb_weight = perturb_weight
        self.l1 = L1Loss()
        self.vgg_loss = make_vgg_term(device, vgg_loss)
        self.perturbation_loss = PerturbationLoss()
        logging.info(f"初始化增强版损失函数: alpha={alpha}, perturb_weight={perturb_weight}")

    def forward(self, model, output, target, inputs, noise=None):
        nsm.require_device(output)
        use_pert = self.training and self.perturb_weight > 0
        ys = self.perturbation_loss.perturbed_outputs(model, inputs, noise) if use_pert else []
        total, l1, pert, bad = _FusedPerturbL1.apply(output, target.detach(), self.alpha,
                                                     self.perturb_weight if use_pert else 0.0, *ys)
        if self.lazy_range_check:
            self.note_range_flag(bad)
        elif float(bad.item()) != 0.0:
            raise AssertionError("输出必须经过Sigmoid激活!")
        if self.vgg_loss is None:
            vgg = torch.zeros((), dtype=torch.float32, device=output.device)
        else:
            with torch.no_grad():
                vgg = torch.as_tensor(self.vgg_loss(output.detach(), target), dtype=torch.float32,
                                      device=output.device).detach()
        total = total + (1 - self.alpha) * vgg
        losses = {"l1_loss": l1, "vgg_loss": vgg, "perturbation_loss": pert, "total_loss": total}
        return total, losses


def measure_temporal_instability(frames, motion_vectors=None, alpha=5.0):
    """Evaluation metric of pert_loss.py:170-199 (mean over frame pairs of exp(alpha*|f_t - f_{t-1}|) - 1).  Not on the
    training hot path; plain tensor ops.  The reference's motion-vector branch is an unimplemented ``pass``."""
    if len(frames) < 2:
        return torch.tensor(0.0)
    if motion_vectors is not None:
        raise NotImplementedError("the reference leaves the motion-vector branch unimplemented (pert_loss.py:187-190)")
    total = 0
    for t in range(1, len(frames)):
        total = total + torch.mean(torch.exp(alpha * torch.abs(frames[t] - frames[t - 1])) - 1)
    return total / (len(frames) - 1)
