"""Drop-in for the reference ``customLoss.py`` objective on B200.

``CustomLoss(device, alpha)(output, target, inputs)`` returns ``alpha * L1 + (1 - alpha) * vgg`` exactly like
customLoss.py:129-193, keeps the attributes the trainer reads (``.l1`` callable, ``.alpha``; main.py:274-277) and
asserts ``0 <= output <= 1`` (customLoss.py:131).  The L1 value, its gradient ``alpha * sign(o - t) / N`` and the range
check come out of ONE vectorised, warp-shuffle-reduced streaming kernel (``nsm_l1_loss_fwd_bwd``).

The reference's VGG19 perceptual term is a detached constant (re-wrapped by ``torch.tensor(..., requires_grad=True)``,
customLoss.py:90): it shifts the loss value and contributes no gradient.  It needs ImageNet weights (not available
offline) and is listed as a follow-up in SURVEY 8f; pass ``vgg_loss=<callable(output, target) -> scalar>`` (for
instance the reference's own ``MultiLayerVGGLoss``) to include it, the default contributes 0.  The high-frequency,
penumbra and Sobel terms the reference computes and then discards (customLoss.py:139-185) are not computed.
"""
from __future__ import annotations

import torch
import torch.nn as nn

import nsm


class _FusedL1(torch.autograd.Function):
    """mean|o - t| with the gradient produced by the same kernel pass."""

    @staticmethod
    def forward(ctx, output, target, check_range):
        o32 = output.detach().to(torch.float32).contiguous()
        acc, sign = nsm.l1_loss_fwd_bwd(o32, target.detach(), (), coef_l1=1.0, want_grad=True)
        ctx.save_for_backward(sign)       # sign(o - t) in {-1, 0, +1}
        ctx.numel = o32.numel()
        ctx.out_dtype = output.dtype
        if check_range and float(acc[2].item()) != 0.0:     # one sync; the reference does two (:131)
            raise AssertionError("输出必须经过Sigmoid激活!")
        return (acc[0] / o32.numel()).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        (sign,) = ctx.saved_tensors
        # same evaluation order as autograd of (o - t).abs().mean(): (g / N) * sign  -> bit-identical gradient
        return (sign * (g / ctx.numel)).to(ctx.out_dtype), None, None


class L1Loss(nn.Module):
    """nn.L1Loss() stand-in (customLoss.py:96) backed by the fused kernel."""

    def __init__(self, check_range=False):
        super().__init__()
        self.check_range = check_range

    def forward(self, output, target):
        nsm.require_device(output)
        return _FusedL1.apply(output, target, self.check_range)


class CustomLoss(nn.Module):
    def __init__(self, device, alpha=0.9, vgg_loss=None):
        super().__init__()
        self.alpha = alpha
        self.device = device
        self.l1 = L1Loss()
        self._l1_checked = L1Loss(check_range=True)
        self.vgg_loss = vgg_loss

    def forward(self, output, target, inputs):
        nsm.require_device(output)
        l1 = self._l1_checked(output, target)
        if self.vgg_loss is None:
            vgg = torch.zeros((), dtype=torch.float32, device=output.device)
        else:
            with torch.no_grad():
                vgg = torch.as_tensor(self.vgg_loss(output.detach(), target), dtype=torch.float32,
                                      device=output.device).detach()
        return self.alpha * l1 + (1 - self.alpha) * vgg
