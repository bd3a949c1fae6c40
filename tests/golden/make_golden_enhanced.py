#!/usr/bin/env python
"""Golden vectors for customLoss.EnhancedCustomLoss (customLoss.py:195-238) from the UNMODIFIED reference class.

Build container only (needs /root/reference).  The VGG19 term (ImageNet weights, not downloadable) is a zero constant on
the reference side, as in make_golden.py; the network is a fixed seeded 3x3 convolution + sigmoid (the class takes any
callable).  Stored: inputs, target, network weight, the randn_like draw of customLoss.py:225 (replayed from the seed),
the three components, the total and its gradient w.r.t. the output.

    python tests/golden/make_golden_enhanced.py
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("graphviz", "pytorch_msssim", "OpenEXR", "Imath", "colorama"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["pytorch_msssim"].ssim = None
sys.path.insert(0, "/root/reference")
import customLoss as ref_loss   # noqa: E402


class _ZeroVGG(torch.nn.Module):
    def __init__(self, device):
        super().__init__()

    def forward(self, output, target):
        return torch.tensor(0.0, device=output.device, requires_grad=True)


ref_loss.MultiLayerVGGLoss = _ZeroVGG
torch.set_num_threads(1)


def main():
    g = torch.Generator().manual_seed(2024)
    inputs = torch.randn(2, 4, 17, 19, generator=g) * 4.0
    inputs[0, 0, 0, :4] = torch.tensor([11.0, -12.0, 9.995, -9.999])      # the clamp of :228 is exercised
    target = torch.rand(2, 1, 17, 19, generator=g)
    w = torch.randn(1, 4, 3, 3, generator=g) * 0.3
    model = lambda x: torch.sigmoid(F.conv2d(x, w, padding=1))             # noqa: E731
    output = model(inputs).detach().requires_grad_(True)
    crit = ref_loss.EnhancedCustomLoss("cpu", alpha=0.9, beta=0.05)
    torch.manual_seed(99)
    total, comp = crit(model, output, target, inputs)
    total.backward()
    torch.manual_seed(99)
    noise = torch.randn_like(inputs)                                        # the same draw as customLoss.py:225
    np.savez_compressed(os.path.join(HERE, "enhanced_vectors.npz"),
                        inputs=inputs.numpy(), target=target.numpy(), w=w.numpy(), noise=noise.numpy(),
                        output=output.detach().numpy(), total=np.float64(total.item()),
                        l1=np.float64(comp["l1_loss"].item()), vgg=np.float64(float(comp["vgg_loss"])),
                        pert=np.float64(comp["perturbation_loss"].item()), grad=output.grad.numpy())
    print("total", total.item(), {k: float(v) for k, v in comp.items()})


if __name__ == "__main__":
    main()
