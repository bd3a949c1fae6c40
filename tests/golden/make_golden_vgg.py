#!/usr/bin/env python
"""Golden vectors for the perceptual term from the UNMODIFIED reference class (customLoss.MultiLayerVGGLoss,
customLoss.py:7-90) -> tests/golden/vgg_vectors.npz.  Build container only (needs /root/reference).

ImageNet weights cannot be downloaded here, so torchvision's constructor is replaced by `seeded_vgg19` (random init under
a fixed seed, biases randomised too so that the bias path is exercised) -- the same function the tests use to rebuild the
identical frozen network.  What is pinned is the reference's ARITHMETIC around the network (clamp, grey -> 3 channels,
normalisation, the five truncated stacks, nan_to_num, weighted L1, detached result), not ImageNet features.

    python tests/golden/make_golden_vgg.py
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


sys.path.insert(0, os.path.dirname(HERE))
from vgg_fixture import cases, seeded_vgg19  # noqa: E402


def main():
    for name in ("graphviz", "pytorch_msssim"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pytorch_msssim"].ssim = None
    sys.path.insert(0, REF)
    torchvision.models.vgg19 = seeded_vgg19
    import customLoss as ref_loss
    torch.set_num_threads(1)
    crit = ref_loss.MultiLayerVGGLoss("cpu")
    res = {}
    for tag, (o, t) in cases().items():
        res[f"vgg_{tag}"] = np.float64(crit(o, t).item())
        with torch.autocast("cpu", dtype=torch.bfloat16):
            res[f"vgg_{tag}_bf16"] = np.float64(crit(o, t).item())
    # the full CustomLoss value with the term in (customLoss.py:160,193)
    cl = ref_loss.CustomLoss("cpu", alpha=0.9)
    o, t = cases()["a"]
    res["custom_loss_a"] = np.float64(cl(o, t, None).item())
    np.savez(os.path.join(HERE, "vgg_vectors.npz"), **res)
    print({k: float(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
