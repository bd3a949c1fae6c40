#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference classes.

Runs only in the build container (needs /root/reference); the GPU box never runs it.  The reference
has no tests or fixtures of its own, so truth = its own classes executed on seeded
inputs.  Modules the reference imports but never uses on the hot path and that are absent here
(graphviz, pytorch_msssim, OpenEXR, Imath, colorama) are stubbed with empty modules; the VGG19
perceptual term (ImageNet weights, not downloadable) is replaced by a zero constant on the reference
side -- it is a detached constant with zero gradient (customLoss.py:90).

    python tests/golden/make_golden.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

for name in ("graphviz", "pytorch_msssim", "OpenEXR", "Imath", "colorama"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["graphviz"].Digraph = object
sys.modules["pytorch_msssim"].ssim = None
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import Unetmodel as ref_model          # noqa: E402
import customLoss as ref_loss          # noqa: E402
import pert_loss as ref_pert           # noqa: E402
import calculate_dataset_stats as ref_stats  # noqa: E402
import setdata as ref_data             # noqa: E402
import oracle                          # noqa: E402

torch.set_num_threads(1)


class _ZeroVGG(torch.nn.Module):
    def __init__(self, device):
        super().__init__()

    def forward(self, output, target):
        return torch.tensor(0.0, device=output.device, requires_grad=True)


ref_loss.MultiLayerVGGLoss = _ZeroVGG


def gen(seed):
    return torch.Generator().manual_seed(seed)


def f64sum(t):
    return float(t.double().sum()), float(t.double().abs().sum())


def main():
    out = {}

    # ---- G0: default init equals torch.manual_seed(42); Unet() ---------------------------------
    torch.manual_seed(42)
    net = ref_model.Unet()
    sd = net.state_dict()
    out["init_keys"] = np.array(list(sd.keys()))
    out["init_sums"] = np.array([f64sum(v.float()) for v in sd.values()])

    # ---- G1/G2: eval forward, calibrated BN, fp32; regular and odd/ragged sizes -----------------
    for tag, shape in (("eval_a", (2, 4, 32, 48)), ("eval_b", (1, 4, 41, 57)), ("eval_c", (1, 4, 80, 112))):
        P = oracle.init_params(42)
        x = torch.randn(*shape, generator=gen(1))
        oracle.calibrate_bn(P, x, generator=gen(2))
        net = ref_model.Unet()
        net.load_state_dict(P, strict=True)
        net.eval()
        with torch.no_grad():
            y = net(x)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                yb = net(x)
        out[f"{tag}_shape"] = np.array(shape)
        out[f"{tag}_out"] = y.numpy()
        out[f"{tag}_out_bf16"] = yb.float().numpy()

    # ---- G3: one training step (train mode, Dropout2d from the global generator, CustomLoss) ----
    torch.manual_seed(42)
    net = ref_model.Unet(dropout_rate=0.2)
    net.train()
    x = torch.randn(2, 4, 32, 48, generator=gen(3)).requires_grad_(True)  # setdata.py:325-326
    t = torch.rand(2, 1, 32, 48, generator=gen(4))
    crit = ref_loss.CustomLoss(torch.device("cpu"), alpha=0.9)
    torch.manual_seed(7)           # pins the Dropout2d draws
    y = net(x)
    loss = crit(y, t, x)
    loss.backward()
    out["train_out"] = y.detach().numpy()
    out["train_loss"] = np.array(loss.item())
    names = [n for n, _ in net.named_parameters()]
    out["train_grad_names"] = np.array(names)
    out["train_grad_norms"] = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
    for n in ("conv10.weight", "conv10.bias", "conv9.conv.4.weight", "conv2.conv.0.weight",
              "conv9.conv.5.weight", "conv9.conv.5.bias", "conv6.conv.1.weight"):
        out["train_grad::" + n] = dict(net.named_parameters())[n].grad.numpy()
    out["train_grad::input"] = x.grad.numpy()
    sd = net.state_dict()
    for n in ("conv2.conv.1.running_mean", "conv2.conv.1.running_var", "conv9.conv.5.running_mean",
              "conv9.conv.5.running_var", "conv5.conv.5.running_var"):
        out["train_buf::" + n] = sd[n].numpy()
    out["train_nbt"] = np.array(int(sd["conv5.conv.1.num_batches_tracked"]))

    # ---- G4: PerturbationLoss on the train-mode model (pert_loss.py:61-90) ----------------------
    torch.manual_seed(42)
    net = ref_model.Unet(dropout_rate=0.2)
    net.train()
    x = torch.randn(2, 4, 32, 48, generator=gen(5))
    torch.manual_seed(11)
    y = net(x)
    pl = ref_pert.PerturbationLoss(perturbation_count=3)
    val = pl(net, x, y)
    out["pert_loss"] = np.array(val.item())
    out["pert_out"] = y.detach().numpy()
    out["pert_nbt"] = np.array(int(net.state_dict()["conv2.conv.1.num_batches_tracked"]))
    torch.manual_seed(13)
    pin = pl.perturb_input(x)
    out["pert_inputs"] = torch.stack(pin).numpy()

    # ---- G5: dataset channel statistics + standardisation (calculate_dataset_stats, setdata) ----
    rng = np.random.default_rng(0)
    mu = np.array([0.1, -1.0, 5.0, 0.0], dtype=np.float32).reshape(1, 4, 1, 1)
    sg = np.array([1.0, 2.0, 3.0, 0.5], dtype=np.float32).reshape(1, 4, 1, 1)
    data = (rng.standard_normal((6, 4, 64, 96), dtype=np.float32) * sg + mu).astype(np.float32)
    labels = rng.random((6, 1, 64, 96), dtype=np.float32)
    with tempfile.TemporaryDirectory() as d:
        np.save(os.path.join(d, "train_inputs.npy"), data)
        np.save(os.path.join(d, "train_labels.npy"), labels)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            stats = ref_stats.calculate_dataset_stats(d)
            ds = ref_data.MmapLiverDataset(d, split="train")
            xi, _ = ds[3]
        finally:
            os.chdir(cwd)
    out["stats_means"] = np.array(stats["means"])
    out["stats_stds"] = np.array(stats["stds"])
    out["stats_sample3_std"] = xi.detach().numpy()

    # ---- G6: CustomLoss value and gradient -------------------------------------------------------
    o = torch.rand(2, 1, 16, 24, generator=gen(6)).requires_grad_(True)
    tt = torch.rand(2, 1, 16, 24, generator=gen(7))
    ls = crit(o, tt, None)
    ls.backward()
    out["loss_o"] = o.detach().numpy()
    out["loss_t"] = tt.numpy()
    out["loss_val"] = np.array(ls.item())
    out["loss_grad"] = o.grad.numpy()

    path = os.path.join(HERE, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
