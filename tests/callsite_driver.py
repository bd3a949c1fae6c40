#!/usr/bin/env python
"""Runs the reference's OWN call sites -- `main.main()` (main.py:869-979 -> train_model :132-581, validate_direct
:583-664) and `infer.main()` (infer.py:22-80) -- unmodified, from the run-time copy in baseline/_ref, in one of two arms:

  --arm dropin     pcss-unet_b200/ is first on sys.path, so `from Unetmodel import Unet`, `from customLoss import
                   CustomLoss`, `from pert_loss import EnhancedCustomLoss` resolve to the B200 drop-in modules;
  --arm reference  only baseline/_ref is on sys.path: the reference's own classes on stock PyTorch.

Everything else is the reference's code path: config.ini parsing, MmapLiverDataset + DataLoader on .npy files,
fp16 autocast + GradScaler, set_detect_anomaly, the backward hooks, gradient hygiene, clip, AdamW, LambdaLR, validation,
checkpoint save; then infer.py loads that checkpoint and runs an odd-sized frame.  A separate process per arm because
main.py configures the process at import (default device cuda, memory fraction, seeds).

Out-of-scope I/O is stubbed, identically in both arms: graphviz / pytorch_msssim / OpenEXR / Imath / colorama modules
(unused on the path), TensorBoard's SummaryWriter (replaced by a recorder that keeps the scalars -- that is how the
per-step losses get out), `read_exr` (returns the synthetic frame) and torchvision's VGG19 ImageNet weights (no network:
seeded random init).  Test infrastructure; prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pcss-unet_b200")

CONFIG = """[base]
batch_size={batch}
num_epochs={epochs}
learning_rate=0.0007
dropout_rate={dropout}
optimizer_type=adamw
warmup_epochs=2
perturbation_count=3
perturb_weight=0.1
save_dir=./checkpoints
processed_data_dir=./data/processed
image_width=96
image_height=64
input_channels=4
output_channels=1
alpha=0.9
loss_type={loss_type}
log_dir=./logs
"""


class Recorder:
    """Stand-in for torch.utils.tensorboard.SummaryWriter: keeps add_scalar() calls."""
    scalars = {}

    def __init__(self, *a, **k):
        pass

    def add_scalar(self, tag, value, step=None):
        Recorder.scalars.setdefault(tag, []).append(float(value))

    def add_images(self, *a, **k):
        pass

    def close(self):
        pass


def make_dataset(workdir, n_train, n_val, H, W):
    import numpy as np
    rng = np.random.default_rng(0)
    d = os.path.join(workdir, "data", "processed")
    os.makedirs(d, exist_ok=True)
    mu = np.array([0.1, -1.0, 5.0, 0.0], np.float32).reshape(1, 4, 1, 1)
    sd = np.array([1.0, 2.0, 3.0, 0.5], np.float32).reshape(1, 4, 1, 1)
    for split, n in (("train", n_train), ("val", n_val)):
        x = (rng.standard_normal((n, 4, H, W), dtype=np.float32) * sd + mu).astype(np.float32)
        # a learnable target: smooth function of the inputs squashed into (0,1)
        t = 1.0 / (1.0 + np.exp(-(0.7 * (x[:, :1] - mu[:, :1]) + 0.2 * (x[:, 1:2] - mu[:, 1:2]))))
        np.save(os.path.join(d, f"{split}_inputs.npy"), x)
        np.save(os.path.join(d, f"{split}_labels.npy"), t.astype(np.float32))
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", required=True, choices=["dropin", "reference"])
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--loss-type", default="standard")
    ap.add_argument("--height", type=int, default=64)
    ap.add_argument("--width", type=int, default=96)
    args = ap.parse_args()

    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_harness
    if not ref_harness.available():
        print(json.dumps({"unavailable": "baseline/_ref not staged"}))
        return
    ref_harness.install_stubs()
    sys.path.insert(0, ref_harness.REF_DIR)
    if args.arm == "dropin":
        sys.path.insert(0, PKG)

    import numpy as np
    import torch
    import torchvision

    # ImageNet weights cannot be downloaded: the same seeded random VGG19 in both arms (SURVEY 8c "VGG19 weights")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from vgg_fixture import seeded_vgg19
    torchvision.models.vgg19 = seeded_vgg19

    os.makedirs(args.workdir, exist_ok=True)
    os.chdir(args.workdir)
    with open("config.ini", "w") as f:
        f.write(CONFIG.format(batch=args.batch, epochs=args.epochs, dropout=args.dropout, loss_type=args.loss_type))
    data_dir = make_dataset(args.workdir, 3 * args.batch, args.batch, args.height, args.width)
    os.makedirs("checkpoints", exist_ok=True)     # main.py:539 does not create save_dir

    # a1: the statistics file main.load_dataset insists on (main.py:793-798) comes from calculate_dataset_stats -- the
    # drop-in's (GPU kernel) in the drop-in arm, the reference's (NumPy) in the reference arm
    import calculate_dataset_stats as cds
    cds.calculate_dataset_stats(data_dir)
    stats = np.load(os.path.join(data_dir, "train_stats.npy"), allow_pickle=True).item()

    import main as ref_main                      # baseline/_ref/main.py, unmodified
    ref_main.SummaryWriter = Recorder
    which = {"Unet": ref_main.Unet.__module__, "file": sys.modules[ref_main.Unet.__module__].__file__,
             "CustomLoss_file": sys.modules[ref_main.CustomLoss.__module__].__file__}
    sys.argv = ["main.py"]
    ref_main.main()

    ckpt = os.path.join(args.workdir, "checkpoints", "best_model.pth")
    out = {"arm": args.arm, "which": which, "scalars": Recorder.scalars, "checkpoint": os.path.exists(ckpt),
           "stats": {k: [float(v) for v in stats[k]] for k in ("means", "stds")}}

    # ---- infer.py on an odd-sized frame with the checkpoint the trainer just wrote ---------------------------------
    if out["checkpoint"]:
        torch.set_default_device("cpu")          # infer.py is its own process in real use
        import infer as ref_infer                # baseline/_ref/infer.py, unmodified
        frame = np.random.default_rng(5).standard_normal((4, 75, 98)).astype(np.float32)
        np.save(os.path.join(args.workdir, "frame.npy"), frame)
        ref_infer.read_exr = lambda path: [frame[c] for c in range(4)]
        png = os.path.join(args.workdir, "out.png")
        sys.argv = ["infer.py", "--input", "frame.exr", "--output", png, "--weights", ckpt, "--device", "cuda"]
        ref_infer.main()
        out["png"] = png if os.path.exists(png) else None
        out["infer_unet_file"] = sys.modules[ref_infer.Unet.__module__].__file__
    print("CALLSITE_RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
