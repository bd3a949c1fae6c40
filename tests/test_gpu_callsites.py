"""The reference's call sites, run for real: `main.main()` -> `train_model` / `validate_direct` (main.py:132-664, fp16
autocast + GradScaler + set_detect_anomaly + backward hooks + gradient hygiene + AdamW + LambdaLR + checkpoint save) and
`infer.main()` (infer.py:22-80, odd-sized frame), imported UNMODIFIED from the run-time copy in baseline/_ref, once with
the B200 drop-in first on sys.path and once with the reference's own classes on stock PyTorch (tests/callsite_driver.py,
one process per arm).  Asserts that the drop-in modules are the ones the call sites imported, that both arms train on
the same loss curve (north_star: within 1 %), and that infer.py's PNG matches the oracle on the saved checkpoint."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "callsite_driver.py")


def run_arm(arm, workdir, *extra):
    cmd = [sys.executable, DRIVER, "--arm", arm, "--workdir", str(workdir)] + list(extra)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("CALLSITE_RESULT ")]
    if not lines:
        if '"unavailable"' in r.stdout:
            pytest.skip("baseline/_ref not staged (run __graft_entry__.build() where /root/reference exists)")
        raise AssertionError(f"{arm} arm failed (exit {r.returncode})\nSTDOUT:\n{r.stdout[-3000:]}\nSTDERR:\n{r.stderr[-6000:]}")
    return json.loads(lines[-1][len("CALLSITE_RESULT "):]), r


@pytest.fixture(scope="module")
def arms(tmp_path_factory):
    out = {}
    for arm in ("dropin", "reference"):
        out[arm] = run_arm(arm, tmp_path_factory.mktemp(arm))
    return out


def test_call_sites_import_the_dropin(arms):
    res, _ = arms["dropin"]
    pkg = os.path.join(ROOT, "pcss-unet_b200")
    assert res["which"]["file"].startswith(pkg), res["which"]
    assert res["which"]["CustomLoss_file"].startswith(pkg), res["which"]
    assert res.get("infer_unet_file", "").startswith(pkg)
    ref, _ = arms["reference"]
    assert "baseline/_ref" in ref["which"]["file"]


def test_dataset_stats_file_matches_reference(arms):
    """a1 through the real consumer: main.load_dataset reads the train_stats.npy each arm's calculate_dataset_stats wrote."""
    a, b = arms["dropin"][0]["stats"], arms["reference"][0]["stats"]
    assert np.allclose(a["means"], b["means"], rtol=0, atol=1e-6) and np.allclose(a["stds"], b["stds"], rtol=1e-6)


def test_train_model_loss_curve_matches_reference_arm(arms):
    a, b = arms["dropin"][0]["scalars"], arms["reference"][0]["scalars"]
    la, lb = a["Loss/Train/Total"], b["Loss/Train/Total"]
    assert len(la) == len(lb) == 12, (len(la), len(lb))        # 4 epochs x 3 batches: no batch was skipped or swallowed
    dev = [abs(x - y) / y for x, y in zip(la, lb)]
    print("train_model loss, drop-in vs reference arm:", [f"{x:.5f}/{y:.5f}" for x, y in zip(la, lb)])
    print("relative deviation per step:", [f"{d:.4f}" for d in dev])
    assert max(dev) <= 0.01
    l1a, l1b = a["Loss/Train/L1"], b["Loss/Train/L1"]
    assert max(abs(x - y) / y for x, y in zip(l1a, l1b)) <= 0.01
    # the VGG term main.py:277 back-derives from the loss (seeded-random VGG19 in both arms)
    va, vb = a["Loss/Train/VGG"], b["Loss/Train/VGG"]
    print("vgg term:", [f"{x:.5f}/{y:.5f}" for x, y in zip(va, vb)])
    assert max(abs(x - y) / max(abs(y), 1e-6) for x, y in zip(va, vb)) <= 0.05
    # validation (model.eval(), inference_mode, fp16 autocast) after every epoch
    vla, vlb = a["Loss/Val/Total"], b["Loss/Val/Total"]
    assert len(vla) == len(vlb) == 4
    assert max(abs(x - y) / y for x, y in zip(vla, vlb)) <= 0.01
    assert a["Learning_Rate"] == b["Learning_Rate"]
    assert la[-1] < la[0]                                       # and it trains


def test_infer_py_png_matches_oracle_on_saved_checkpoint(arms, tmp_path):
    import cv2
    res, _ = arms["dropin"]
    assert res["checkpoint"] and res["png"]
    png = cv2.imread(res["png"], cv2.IMREAD_UNCHANGED)
    assert png.shape == (74, 98) and png.dtype == np.uint8      # infer.py:55-59 even-size fix
    wd = os.path.dirname(res["png"])
    sd = torch.load(os.path.join(wd, "checkpoints", "best_model.pth"), map_location="cpu")["model_state_dict"]
    frame = torch.from_numpy(np.load(os.path.join(wd, "frame.npy"))).unsqueeze(0)
    x = torch.nn.functional.interpolate(frame, (74, 98), mode="bilinear", align_corners=True)
    with torch.no_grad():
        ref = oracle.unet_forward(x, {k: v.float() if v.is_floating_point() else v for k, v in sd.items()},
                                  training=False)
    # infer.py:64-79 as the reference runs it: the model output has the autocast dtype (fp16), `.numpy() * 255` is an fp16
    # product, `.astype(uint8)` truncates.  The drop-in's fp32-mode result (<= 1e-4 from the truth) goes through the same
    # fp16 rounding; a pixel can differ by one grey level only where the two fp32 values straddle an fp16 rounding boundary
    # that also straddles an integer
    want = (ref.to(torch.float16).squeeze().numpy() * 255).astype(np.uint8)
    diff = np.abs(png.astype(np.int32) - want.astype(np.int32))
    print("infer.py PNG vs oracle: max level diff", diff.max(), "pixels differing", int((diff > 0).sum()), "of", diff.size)
    assert diff.max() <= 1 and (diff > 0).mean() <= 0.02
