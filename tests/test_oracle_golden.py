"""Pins the CPU oracle (oracle/) against vectors produced by the unmodified reference classes
(tests/golden/make_golden.py -> tests/golden/reference_vectors.npz).  CPU only."""
import os

import numpy as np
import torch

import oracle


def gen(seed):
    return torch.Generator().manual_seed(seed)


def test_default_init_matches_reference(golden):
    P = oracle.init_params(42)
    keys = list(golden["init_keys"])
    assert keys == list(P.keys()), "state_dict key set / order differs from Unet().state_dict()"
    assert len(keys) == 114
    sums = golden["init_sums"]
    for k, (s, a) in zip(keys, sums):
        v = P[k].double()
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k
    assert sum(P[k].numel() for k in oracle.param_names()) == 15_739_556


def _eval_case(golden, tag, bf16):
    shape = tuple(int(v) for v in golden[f"{tag}_shape"])
    P = oracle.init_params(42)
    x = torch.randn(*shape, generator=gen(1))
    oracle.calibrate_bn(P, x, generator=gen(2))
    with torch.no_grad():
        y = oracle.unet_forward(x, P, training=False, bf16=bf16)
    ref = torch.from_numpy(golden[f"{tag}_out_bf16" if bf16 else f"{tag}_out"])
    assert y.shape == ref.shape
    assert ref.shape[2] == shape[2] - shape[2] % 2 and ref.shape[3] == shape[3] - shape[3] % 2
    return (y.float() - ref).abs().max().item()


def test_eval_forward_fp32(golden):
    # accumulation-order noise across thread counts / ISA is ~1e-6
    for tag in ("eval_a", "eval_b", "eval_c"):
        assert _eval_case(golden, tag, False) <= 5e-6, tag


def test_eval_forward_bf16_autocast(golden):
    # same rounding points; a bf16 ulp near 1.0 is 7.8e-3, allow isolated one-ulp flips
    for tag in ("eval_a", "eval_b", "eval_c"):
        assert _eval_case(golden, tag, True) <= 8e-3, tag


def test_train_step(golden):
    P = oracle.init_params(42)
    x = torch.randn(2, 4, 32, 48, generator=gen(3))
    t = torch.rand(2, 1, 32, 48, generator=gen(4))
    torch.manual_seed(7)
    out, loss, grads = oracle.train_step_grads(x, t, P, masks=None, alpha=0.9, dropout_rate=0.2,
                                               input_grad=True)
    assert (out - torch.from_numpy(golden["train_out"])).abs().max().item() <= 5e-6
    assert abs(loss.item() - float(golden["train_loss"])) <= 1e-6
    names = list(golden["train_grad_names"])
    assert names == oracle.param_names()
    norms = golden["train_grad_norms"]
    num = den = 0.0
    for n, ref in zip(names, norms):
        num += (float(grads[n].double().norm()) - ref) ** 2
        den += ref ** 2
    assert (num / den) ** 0.5 <= 5e-3            # reference-vs-reference floor is 1e-3..2.5e-3
    for k in golden.files:
        if k.startswith("train_grad::"):
            g = grads[k.split("::")[1]]
            ref = torch.from_numpy(golden[k])
            rel = (g - ref).norm() / ref.norm().clamp_min(1e-12)
            assert rel <= 5e-3, (k, float(rel))
        if k.startswith("train_buf::"):
            ref = torch.from_numpy(golden[k])
            assert torch.allclose(P[k.split("::")[1]], ref, rtol=1e-5, atol=1e-6), k
    # checkpoint(conv5) re-runs conv5 in backward: its BN buffers advance twice per step
    assert int(P["conv5.conv.1.num_batches_tracked"]) == int(golden["train_nbt"]) == 2
    assert int(P["conv4.conv.1.num_batches_tracked"]) == 1


def test_perturbation_loss(golden):
    P = oracle.init_params(42)
    x = torch.randn(2, 4, 32, 48, generator=gen(5))
    torch.manual_seed(11)
    fwd = lambda inp: oracle.unet_forward(inp, P, training=True, dropout_rate=0.2)  # noqa: E731
    with torch.no_grad():
        y = fwd(x)
    assert (y - torch.from_numpy(golden["pert_out"])).abs().max().item() <= 5e-6
    val, ys = oracle.perturbation_loss(fwd, x, y, count=3)
    assert abs(val.item() - float(golden["pert_loss"])) <= 2e-6
    # 1 + 3 train-mode forwards advance the BN counters four times
    assert int(P["conv2.conv.1.num_batches_tracked"]) == int(golden["pert_nbt"]) == 4
    torch.manual_seed(13)
    pin = torch.stack(oracle.perturb_inputs(x, 3))
    assert torch.equal(pin, torch.from_numpy(golden["pert_inputs"]))
    g = oracle.perturbation_loss_grad(y, ys)
    yy = y.clone().requires_grad_(True)
    tot = sum(torch.nn.functional.l1_loss(yy, v) for v in ys) / 3
    tot.backward()
    assert torch.allclose(g, yy.grad, atol=1e-12)


def test_channel_stats_and_standardise(golden):
    rng = np.random.default_rng(0)
    mu = np.array([0.1, -1.0, 5.0, 0.0], dtype=np.float32).reshape(1, 4, 1, 1)
    sg = np.array([1.0, 2.0, 3.0, 0.5], dtype=np.float32).reshape(1, 4, 1, 1)
    data = (rng.standard_normal((6, 4, 64, 96), dtype=np.float32) * sg + mu).astype(np.float32)
    st = oracle.channel_stats(data)
    assert np.array_equal(np.array(st["means"]), golden["stats_means"])
    assert np.array_equal(np.array(st["stds"]), golden["stats_stds"])
    xs = oracle.standardise(torch.from_numpy(data[3]), st["means"], st["stds"])
    assert torch.equal(xs, torch.from_numpy(golden["stats_sample3_std"]))


def test_custom_loss_value_and_grad(golden):
    o = torch.from_numpy(golden["loss_o"])
    t = torch.from_numpy(golden["loss_t"])
    assert abs(oracle.custom_loss(o, t, 0.9).item() - float(golden["loss_val"])) <= 1e-7
    assert torch.equal(oracle.custom_loss_grad(o, t, 0.9), torch.from_numpy(golden["loss_grad"]))


def test_vgg_perceptual_term_matches_reference():
    """oracle.vgg_perceptual_loss vs the unmodified reference MultiLayerVGGLoss / CustomLoss (tests/golden/
    make_golden_vgg.py; same seeded VGG19 stand-in on both sides)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from vgg_fixture import cases, seeded_vgg19
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vgg_vectors.npz"))
    feats = seeded_vgg19().features.eval()
    torch.set_num_threads(1)
    for tag, (o, t) in cases().items():
        v = oracle.vgg_perceptual_loss(o, t, feats).item()
        assert abs(v - float(gold[f"vgg_{tag}"])) <= 1e-6 * abs(float(gold[f"vgg_{tag}"])), (tag, v)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            vb = oracle.vgg_perceptual_loss(o, t, feats).item()
        assert abs(vb - float(gold[f"vgg_{tag}_bf16"])) <= 1e-6 * abs(vb), (tag, vb)
    o, t = cases()["a"]
    full = oracle.custom_loss(o, t, 0.9, vgg_const=oracle.vgg_perceptual_loss(o, t, feats)).item()
    assert abs(full - float(gold["custom_loss_a"])) <= 1e-6


def test_enhanced_custom_loss_of_customloss_py_matches_reference():
    """oracle.enhanced_mse_loss vs the unmodified reference customLoss.EnhancedCustomLoss (customLoss.py:195-238;
    tests/golden/make_golden_enhanced.py): total, components and the gradient w.r.t. the output, with the reference's own
    randn_like draw replayed."""
    import torch.nn.functional as F
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "enhanced_vectors.npz"))
    torch.set_num_threads(1)
    inputs, target, w, noise = (torch.from_numpy(gold[k]) for k in ("inputs", "target", "w", "noise"))
    model = lambda x: torch.sigmoid(F.conv2d(x, w, padding=1))             # noqa: E731
    output = model(inputs).detach()
    assert torch.equal(output, torch.from_numpy(gold["output"]))
    output.requires_grad_(True)
    total, comp = oracle.enhanced_mse_loss(model, output, target, inputs, alpha=0.9, beta=0.05, noise=noise)
    total.backward()
    assert abs(total.item() - float(gold["total"])) <= 1e-7
    assert abs(comp["l1_loss"].item() - float(gold["l1"])) <= 1e-7 and float(comp["vgg_loss"]) == float(gold["vgg"])
    assert abs(comp["perturbation_loss"].item() - float(gold["pert"])) <= 1e-9
    assert torch.allclose(output.grad, torch.from_numpy(gold["grad"]), rtol=1e-6, atol=1e-10)
