"""Objective, statistics and perturbation kernels against the oracle and the reference golden vectors (GPU)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nsm():
    import nsm as _nsm
    _nsm.require_device()
    return _nsm


def gen(seed):
    return torch.Generator().manual_seed(seed)


def test_custom_loss_matches_reference_golden(nsm, golden):
    from customLoss import CustomLoss
    o = torch.from_numpy(golden["loss_o"]).cuda().requires_grad_(True)
    t = torch.from_numpy(golden["loss_t"]).cuda()
    crit = CustomLoss("cuda", alpha=0.9)
    loss = crit(o, t, None)
    loss.backward()
    assert abs(loss.item() - float(golden["loss_val"])) <= 1e-6
    # alpha * sign(o - t) / N; ATen's CUDA "divide by a scalar" multiplies by the reciprocal whereas the CPU golden
    # divides, so the coefficient may differ in the last bit -- the signs must agree exactly.
    ref_grad = torch.from_numpy(golden["loss_grad"])
    assert torch.equal(torch.sign(o.grad.cpu()), torch.sign(ref_grad))
    assert torch.allclose(o.grad.cpu(), ref_grad, rtol=2e-7, atol=0)
    l1 = crit.l1(o, t)
    assert abs(((loss - crit.alpha * l1) / (1 - crit.alpha)).item()) < 1e-5     # main.py:277 back-derivation of vgg


def test_l1_misaligned_views(nsm):
    """Contiguous views whose first element is not 16-byte aligned (a slice of a larger buffer): the kernel must fall back to
    its scalar loop instead of faulting on a 16-byte access."""
    big_o = torch.rand(4 * 33 * 40 + 3, generator=gen(3)).cuda()
    big_t = torch.rand(4 * 33 * 40 + 3, generator=gen(4)).cuda()
    o = big_o[3:].view(4, 1, 33, 40)
    t = big_t[1:-2].view(4, 1, 33, 40)
    assert o.data_ptr() % 16 and t.data_ptr() % 16 and o.is_contiguous()
    acc, grad = nsm.l1_loss_fwd_bwd(o, t, (), coef_l1=0.9 / o.numel())
    assert abs(acc[0].item() / o.numel() - oracle.l1_loss(o.cpu(), t.cpu()).item()) <= 1e-6
    assert torch.allclose(grad.cpu(), oracle.custom_loss_grad(o.cpu(), t.cpu(), 0.9), rtol=1e-6, atol=0)


@pytest.mark.parametrize("shape", [(1, 1, 7, 9), (3, 1, 64, 96), (2, 1, 1080, 1920)])
def test_l1_value_and_sign_gradient(nsm, shape):
    o = torch.rand(*shape, generator=gen(1))
    t = torch.rand(*shape, generator=gen(2))
    o.view(-1)[::5] = t.view(-1)[::5]                       # exact ties: sign(0) = 0
    acc, grad = nsm.l1_loss_fwd_bwd(o.cuda(), t.cuda(), (), coef_l1=0.9 / o.numel())
    assert abs(acc[0].item() / o.numel() - oracle.l1_loss(o, t).item()) <= 1e-6
    assert torch.allclose(grad.cpu(), oracle.custom_loss_grad(o, t, 0.9), rtol=1e-6, atol=0)
    assert acc[2].item() == 0


def test_range_assert(nsm):
    from customLoss import CustomLoss
    o = torch.rand(1, 1, 16, 16, device="cuda") * 1.5
    with pytest.raises(AssertionError):
        CustomLoss("cuda")(o, torch.rand(1, 1, 16, 16, device="cuda"), None)


def test_perturbation_loss_value_and_grad(nsm):
    out = torch.rand(2, 1, 32, 48, generator=gen(3))
    ys = [torch.rand(2, 1, 32, 48, generator=gen(4 + i)) for i in range(3)]
    ys[1].view(-1)[::7] = out.view(-1)[::7]
    o = out.cuda().requires_grad_(True)
    from pert_loss import _FusedPerturbL1
    total, _, pert, _ = _FusedPerturbL1.apply(o, None, 0.0, 1.0, *[y.cuda() for y in ys])
    total.backward()
    ref = sum(oracle.l1_loss(out, y) for y in ys) / 3
    assert abs(total.item() - ref.item()) <= 1e-6 and abs(pert.item() - ref.item()) <= 1e-6
    assert torch.allclose(o.grad.cpu(), oracle.perturbation_loss_grad(out, ys), rtol=1e-6, atol=0)


def test_perturb_inputs_match_reference_order(nsm):
    from pert_loss import PerturbationLoss
    x = torch.randn(2, 4, 24, 40, generator=gen(8)) * torch.tensor([1.0, 2.0, 3.0, 0.5]).view(1, 4, 1, 1)
    noises = [[torch.randn(2, 1, 24, 40, generator=gen(100 + 4 * i + c)) for c in range(4)] for i in range(3)]
    ref = oracle.perturb_inputs(x, 3, noises=noises)
    nz = torch.stack([torch.stack(n) for n in noises]).cuda()             # [count, C, B, 1, H, W]
    got = PerturbationLoss(3).perturb_input(x.cuda(), noise=nz)
    for g, r in zip(got, ref):
        assert (g.cpu() - r).abs().max().item() <= 2e-7 * 4


def test_channel_stats_match_reference_golden(nsm, golden, tmp_path):
    rng = np.random.default_rng(0)
    mu = np.array([0.1, -1.0, 5.0, 0.0], dtype=np.float32).reshape(1, 4, 1, 1)
    sg = np.array([1.0, 2.0, 3.0, 0.5], dtype=np.float32).reshape(1, 4, 1, 1)
    data = (rng.standard_normal((6, 4, 64, 96), dtype=np.float32) * sg + mu).astype(np.float32)
    np.save(tmp_path / "train_inputs.npy", data)
    from calculate_dataset_stats import calculate_dataset_stats
    st = calculate_dataset_stats(str(tmp_path))
    assert st is not None
    assert np.allclose(st["means"], golden["stats_means"], rtol=1e-6, atol=1e-7)
    assert np.allclose(st["stds"], golden["stats_stds"], rtol=1e-6)
    loaded = np.load(tmp_path / "train_stats.npy", allow_pickle=True).item()   # same file format as the reference
    assert set(loaded) == {"means", "stds"} and len(loaded["means"]) == 4
    # ragged / unaligned plane sizes
    x = torch.randn(3, 4, 13, 7, generator=gen(5))
    s = nsm.channel_sums(x.cuda()).cpu()
    assert torch.allclose(s, x.double().sum(dim=(0, 2, 3)), rtol=1e-12, atol=1e-9)


def test_device_feeder_matches_reference_standardisation(nsm, tmp_path):
    """Pinned double-buffered feed + on-GPU standardise == MmapLiverDataset.__getitem__ of the reference (setdata.py:316),
    bit for bit, in the loader's sequential order."""
    from setdata_b200 import DeviceFeeder, MmapLiverDataset
    rng = np.random.default_rng(3)
    data = (rng.standard_normal((7, 4, 32, 48), dtype=np.float32) * 3 + 1).astype(np.float32)
    labels = rng.random((7, 1, 32, 48), dtype=np.float32)
    np.save(tmp_path / "train_inputs.npy", data)
    np.save(tmp_path / "train_labels.npy", labels)
    st = oracle.channel_stats(data)
    np.save(tmp_path / "train_stats.npy", st)
    ds = MmapLiverDataset(str(tmp_path), split="train")
    assert len(ds) == 7 and torch.allclose(ds.means, torch.tensor(st["means"], dtype=torch.float32))
    seen = 0
    for xb, yb in DeviceFeeder(ds, batch_size=3):
        assert xb.is_cuda and xb.requires_grad and yb.is_cuda
        for j in range(xb.shape[0]):
            ref = oracle.standardise(torch.from_numpy(data[seen]), ds.means.tolist(), ds.stds.tolist())
            assert torch.equal(xb[j].detach().cpu(), ref)
            assert torch.equal(yb[j].cpu(), torch.from_numpy(labels[seen]))
            seen += 1
    assert seen == 7


def test_vgg_perceptual_term(nsm):
    """MultiLayerVGGLoss (customLoss.py:7-90) on the tcgen05 conv kernel: against the golden values of the unmodified
    reference class (seeded VGG19 stand-in, tests/golden/make_golden_vgg.py), against the oracle on a larger, odd-sized
    batch, and inside CustomLoss (value parity of `alpha*L1 + (1-alpha)*vgg`, customLoss.py:160,193)."""
    import os
    import sys
    import numpy as np
    import torchvision
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from vgg_fixture import cases, seeded_vgg19
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vgg_vectors.npz"))
    saved = torchvision.models.vgg19
    torchvision.models.vgg19 = seeded_vgg19
    try:
        from customLoss import CustomLoss, MultiLayerVGGLoss, vgg_weights_available
        assert vgg_weights_available()
        crit = MultiLayerVGGLoss("cuda")
        full = CustomLoss("cuda", alpha=0.9)          # vgg_loss="auto" picks the term up
        assert isinstance(full.vgg_loss, MultiLayerVGGLoss)
    finally:
        torchvision.models.vgg19 = saved
    assert [k for k in crit.state_dict()][:3] == ["weights", "mean", "std"]
    for tag, (o, t) in cases().items():
        v = crit(o.cuda(), t.cuda())
        assert v.dim() == 0 and not v.requires_grad
        want = float(gold[f"vgg_{tag}"])
        print(f"vgg {tag}: fp32 mode {v.item():.8f} reference {want:.8f}")
        assert abs(v.item() - want) <= 1e-4 * want
        with torch.autocast("cuda", dtype=torch.bfloat16):
            vb = crit(o.cuda(), t.cuda())
        wantb = float(gold[f"vgg_{tag}_bf16"])
        print(f"vgg {tag}: bf16 mode {vb.item():.8f} reference (CPU autocast bf16) {wantb:.8f}")
        assert abs(vb.item() - wantb) <= 1e-2 * wantb
    o, t = cases()["a"]
    og = o.cuda().requires_grad_(True)
    loss = full(og, t.cuda(), None)
    loss.backward()
    assert abs(loss.item() - float(gold["custom_loss_a"])) <= 1e-5
    ref_grad = oracle.custom_loss_grad(o, t, 0.9)                              # the term carries no gradient
    assert torch.equal(torch.sign(og.grad.cpu()), torch.sign(ref_grad))
    assert torch.allclose(og.grad.cpu(), ref_grad, rtol=2e-7, atol=0)
    # larger, odd-sized batch (sizes that do not divide by 16: ragged tiles, floor in every max-pool) vs stock PyTorch
    g = gen(5)
    o = torch.rand(3, 1, 150, 202, generator=g)
    t = torch.rand(3, 1, 150, 202, generator=g)
    o[0, 0, 0, 0], t[0, 0, 1, 1] = float("nan"), float("inf")                    # scrubbed like customLoss.py:48-52
    feats = seeded_vgg19().features.eval().cuda()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        want = oracle.vgg_perceptual_loss(o.cuda(), t.cuda(), feats).item()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    got = crit(o.cuda(), t.cuda()).item()
    print(f"vgg 3x150x202: {got:.8f} vs stock PyTorch strict fp32 {want:.8f}")
    assert abs(got - want) <= 1e-4 * want


def test_enhanced_custom_loss_of_customloss_py(nsm):
    """customLoss.EnhancedCustomLoss (customLoss.py:195-238): alpha * L1 + (1 - alpha) * vgg + beta * MSE(output,
    model(clamp(inputs + 0.01 * randn_like(inputs), -10, 10))) -- value, components and the gradient w.r.t. the output
    against the same formula in stock PyTorch under the same generator state (vgg term off on both sides)."""
    import torch.nn.functional as F
    from customLoss import EnhancedCustomLoss
    g = torch.Generator().manual_seed(11)
    inputs = (torch.randn(2, 4, 33, 47, generator=g) * 4.0).cuda()
    inputs[0, 0, 0, :4] = torch.tensor([11.0, -12.0, 9.995, -9.999])       # clamp is exercised
    target = torch.rand(2, 1, 33, 47, generator=g).cuda()
    w = torch.randn(1, 4, 3, 3, generator=g).cuda() * 0.3
    model = lambda x: torch.sigmoid(F.conv2d(x, w, padding=1))              # noqa: E731  stand-in network (any callable)
    crit = EnhancedCustomLoss("cuda", alpha=0.9, beta=0.05, vgg_loss=None)

    out_a = model(inputs).detach().requires_grad_(True)
    torch.manual_seed(123)
    total, comp = crit(model, out_a, target, inputs)
    total.backward()

    out_b = out_a.detach().clone().requires_grad_(True)
    torch.manual_seed(123)
    noise = torch.randn_like(inputs) * 0.01                                  # customLoss.py:225-231
    pin = torch.clamp(inputs + noise, -10.0, 10.0)
    with torch.no_grad():
        pout = model(pin)
    l1 = F.l1_loss(out_b, target)
    mse = F.mse_loss(out_b, pout)
    ref = 0.9 * l1 + 0.05 * mse
    ref.backward()

    assert set(comp) == {"l1_loss", "vgg_loss", "perturbation_loss"}
    l1, mse, ref, total = l1.detach(), mse.detach(), ref.detach(), total.detach()
    comp = {k: v.detach() for k, v in comp.items()}
    assert abs(float(comp["l1_loss"]) - float(l1)) <= 1e-6 and float(comp["vgg_loss"]) == 0.0
    assert float(mse) > 0 and abs(float(comp["perturbation_loss"]) - float(mse)) <= 1e-5 * float(mse)
    assert abs(float(total) - float(ref)) <= 1e-6
    assert torch.allclose(out_a.grad, out_b.grad, rtol=1e-5, atol=1e-9)
    # the jitter kernel itself, bit for bit (two roundings like torch: noise * eps, then the add, then clamp)
    torch.manual_seed(5)
    nz = torch.randn_like(inputs)
    assert torch.equal(nsm.add_noise_clamp(inputs, nz, 0.01, -10.0, 10.0), torch.clamp(inputs + nz * 0.01, -10.0, 10.0))
