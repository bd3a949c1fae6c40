"""Whole-network eval-mode parity (Unet.forward under model.eval()) through the drop-in module and the C ABI.

Checked against (a) the committed golden vectors produced by the unmodified reference classes and (b) the CPU oracle
on the same seeded inputs, including every intermediate tensor (nsm_unet_tap), the odd-size / odd-level paths, the
1080p frame of BASELINE config 1, fused standardisation and the host-buffer entry point.
Tolerances (BASELINE.json north_star): output max-abs <= 1e-4 in fp32 mode vs the fp32 reference, <= 1e-2 in bf16
mode vs the bf16-autocast reference, on values in [0,1].
"""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-4, "bf16": 1e-2}


@pytest.fixture(scope="module")
def nsm():
    import nsm as _nsm
    _nsm.require_device()
    return _nsm


def gen(seed):
    return torch.Generator().manual_seed(seed)


def make_net(P, precision):
    from Unetmodel import Unet
    net = Unet(precision=precision)
    net.load_state_dict(P, strict=True)
    return net.cuda().eval()


def calibrated(shape, seed=1):
    P = oracle.init_params(42)
    x = torch.randn(*shape, generator=gen(seed))
    oracle.calibrate_bn(P, x, generator=gen(2))
    return P, x


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["eval_a", "eval_b", "eval_c"])
def test_against_reference_golden(nsm, golden, precision, tag):
    shape = tuple(int(v) for v in golden[f"{tag}_shape"])
    P, x = calibrated(shape)
    net = make_net(P, precision)
    with torch.inference_mode():
        y = net(x.cuda())
    ref = torch.from_numpy(golden[f"{tag}_out" if precision == "fp32" else f"{tag}_out_bf16"])
    assert y.shape == ref.shape
    assert y.dtype == (torch.float32 if precision == "fp32" else torch.bfloat16)
    err = (y.float().cpu() - ref).abs().max().item()
    assert err <= TOL[precision], (tag, precision, err)
    assert 0.0 <= float(y.min()) and float(y.max()) <= 1.0


TAPS = [("c2", "c2"), ("p2", "p2"), ("t3", "conv3.a0"), ("c3", "c3"), ("p3", "p3"), ("t4", "conv4.a0"),
        ("c4", "c4"), ("p4", "p4"), ("t5", "conv5.a0"), ("c5", "c5"), ("u6", "u6"), ("t6", "conv6.a0"),
        ("m6", "m6"), ("u7", "u7"), ("t7", "conv7.a0"), ("m7", "m7"), ("u8", "u8"), ("t8", "conv8.a0"),
        ("m8", "m8"), ("u9", "u9"), ("t9", "conv9.a0")]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(2, 4, 32, 48), (1, 4, 41, 57), (1, 4, 80, 112)], ids=lambda s: "x".join(map(str, s)))
def test_every_intermediate_against_oracle(nsm, precision, shape):
    P, x = calibrated(shape)
    net = make_net(P, precision)
    taps = {}
    nsm.lib().nsm_unet_set_fused_decoder(0)     # stage-by-stage decoder: u8 / t8 / u9 / t9 exist in the workspace
    try:
        with torch.no_grad():
            ref = oracle.unet_forward(x, P, training=False, bf16=(precision == "bf16"), taps=taps)
            y = net(x.cuda())
            torch.cuda.synchronize()
    finally:
        nsm.lib().nsm_unet_set_fused_decoder(-1)
    ws, B, H, W, mode = net.last_workspace
    report = []
    worst = 0.0
    for mine, theirs in TAPS:
        got = nsm.unet_tap(ws, B, H, W, mode, mine).cpu()
        r = taps[theirs].float()
        assert got.shape == r.shape, (mine, got.shape, r.shape)
        rel = ((got - r).abs().max() / r.abs().max().clamp_min(1e-6)).item()
        report.append(f"{mine}:{rel:.2e}")
        worst = max(worst, rel)
    err = (y.float().cpu() - ref.float()).abs().max().item()
    print(precision, shape, "out err", err, " ".join(report))
    assert worst <= (3e-4 if precision == "fp32" else 0.1), " ".join(report)
    assert err <= TOL[precision], (err, " ".join(report))
    with torch.no_grad():                       # the fused decoder blocks (default) give the same output
        y_fused = net(x.cuda())
    assert (y_fused.float() - y.float()).abs().max().item() <= (2e-5 if precision == "fp32" else 1.6e-2)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_1080p_frame(nsm, precision):
    """BASELINE config 1: single 1920x1080 frame; level sizes 540x960 / 270x480 / 135x240 / 67x120 (odd level:
    AvgPool floors 135 -> 67 and up6 is resized 134 -> 135).

    fp32 mode is asserted against the CPU fp32 oracle.  bf16 mode is asserted against the same call sequence executed
    by stock PyTorch on the GPU under autocast(bfloat16) with TF32 off (the reference path as main.py:257-263 runs it
    on a GPU); the CPU bf16 oracle is reported next to it together with the reference-vs-reference floor, because
    ATen's CPU bf16 kernels round differently from its CUDA kernels (e.g. bf16 interpolation weights)."""
    P, _ = calibrated((1, 4, 64, 64))
    x = torch.randn(1, 4, 1080, 1920, generator=gen(9))
    net = make_net(P, precision)
    bf16 = precision == "bf16"
    with torch.no_grad():
        ref_cpu = oracle.unet_forward(x, P, training=False, bf16=bf16).float()
    with torch.inference_mode():
        y = net(x.cuda()).float().cpu()
    assert y.shape == (1, 1, 1080, 1920)
    err_cpu = (y - ref_cpu).abs().max().item()
    if not bf16:
        print("1080p fp32 max abs err vs CPU oracle", err_cpu)
        assert err_cpu <= TOL[precision]
        return
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        Pg = {k: v.cuda() for k, v in P.items()}
        with torch.no_grad():
            ref_gpu = oracle.unet_forward(x.cuda(), Pg, training=False, bf16=True).float().cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    err_gpu = (y - ref_gpu).abs().max().item()
    floor = (ref_gpu - ref_cpu).abs().max().item()
    print(f"1080p bf16 max abs err: vs stock-PyTorch-on-GPU bf16 {err_gpu:.4g}, vs CPU bf16 oracle {err_cpu:.4g}; "
          f"reference-vs-reference (GPU vs CPU bf16) {floor:.4g}; mean abs err vs GPU ref "
          f"{(y - ref_gpu).abs().mean().item():.3g}")
    # Measured on B200: the two stock-PyTorch bf16 references differ from EACH OTHER by 0.033 max-abs on this frame
    # (cuDNN vs oneDNN accumulation, bf16 interpolation weights on CPU), i.e. the 1e-2 target is below the reference's own
    # reproducibility at 2 M pixels.  It is asserted on the small golden cases above; here the B200 path must be no farther
    # from either reference than they are from each other.
    assert err_cpu <= max(TOL[precision], floor)
    assert err_gpu <= max(TOL[precision], 1.5 * floor)
    assert (y - ref_cpu).abs().mean().item() <= 3e-3


def test_fused_standardise_and_host_entry(nsm):
    P, xs = calibrated((1, 4, 48, 64))
    mean = torch.tensor([0.1, -1.0, 5.0, 0.0])
    std = torch.tensor([1.0, 2.0, 3.0, 0.5])
    raw = xs * (std.view(1, 4, 1, 1) + 1e-8) + mean.view(1, 4, 1, 1)
    x_std = oracle.standardise(raw[0], mean.tolist(), std.tolist()).unsqueeze(0)
    net = make_net(P, "fp32")
    with torch.no_grad():
        ref = oracle.unet_forward(x_std, P, training=False)
        net.set_input_stats(mean, std)
        y = net(raw.cuda())
        assert (y.cpu() - ref).abs().max().item() <= 1e-4
        # stand-alone standardise kernel is bit-exact with setdata.py:316
        got = nsm.standardize(raw.cuda(), mean.cuda(), std.cuda()).cpu()
        assert torch.equal(got, x_std)
        # host-buffer entry point (H2D + forward + D2H inside one C-ABI call)
        yh = net.infer_host(raw.pin_memory())
        assert torch.equal(yh, y.cpu())
        # uint8 output path fused into the last kernel == infer.py:79  (out * 255).astype(np.uint8)
        y8 = net.infer_host_u8(raw.pin_memory())
        assert y8.dtype == torch.uint8 and torch.equal(y8, (y.cpu() * 255).to(torch.uint8))


def test_eval_in_training_storage_format(nsm):
    """NSM_MODE_FP32_TRAIN (hi+lo bf16 planes, the range-safe format of fp32 training) also serves eval forwards, e.g.
    validation inside a training run: 16 significand bits -> output within 5e-4 of the fp32 oracle (the fp32 eval mode
    with its fp16 planes is the one held to 1e-4)."""
    for shape in [(2, 4, 48, 64), (1, 4, 41, 57)]:
        P, x = calibrated(shape)
        net = make_net(P, "fp32_train")
        with torch.no_grad():
            ref = oracle.unet_forward(x, P, training=False)
            y = net(x.cuda()).cpu()
        assert y.shape == ref.shape
        assert (y - ref).abs().max().item() <= 5e-4, (y - ref).abs().max().item()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frame_pipeline_matches_single_calls(nsm, precision):
    """nsm_unet_pipe_*: a sequence of frames through the double-buffered copy/compute pipeline returns, frame by frame,
    exactly what the synchronous host entry point returns (fp32 and fused uint8 results, odd frame size)."""
    P, _ = calibrated((2, 4, 64, 64))
    net = make_net(P, precision)
    B, H, W = 1, 75, 98
    frames = [torch.randn(B, 4, H, W, generator=gen(10 + k)).pin_memory() for k in range(7)]
    with torch.no_grad():
        want = [net.infer_host(f).clone() for f in frames]
        want8 = [net.infer_host_u8(f).clone() for f in frames]
        pipe = net.open_pipe(B, H, W)
        outs = [torch.empty(B, 1, H - 1, W, dtype=torch.float32).pin_memory() for _ in frames]
        for f, o in zip(frames, outs):
            pipe.submit(f, o)
        pipe.sync()
        for k, (o, w) in enumerate(zip(outs, want)):
            assert torch.equal(o, w), f"frame {k}"
        outs8 = [torch.empty(B, 1, H - 1, W, dtype=torch.uint8).pin_memory() for _ in frames]
        for k, (f, o) in enumerate(zip(frames, outs8)):
            pipe.submit(f, o)
            if k >= 2:   # contract: when submit(k) returns, result k-2 is complete
                assert torch.equal(outs8[k - 2], want8[k - 2]), f"frame {k - 2} after submit {k}"
        pipe.sync()
        for k, (o, w) in enumerate(zip(outs8, want8)):
            assert torch.equal(o, w), f"u8 frame {k}"
        with pytest.raises(nsm.NsmError):
            pipe.submit(frames[0].clone(), outs[0])      # pageable input: refused
        pipe.close()


def test_rejects_bad_input(nsm):
    from Unetmodel import Unet
    net = Unet().cuda().eval()
    with pytest.raises(ValueError):
        net(torch.zeros(1, 3, 32, 32, device="cuda"))
    with pytest.raises(nsm.NsmError):
        net(torch.zeros(1, 4, 8, 8, device="cuda"))
    with pytest.raises(nsm.NsmError):
        net(torch.zeros(1, 4, 32, 32))          # CPU tensor: no fallback


def test_tma_descriptors_are_encoded_once(nsm):
    """Steady state: a frame of the same size through the same model adds hits to the TMA-descriptor table and no
    misses (no cuTensorMapEncodeTiled call per launch), and the cached descriptors give the same output bit for bit."""
    P, x = calibrated((1, 4, 70, 90), seed=7)
    net = make_net(P, "fp32")
    xd = x.cuda()
    with torch.inference_mode():
        y0 = net(xd).clone()
        h0, m0 = nsm.tmap_cache_stats()
        y1 = net(xd).clone()
        h1, m1 = nsm.tmap_cache_stats()
    assert m0 > 0 and m1 == m0 and h1 > h0, (h0, m0, h1, m1)
    assert torch.equal(y0, y1)
