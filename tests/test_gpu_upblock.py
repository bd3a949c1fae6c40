"""Fused decoder block (nsm_upblock, csrc/upblock.cu): composite up-sample on the operand path -> 3x3 conv + BN + LReLU
-> 1x1 conv + BN + LReLU -> skip add | conv10 + sigmoid + pixel_shuffle, ONE launch (Unetmodel.py:134-148).

Checked (a) against the CPU oracle chain of the same reference ops and (b) against the stage-by-stage CUDA path
(nsm_upsample_match -> nsm_conv_fwd x2), whose rounding points are identical: in bf16 mode only isolated one-ulp flips
from the accumulation order may differ.
"""
import pytest
import torch
import torch.nn.functional as F

import oracle
from test_gpu_stages import bf, check_close, describe, gen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nsm():
    import nsm as _nsm
    _nsm.require_device()
    return _nsm


CASES = [
    # N, Cmid, Cout, Hs, Ws, H, W, tail
    (1, 128, 64, 8, 4, 16, 8, False),      # exactly one tile, plain x2
    (1, 128, 64, 12, 20, 24, 40, False),   # conv8-like: x2, several tiles
    (2, 128, 64, 9, 13, 19, 27, False),    # odd skip size (the 67 -> 135 case of 1080p), ragged tiles, batch 2
    (1, 64, 16, 16, 8, 16, 8, True),       # one tile, same-resolution composite (up9), tail
    (1, 64, 16, 24, 40, 24, 40, True),     # conv9-like
    (2, 64, 16, 19, 27, 19, 27, True),     # ragged, batch 2
    (1, 64, 64, 10, 14, 20, 28, False),    # 64-channel block with a skip
    (1, 128, 64, 40, 36, 80, 72, False),   # more tiles than one wave of a small grid would take per CTA
]


def _params(Cmid, Cout, g):
    w3 = torch.randn(Cmid, Cmid, 3, 3, generator=g) / (Cmid * 9) ** 0.5
    w1 = torch.randn(Cout, Cmid, 1, 1, generator=g) / Cmid ** 0.5
    vs = []
    for C in (Cmid, Cout):
        b = torch.randn(C, generator=g) * 0.1
        gamma = torch.empty(C).uniform_(0.5, 1.5, generator=g)
        beta = torch.empty(C).uniform_(-0.5, 0.5, generator=g)
        rm = torch.randn(C, generator=g) * 0.2
        rv = torch.empty(C).uniform_(0.5, 1.5, generator=g)
        vs.append((b, gamma, beta, rm, rv))
    w10 = torch.randn(4, 16, 1, 1, generator=g) * 0.4
    b10 = torch.randn(4, generator=g) * 0.1
    return w3, w1, vs, w10, b10


@pytest.mark.parametrize("mode_name", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_upblock(nsm, mode_name, case):
    N, Cmid, Cout, Hs, Ws, H, W, tail = case
    mode = nsm.MODES[mode_name]
    rbf = mode_name == "bf16"
    g = gen(hash(case) % 1000)
    src = torch.randn(N, Cmid, Hs, Ws, generator=g)
    res = None if tail else torch.randn(N, Cout, H, W, generator=g)
    if rbf:
        src, res = bf(src), (bf(res) if res is not None else None)
    w3, w1, vs, w10, b10 = _params(Cmid, Cout, g)
    (b3, g3, be3, rm3, rv3), (b1, g1, be1, rm1, rv1) = vs

    def fold(b, gamma, beta, rm, rv):
        scale = gamma / torch.sqrt(rv + 1e-5)
        return (bf(b) if rbf else b).cuda(), scale.cuda(), (beta - rm * scale).cuda()

    v3, v1 = fold(*vs[0]), fold(*vs[1])
    w10d = (bf(w10) if rbf else w10).reshape(4, 16).contiguous().cuda()
    b10d = (bf(b10) if rbf else b10).cuda()

    # ---- (a) CPU oracle chain (fp32 mode only: in bf16 mode the product rounds the composite resize once)
    srct = nsm.PlaneTensor.from_nchw(src.cuda(), mode)
    rest = nsm.PlaneTensor.from_nchw(res.cuda(), mode) if res is not None else None
    w3p = nsm.pack_conv_weight(w3.cuda(), nsm.FMT_F16_X8 if mode_name == "fp32" else mode)
    w1p = nsm.pack_conv_weight(w1.cuda(), mode)
    got = nsm.upblock(srct, H, W, w3p, w1p, Cout, v3, v1, residual=rest, tail=(w10d, b10d) if tail else None,
                      want_u8=tail)
    torch.cuda.synchronize()
    if not rbf:
        u = oracle.upsample_and_match(src, (H, W))
        t, _ = oracle.conv_stage_eval(u, w3, b3, rm3, rv3, g3, be3)
        o, _ = oracle.conv_stage_eval(t, w1, b1, rm1, rv1, g1, be1, residual=res)
        if tail:
            ref = torch.sigmoid(F.pixel_shuffle(F.conv2d(o, w10, b10), 2))
            err = (got[0].cpu() - ref).abs().max().item()
            assert err <= 2e-5, f"tail output vs oracle: {err}\n" + describe(got[0].cpu() - ref, ref)
        else:
            tol = 6e-5 * max(1.0, o.abs().max().item())       # two GEMMs, 8-bit cross operands in the first
            e = got.to_nchw().cpu() - o
            assert e.abs().max().item() <= tol, "block output vs oracle\n" + describe(e, o)

    # ---- (b) the stage-by-stage CUDA path on the same operands
    fmt3 = nsm.FMT_F16_X8 if mode_name == "fp32" else mode
    u_t = nsm.upsample_match(srct, H, W, out_x8=(mode_name == "fp32"))
    t_t, _, _ = nsm.conv_fwd(u_t, w3p, 3, Cmid, fmt3, bias=v3[0], bn_scale=v3[1], bn_shift=v3[2], lrelu=True)
    if not tail:
        o_t, _, _ = nsm.conv_fwd(t_t, w1p, 1, Cout, mode, bias=v1[0], bn_scale=v1[1], bn_shift=v1[2], lrelu=True,
                                 residual=rest)
        a, b = got.to_nchw().cpu(), o_t.to_nchw().cpu()
        if rbf:
            # same rounding points, different accumulation order (chunk-major halo K loop): isolated one-ulp flips of the bf16
            # intermediate propagate through the 1x1 GEMM and the skip add (where they may cancel down to a small value)
            e = (a - b).abs()
            m = b.abs().max().item()
            assert e.max().item() <= 2.0 ** -6 * m and e.mean().item() <= 1e-3 * m and (e != 0).float().mean() < 0.2, \
                "fused block vs staged path [bf16]\n" + describe(a - b, b)
        else:
            check_close(a, b, mode_name, "fused block vs staged path")
    else:
        t_ref = t_t.to_nchw().cpu()
        with torch.autocast("cpu", dtype=torch.bfloat16) if rbf else torch.autocast("cpu", enabled=False):
            o = F.leaky_relu(F.batch_norm(F.conv2d(t_ref, w1, b1), rm1, rv1, g1, be1, False, 0.1, 1e-5), 0.2)
            ref = torch.sigmoid(F.pixel_shuffle(F.conv2d(o, w10, b10), 2)).float()
        y, y8 = got
        err = (y.cpu() - ref).abs().max().item()
        assert err <= (1.6e-2 if rbf else 2e-5), f"tail vs staged 3x3 + torch 1x1/conv10: {err}"
        assert y8 is not None and torch.equal(y8.cpu(), (y.cpu() * 255).to(torch.uint8))


def test_fused_decoder_matches_staged_network(nsm, monkeypatch):
    """Whole eval forward: fused decoder blocks (default) against the oracle at an odd-level size."""
    from Unetmodel import Unet
    assert nsm.lib().nsm_unet_fused_decoder() == 1
    P = oracle.init_params(42)
    g = gen(3)
    x = torch.randn(1, 4, 88, 120, generator=g)
    oracle.calibrate_bn(P, x, generator=g)
    ref = oracle.unet_forward(x, P, training=False)
    net = Unet(precision="fp32")
    net.load_state_dict(P)
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(x.cuda()).float().cpu()
    err = (y - ref).abs().max().item()
    assert err <= 1e-4, err
