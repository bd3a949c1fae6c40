"""Training-mode forward / backward of the drop-in Unet on B200 (Unetmodel.py:90-149 under model.train() and its
autograd backward, main.py:263-281).

One ``torch.autograd.Function`` spans the whole network: forward launches the sm_100a kernels stage by stage and keeps
the raw convolution outputs (pre-BatchNorm) plus the per-channel BN statistics; backward produces the gradients of all
66 parameters (and of the input when it requires grad, as the reference dataset sets it, setdata.py:325-326) in one go.
PyTorch contributes only allocation, the Dropout2d random draws (same generator calls as ``F.dropout2d``) and the
int64 ``num_batches_tracked`` counters.

Stage graph per DoubleConv block (all arithmetic in libnsm_b200.so):
  forward : conv3x3 GEMM (+bias) -> bn_stats -> bn_finalize -> bn_act(BN, LeakyReLU, Dropout2d mask)
            conv1x1 GEMM (+bias) -> bn_stats -> bn_finalize -> bn_act(BN, LeakyReLU [, +skip] [, AvgPool2d])
  backward: bn_bwd (LeakyReLU', mask, BN backward, bias grad) -> wgrad GEMM -> dgrad GEMM (x2)
Thin layers (16 / 4 channels) are zero-padded to 64 channels so that every convolution runs on the tensor cores.
"""
from __future__ import annotations

import torch

import nsm

BLOCKS = (("conv2", 16, 64), ("conv3", 64, 128), ("conv4", 128, 512), ("conv5", 512, 1024),
          ("conv6", 1024, 512), ("conv7", 512, 128), ("conv8", 128, 64), ("conv9", 64, 16))


def _pad64(c):
    return max(64, c)


def _param_list(model):
    ps = []
    for name, _, _ in BLOCKS:
        seq = getattr(model, name).conv
        for idx in (0, 1, 4, 5):
            ps += [seq[idx].weight, seq[idx].bias]
    return ps + [model.conv10.weight, model.conv10.bias]


class _Packed:
    """Per-step packed operands: forward and dgrad weight planes, padded bias / BN vectors."""

    def __init__(self, model, mode):
        rb = mode == nsm.MODE_BF16
        self.blocks = []
        for name, cin, cout in BLOCKS:
            seq = getattr(model, name).conv
            cip, cop = _pad64(cin), _pad64(cout)
            d = {}
            d["w3"] = nsm.pack_conv_weight_padded(seq[0].weight, mode, cip, cip)
            d["w3t"] = nsm.pack_conv_weight_padded(seq[0].weight, mode, cip, cip, dgrad=True)
            d["b3"] = nsm.pad_vector(seq[0].bias, cip, 0.0, rb)
            d["g3"] = nsm.pad_vector(seq[1].weight, cip, 0.0)
            d["be3"] = nsm.pad_vector(seq[1].bias, cip, 0.0)
            d["w1"] = nsm.pack_conv_weight_padded(seq[4].weight, mode, cop, cip)
            d["w1t"] = nsm.pack_conv_weight_padded(seq[4].weight, mode, cop, cip, dgrad=True)
            d["b1"] = nsm.pad_vector(seq[4].bias, cop, 0.0, rb)
            d["g1"] = nsm.pad_vector(seq[5].weight, cop, 0.0)
            d["be1"] = nsm.pad_vector(seq[5].bias, cop, 0.0)
            self.blocks.append(d)
        self.w10 = nsm.pack_conv_weight_padded(model.conv10.weight, mode, 64, 64)
        self.w10t = nsm.pack_conv_weight_padded(model.conv10.weight, mode, 64, 64, dgrad=True)
        self.b10 = nsm.pad_vector(model.conv10.bias, 64, 0.0, rb)


def _packed(model, mode):
    ps = _param_list(model)
    key = (mode,) + tuple((p.data_ptr(), p._version) for p in ps)
    hit = getattr(model, "_train_packed", None)
    if hit is None or hit[0] != key:
        model._train_packed = (key, _Packed(model, mode))
    return model._train_packed[1]


def _running(bn, cpad):
    """Padded copies of the running statistics (only the 16-channel BNs need padding)."""
    c = bn.running_mean.numel()
    if c == cpad:
        return bn.running_mean, bn.running_var, None
    rm = torch.zeros(cpad, dtype=torch.float32, device=bn.running_mean.device)
    rv = torch.ones(cpad, dtype=torch.float32, device=bn.running_mean.device)
    rm[:c].copy_(bn.running_mean)
    rv[:c].copy_(bn.running_var)
    return rm, rv, c


def _conv_stats(x, w, ksize, cout, mode, bias):
    """Raw convolution (+bias) with the BatchNorm batch statistics accumulated by the conv epilogue itself."""
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=x.p0.device)
    z, _, _ = nsm.conv_fwd(x, w, ksize, cout, mode, bias=bias, stats=sums)
    return z, sums


def _bn_train(z, gamma, beta, bn, updates=1, sums=None):
    """Batch statistics + running-stat update of one nn.BatchNorm2d (eps 1e-5, momentum 0.1).  Returns the [4, C]
    (scale, shift, mean, invstd) tensor and the fp64 sums (kept for the checkpoint replay of conv5)."""
    N, C, H, W = z.shape
    if sums is None:
        sums = nsm.bn_stats(z)
    rm, rv, real = _running(bn, C)
    st = nsm.bn_finalize(sums, N * H * W, gamma, beta, rm, rv, updates=updates, eps=bn.eps,
                         momentum=bn.momentum if bn.momentum is not None else 0.1)
    if real is not None:
        bn.running_mean.copy_(rm[:real])
        bn.running_var.copy_(rv[:real])
    bn.num_batches_tracked += updates
    return st, sums


def _bn_replay(sums, P, gamma, beta, bn):
    """Second running-stat update of the reference's checkpoint(conv5) recomputation (runs during backward)."""
    rm, rv, real = _running(bn, gamma.numel())
    nsm.bn_finalize(sums, P, gamma, beta, rm, rv, updates=1, eps=bn.eps,
                    momentum=bn.momentum if bn.momentum is not None else 0.1)
    if real is not None:
        bn.running_mean.copy_(rm[:real])
        bn.running_var.copy_(rv[:real])
    bn.num_batches_tracked += 1


def _draw_masks(model, N, mode, device):
    """The eight Dropout2d draws in block order, exactly as F.dropout2d / feature_dropout makes them
    (noise = empty(N, C, 1, 1).bernoulli_(1 - p).div_(1 - p) in the activation dtype)."""
    # parity tests replay given draws: set by `replay_masks(model, masks)` for exactly ONE forward (popped here, so a
    # model object reused afterwards draws fresh masks again)
    replay = model.__dict__.pop("_replay_masks", None)
    dt = torch.bfloat16 if mode == nsm.MODE_BF16 else torch.float32
    masks = []
    for i, (name, cin, _) in enumerate(BLOCKS):
        p = getattr(model, name).conv[3].p
        if replay is not None:
            m = replay[i]
            m = None if m is None else m.to(device=device, dtype=dt)
        elif p > 0:
            m = torch.empty(N, cin, 1, 1, dtype=dt, device=device).bernoulli_(1 - p).div_(1 - p)
        else:
            m = None
        if m is not None:
            full = torch.zeros(N, _pad64(cin), dtype=torch.float32, device=device)
            full[:, :cin] = m.reshape(N, cin).to(torch.float32)
            m = full
        masks.append(m)
    return masks


def replay_masks(model, masks):
    """Parity-test hook: the NEXT training-mode forward of `model` uses these eight Dropout2d masks ([N,Cin,1,1], already
    divided by 1-p, block order conv2..conv9; None = no dropout for that block) instead of drawing from the CUDA generator.
    One-shot: the forward consumes them."""
    model.__dict__["_replay_masks"] = list(masks)


def _block_forward(model, pk, i, x, mode, mask, residual=None, pool=False, save=True):
    name, cin, cout = BLOCKS[i]
    seq = getattr(model, name).conv
    d = pk.blocks[i]
    cip, cop = _pad64(cin), _pad64(cout)
    z0, s0 = _conv_stats(x, d["w3"], 3, cip, mode, d["b3"])
    st0, sums0 = _bn_train(z0, d["g3"], d["be3"], seq[1], sums=s0)
    a0, _ = nsm.bn_act(z0, st0[0], st0[1], mask=mask, lrelu=True)
    z1, s1 = _conv_stats(a0, d["w1"], 1, cop, mode, d["b1"])
    st1, sums1 = _bn_train(z1, d["g1"], d["be1"], seq[5], sums=s1)
    y, pooled = nsm.bn_act(z1, st1[0], st1[1], mask=None, lrelu=True, residual=residual, pool=pool)
    saved = dict(x=x, z0=z0, a0=a0, z1=z1, st0=st0, st1=st1, mask=mask, sums0=sums0, sums1=sums1) if save else None
    return y, pooled, saved


def _block_backward(model, pk, i, sv, dy, mode, need_dx=True):
    """dy: gradient w.r.t. the block output (after the final LeakyReLU).  Returns (dx planes | None, grads dict)."""
    name, cin, cout = BLOCKS[i]
    d = pk.blocks[i]
    st0, st1 = sv["st0"], sv["st1"]
    g = {}
    dz1, dg1, dbe1, db1 = nsm.bn_bwd(dy, sv["z1"], st1[0], st1[1], st1[2], st1[3], mask=None, lrelu=True)
    g[f"{name}.conv.5.weight"], g[f"{name}.conv.5.bias"], g[f"{name}.conv.4.bias"] = dg1[:cout], dbe1[:cout], db1[:cout]
    g[f"{name}.conv.4.weight"] = nsm.wgrad(dz1, sv["a0"], 1, cout, cin)
    da0, _, _ = nsm.conv_fwd(dz1, d["w1t"], 1, _pad64(cin), mode)
    dz0, dg0, dbe0, db0 = nsm.bn_bwd(da0, sv["z0"], st0[0], st0[1], st0[2], st0[3], mask=sv["mask"], lrelu=True)
    g[f"{name}.conv.1.weight"], g[f"{name}.conv.1.bias"], g[f"{name}.conv.0.bias"] = dg0[:cin], dbe0[:cin], db0[:cin]
    g[f"{name}.conv.0.weight"] = nsm.wgrad(dz0, sv["x"], 3, cin, cin)
    dx = nsm.conv_fwd(dz0, d["w3t"], 3, _pad64(cin), mode)[0] if need_dx else None
    return dx, g


def _forward(model, x, mode, save):
    dev = x.device
    N, _, Hin, Win = x.shape
    pk = _packed(model, mode)
    masks = _draw_masks(model, N, mode, dev)
    st = {"pk": pk, "in_shape": (N, Hin, Win)} if save else None
    x16 = nsm.train_input_prep(x.detach().to(torch.float32).contiguous(), mode)
    sv = [None] * 8
    c2, p2, sv[0] = _block_forward(model, pk, 0, x16, mode, masks[0], pool=True, save=save)
    c3, p3, sv[1] = _block_forward(model, pk, 1, p2, mode, masks[1], pool=True, save=save)
    c4, p4, sv[2] = _block_forward(model, pk, 2, p3, mode, masks[2], pool=True, save=save)
    c5, _, sv[3] = _block_forward(model, pk, 3, p4, mode, masks[3], save=save)
    u6 = nsm.upsample_match(c5, c4.shape[2], c4.shape[3])
    m6, _, sv[4] = _block_forward(model, pk, 4, u6, mode, masks[4], residual=c4, save=save)
    u7 = nsm.upsample_match(m6, c3.shape[2], c3.shape[3])
    m7, _, sv[5] = _block_forward(model, pk, 5, u7, mode, masks[5], residual=c3, save=save)
    u8 = nsm.upsample_match(m7, c2.shape[2], c2.shape[3])
    m8, _, sv[6] = _block_forward(model, pk, 6, u8, mode, masks[6], residual=c2, save=save)
    u9 = nsm.upsample_match(m8, x16.shape[2], x16.shape[3])
    c9, _, sv[7] = _block_forward(model, pk, 7, u9, mode, masks[7], save=save)
    c10, _, _ = nsm.conv_fwd(c9, pk.w10, 1, 64, mode, bias=pk.b10)
    y = nsm.sigmoid_shuffle_fwd(c10)
    if save:
        st.update(sv=sv, c9=c9, y=y,
                  sizes=dict(c5=c5.shape[2:], m6=m6.shape[2:], m7=m7.shape[2:], m8=m8.shape[2:],
                             c4=c4.shape, c3=c3.shape, c2=c2.shape))
    return y, st


def _backward(model, st, dy, mode, need_dx):
    pk, sv = st["pk"], st["sv"]
    sync = getattr(model, "_grad_sync", None)      # parallel.GradSync: bucketed all-reduce overlapped with backward

    class _G(dict):
        def update(self, g):                        # every finished block is handed to the gradient synchroniser
            super().update(g)
            if sync is not None:
                sync.reduce_ready(g)

    G = _G()
    dc10 = nsm.sigmoid_shuffle_bwd(dy, st["y"], mode)
    G.update({"conv10.weight": nsm.wgrad(dc10, st["c9"], 1, 4, 16),
              "conv10.bias": nsm.bn_stats(dc10)[:4].to(torch.float32)})
    dc9, _, _ = nsm.conv_fwd(dc10, pk.w10t, 1, 64, mode)
    du9, g = _block_backward(model, pk, 7, sv[7], dc9, mode); G.update(g)
    dm8 = nsm.upsample_match_bwd(du9, *st["sizes"]["m8"])
    du8, g = _block_backward(model, pk, 6, sv[6], dm8, mode); G.update(g)      # skip: d c2 += dm8
    dm7 = nsm.upsample_match_bwd(du8, *st["sizes"]["m7"])
    du7, g = _block_backward(model, pk, 5, sv[5], dm7, mode); G.update(g)      # skip: d c3 += dm7
    dm6 = nsm.upsample_match_bwd(du7, *st["sizes"]["m6"])
    du6, g = _block_backward(model, pk, 4, sv[4], dm6, mode); G.update(g)      # skip: d c4 += dm6
    dc5 = nsm.upsample_match_bwd(du6, *st["sizes"]["c5"])
    dp4, g = _block_backward(model, pk, 3, sv[3], dc5, mode); G.update(g)
    # the reference's checkpoint(conv5) re-runs conv5 here: its two BatchNorms see the batch a second time
    seq5 = model.conv5.conv
    P5 = sv[3]["z0"].shape[0] * sv[3]["z0"].shape[2] * sv[3]["z0"].shape[3]
    _bn_replay(sv[3]["sums0"], P5, pk.blocks[3]["g3"], pk.blocks[3]["be3"], seq5[1])
    _bn_replay(sv[3]["sums1"], P5, pk.blocks[3]["g1"], pk.blocks[3]["be1"], seq5[5])
    dc4 = nsm.pool_bwd_add(dm6, dp4, st["sizes"]["c4"])
    dp3, g = _block_backward(model, pk, 2, sv[2], dc4, mode); G.update(g)
    dc3 = nsm.pool_bwd_add(dm7, dp3, st["sizes"]["c3"])
    dp2, g = _block_backward(model, pk, 1, sv[1], dc3, mode); G.update(g)
    dc2 = nsm.pool_bwd_add(dm8, dp2, st["sizes"]["c2"])
    dx16, g = _block_backward(model, pk, 0, sv[0], dc2, mode, need_dx=need_dx); G.update(g)
    dx = None
    if need_dx:
        N, Hin, Win = st["in_shape"]
        dx = nsm.train_input_grad(dx16, Hin, Win)
    if sync is not None:
        sync.flush()
    model._packed.clear()      # _bn_replay moved conv5's running statistics after the forward's clear
    return dx, G


class _UnetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, mode, *params):
        y, st = _forward(model, x, mode, save=True)
        ctx.model, ctx.mode, ctx.st = model, mode, st
        ctx.need_dx = x.requires_grad
        ctx.names = [n for n, _ in model.named_parameters()]
        return y

    @staticmethod
    def backward(ctx, dy):
        dx, G = _backward(ctx.model, ctx.st, dy.contiguous(), ctx.mode, ctx.need_dx)
        ctx.st = None
        grads = []
        for n, need in zip(ctx.names, ctx.needs_input_grad[3:]):
            grads.append(G[n].contiguous() if need else None)
        return (dx, None, None) + tuple(grads)


def unet_train_forward(model, x, mode):
    """Entry point used by Unet.forward when model.training is True."""
    tmode = nsm.MODE_BF16 if mode == nsm.MODE_BF16 else nsm.MODE_FP32_TRAIN
    # the BN kernels update running_mean / running_var through raw pointers (no Tensor._version bump): drop the eval-mode
    # packed blob (it folds those buffers) so that the next model.eval() forward re-packs
    model._packed.clear()
    params = [p for _, p in model.named_parameters()]
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    if needs_grad:
        y = _UnetTrainFn.apply(x, model, tmode, *params)
    else:
        with torch.no_grad():
            y, _ = _forward(model, x, tmode, save=False)
    dt = model._out_dtype(tmode)
    return y if dt == torch.float32 else y.to(dt)
