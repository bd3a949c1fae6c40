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

import os

import torch

import nsm

_NO_PX4 = bool(os.environ.get("NSM_NO_PX4"))     # A/B switch: zero-pad the thin layers to 64 channels instead

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


class _Conv:
    """One convolution of the training path with its packed operands (forward and dgrad weight planes, bias).

    Layers with fewer than 64 input or output channels (conv2, conv9 1x1, conv10) run PIXEL-PACKED when the level width is
    a multiple of 4 (``px``): four horizontally adjacent pixels x C channels are one pixel of 4C virtual channels, the
    weight becomes block-diagonal (1x1) / banded (3x3) -- tensors keep their real channel counts (``cin_s`` / ``cout_s``)
    and bytes (see nsm_b200.h, nsm_pack_conv_weight_px4).  Otherwise thin layers are zero-padded to 64 channels."""

    def __init__(self, conv, mode, px4):
        rb = mode == nsm.MODE_BF16
        self.cout, self.cin, self.k = conv.weight.shape[0], conv.weight.shape[1], conv.weight.shape[2]
        self.px = bool(px4) and (self.cin < 64 or self.cout < 64)
        if self.px:
            self.cin_v, self.cout_v = max(64, 4 * self.cin), max(64, 4 * self.cout)
            self.cin_s, self.cout_s = self.cin, self.cout
            self.w = nsm.pack_conv_weight_px4(conv.weight, mode, self.cout_v, self.cin_v)
            self.wt = nsm.pack_conv_weight_px4(conv.weight, mode, self.cout_v, self.cin_v, dgrad=True)
            self.b = nsm.tile_vector(conv.bias, 4, self.cout_v, 0.0, rb)
        else:
            self.cin_v = self.cin_s = _pad64(self.cin)
            self.cout_v = self.cout_s = _pad64(self.cout)
            self.w = nsm.pack_conv_weight_padded(conv.weight, mode, self.cout_v, self.cin_v)
            self.wt = nsm.pack_conv_weight_padded(conv.weight, mode, self.cout_v, self.cin_v, dgrad=True)
            self.b = nsm.pad_vector(conv.bias, self.cout_v, 0.0, rb)

    def _in(self, x, c_v):      # stored tensor -> the view the GEMM reads (c_v virtual channels)
        return x if (not self.px or x.shape[1] == c_v) else x.px_view(4)

    def _stored(self, t, c_s):   # GEMM result (virtual channels) -> stored tensor with c_s channels per pixel
        if not self.px or t.shape[1] == c_s:
            return t
        return t.px_view(-(t.shape[1] // c_s)) if t.shape[1] == 4 * c_s else t   # conv10: 16 of 64 virtual channels stay

    def forward(self, x, mode, stats=True):
        """Raw convolution + bias; returns (z with cout_s channels per pixel, [2*cout_s] accumulator slots | None)."""
        sums = nsm.acc_zeros(2 * self.cout_v, x.p0.device) if stats else None
        z, _, _ = nsm.conv_fwd(self._in(x, self.cin_v), self.w, self.k, self.cout_v, mode, bias=self.b, stats=sums)
        if self.px and stats:
            sums = nsm.fold_channel_sums(sums, 2, 4, self.cout)
        return self._stored(z, self.cout_s), sums

    def dgrad(self, dz, mode):
        dx, _, _ = nsm.conv_fwd(self._in(dz, self.cout_v), self.wt, self.k, self.cin_v, mode)
        return self._stored(dx, self.cin_s)

    def wgrad(self, dz, x):
        if not self.px:
            return nsm.wgrad(dz, x, self.k, self.cout, self.cin)
        dwv = nsm.wgrad(self._in(dz, self.cout_v), self._in(x, self.cin_v), self.k, self.cout_v, self.cin_v)
        return nsm.px4_reduce_dw(dwv, self.cout, self.cin, self.k)


class _Packed:
    """Per-step packed operands of all 17 convolutions and the padded BN vectors."""

    def __init__(self, model, mode, px4):
        self.px4 = px4
        self.blocks = []
        for name, cin, cout in BLOCKS:
            seq = getattr(model, name).conv
            d = {"c3": _Conv(seq[0], mode, px4), "c1": _Conv(seq[4], mode, px4)}
            d["g3"] = nsm.pad_vector(seq[1].weight, d["c3"].cout_s, 0.0)
            d["be3"] = nsm.pad_vector(seq[1].bias, d["c3"].cout_s, 0.0)
            d["g1"] = nsm.pad_vector(seq[5].weight, d["c1"].cout_s, 0.0)
            d["be1"] = nsm.pad_vector(seq[5].bias, d["c1"].cout_s, 0.0)
            self.blocks.append(d)
        self.c10 = _Conv(model.conv10, mode, px4)


def _packed(model, mode, px4):
    ps = _param_list(model)
    key = (mode, px4) + tuple((p.data_ptr(), p._version) for p in ps)
    hit = getattr(model, "_train_packed", None)
    if hit is None or hit[0] != key:
        model._train_packed = (key, _Packed(model, mode, px4))
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


def _bn_train(z, gamma, beta, bn, updates=1, sums=None):
    """Batch statistics + running-stat update of one nn.BatchNorm2d (eps 1e-5, momentum 0.1).  Returns the [4, C]
    (scale, shift, mean, invstd) tensor and the accumulator slots of the sums (kept for the checkpoint replay of conv5)."""
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


def _draw_masks(model, N, mode, device, cpads):
    """The eight Dropout2d draws in block order, exactly as F.dropout2d / feature_dropout makes them
    (noise = empty(N, C, 1, 1).bernoulli_(1 - p).div_(1 - p) in the activation dtype).  cpads[i] = channels per pixel of
    block i's stored 3x3 output (the real count, or 64 when a thin layer runs zero-padded)."""
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
            cs = cpads[i]
            if cs == cin:
                m = m.reshape(N, cin).to(torch.float32).contiguous()
            else:
                full = torch.zeros(N, cs, dtype=torch.float32, device=device)
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
    z0, s0 = d["c3"].forward(x, mode)
    st0, sums0 = _bn_train(z0, d["g3"], d["be3"], seq[1], sums=s0)
    a0, _ = nsm.bn_act(z0, st0[0], st0[1], mask=mask, lrelu=True)
    z1, s1 = d["c1"].forward(a0, mode)
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
    g[f"{name}.conv.4.weight"] = d["c1"].wgrad(dz1, sv["a0"])
    da0 = d["c1"].dgrad(dz1, mode)
    dz0, dg0, dbe0, db0 = nsm.bn_bwd(da0, sv["z0"], st0[0], st0[1], st0[2], st0[3], mask=sv["mask"], lrelu=True)
    g[f"{name}.conv.1.weight"], g[f"{name}.conv.1.bias"], g[f"{name}.conv.0.bias"] = dg0[:cin], dbe0[:cin], db0[:cin]
    g[f"{name}.conv.0.weight"] = d["c3"].wgrad(dz0, sv["x"])
    dx = d["c3"].dgrad(dz0, mode) if need_dx else None
    return dx, g


def _forward(model, x, mode, save):
    dev = x.device
    N, _, Hin, Win = x.shape
    # thin layers run pixel-packed (real channel counts) when the first level's width is a multiple of 4
    px4 = ((Win - Win % 2) // 2) % 4 == 0 and not _NO_PX4
    pk = _packed(model, mode, px4)
    masks = _draw_masks(model, N, mode, dev, [d["c3"].cout_s for d in pk.blocks])
    st = {"pk": pk, "in_shape": (N, Hin, Win)} if save else None
    x16 = nsm.train_input_prep(x.detach().to(torch.float32).contiguous(), mode, c16=px4)
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
    c10, _ = pk.c10.forward(c9, mode, stats=False)
    y = nsm.sigmoid_shuffle_fwd(c10, px4=pk.c10.px)
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
    dc10 = nsm.sigmoid_shuffle_bwd(dy, st["y"], mode, px4=pk.c10.px)
    db10 = nsm.bn_stats(dc10)                       # per-channel sums of dc10 = the bias gradient
    if pk.c10.px:
        db10 = nsm.fold_channel_sums(db10, 2, 4, 4)
    G.update({"conv10.weight": pk.c10.wgrad(dc10, st["c9"]), "conv10.bias": nsm.acc_to_double(db10[:4]).to(torch.float32)})
    dc9 = pk.c10.dgrad(dc10, mode)
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
