"""Drop-in replacement for the reference ``Unetmodel.py`` (SDU-Gary/PCSS-Unet) on B200.

Same public surface as the reference module (Unetmodel.py:17-33 ``DoubleConv``, :36-149 ``Unet``): identical
constructor signature, identical module tree -- hence identical ``state_dict()`` keys (114 entries), parameter order
and default initialisation under the same seed -- and the same ``forward(x:[B,4,H,W]) -> [B,1,H-H%2,W-W%2]``
contract, so ``main.py:894``, ``infer.py:34``, ``inference.py:256`` and ``validate_consistency.py:151`` keep working
unchanged with this directory ahead of the reference on ``sys.path``.

Only ``forward`` differs: the nn.Modules are parameter containers, the arithmetic runs in the hand-written sm_100a
kernels of ``libnsm_b200.so`` (tcgen05 implicit-GEMM convolutions with fused BatchNorm/LeakyReLU/skip/pool epilogues,
warp-level tensor-core head/tail stages, bilinear up-sampling).  There is no PyTorch or CPU fallback: a CPU tensor raises.

Precision ("mode"):
  * ``fp32`` (default outside autocast): fp32-accurate tensor-core arithmetic on hi+lo half-precision planes (fp16
    pairs in eval, bf16 pairs in training) with chunked fp32 accumulation, output within 1e-4 of the fp32 reference;
  * ``bf16`` (default under ``torch.autocast``): bf16 storage with the rounding points of the autocast reference.
Force one with ``Unet(..., precision="bf16")`` or the ``NSM_PRECISION`` environment variable.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

import nsm

_BLOCKS = (("conv2", 16, 64), ("conv3", 64, 128), ("conv4", 128, 512), ("conv5", 512, 1024),
           ("conv6", 1024, 512), ("conv7", 512, 128), ("conv8", 128, 64), ("conv9", 64, 16))


class DoubleConv(nn.Module):
    """Parameter container with the reference's layout: ``conv`` = Sequential(Conv3x3(in,in) [0], BatchNorm2d [1],
    LeakyReLU [2], Dropout2d [3], Conv1x1(in,out) [4], BatchNorm2d [5], LeakyReLU [6]).  ``dilation`` is accepted and
    ignored, as in the reference."""

    def __init__(self, in_ch, out_ch, dropout_rate=0.2, dilation=1):
        super().__init__()
        layers = [nn.Conv2d(in_ch, in_ch, 3, padding=1), nn.BatchNorm2d(in_ch, eps=1e-5, momentum=0.1),
                  nn.LeakyReLU(0.2), nn.Dropout2d(p=dropout_rate),
                  nn.Conv2d(in_ch, out_ch, 1), nn.BatchNorm2d(out_ch, eps=1e-5, momentum=0.1), nn.LeakyReLU(0.2)]
        self.conv = nn.Sequential(*layers)

    def forward(self, x):  # pragma: no cover - the fused network never calls block modules individually
        raise nsm.NsmError("DoubleConv is a parameter container here; call the enclosing Unet "
                           "(the B200 path runs whole fused stages, not per-module forwards)")

    def tensors(self):
        """The 12 tensors nsm_unet_pack consumes for this block, in header order."""
        c0, b0, c1, b1 = self.conv[0], self.conv[1], self.conv[4], self.conv[5]
        return [c0.weight, c0.bias, b0.weight, b0.bias, b0.running_mean, b0.running_var,
                c1.weight, c1.bias, b1.weight, b1.bias, b1.running_mean, b1.running_var]


class Unet(nn.Module):
    def __init__(self, in_ch=4, out_ch=1, dropout_rate=0.2, precision=None):
        super().__init__()
        # in_ch / out_ch are accepted and ignored exactly like the reference (hard-wired 4 -> 16 ... 4 -> 1)
        drops = {"conv9": dropout_rate / 2}
        for i, (name, cin, cout) in enumerate(_BLOCKS):
            setattr(self, name, DoubleConv(cin, cout, drops.get(name, dropout_rate)))
            if name in ("conv2", "conv3", "conv4"):
                setattr(self, "pool" + name[-1], nn.AvgPool2d(2))
            nxt = _BLOCKS[i + 1][0] if i + 1 < len(_BLOCKS) else None
            if nxt in ("conv6", "conv7", "conv8", "conv9"):
                setattr(self, "up" + nxt[-1], nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True))
        self.conv10 = nn.Conv2d(16, 4, 1)
        self.precision = precision or os.environ.get("NSM_PRECISION") or None
        self.input_stats = None      # optional (mean[4], std[4]) fused into the first load (setdata.py:316)
        self._packed = {}            # mode -> (version key, blob)
        self._ws = {}                # (B,H,W,mode) -> workspace
        self.last_workspace = None   # (ws, B, H, W, mode) of the latest forward, for nsm.unet_tap

    # -- same helpers as the reference keeps public -------------------------------------------------------------
    def rearrange_to_channels(self, x):
        return torch.nn.functional.pixel_unshuffle(x, downscale_factor=2)

    def reconstruct_from_channels(self, x):
        return torch.nn.functional.pixel_shuffle(x, upscale_factor=2)

    # -- plumbing -----------------------------------------------------------------------------------------------
    def _mode(self):
        """Arithmetic mode of this call.  Outside autocast: fp32 mode.  Under ``autocast(bfloat16)``: bf16 mode.  Under
        ``autocast(float16)`` -- what main.py:257 / infer.py:64 / inference.py:188 use on CUDA -- eval forwards run the
        fp32 mode (at least the reference's fp16 accuracy; its 2^-11 output step matters for infer.py:79's 8-bit
        quantisation) and training forwards the bf16 tensor path (same tensor-core rate as fp16, and GradScaler's 2^16
        loss scale cannot overflow bf16 planes)."""
        if self.precision is not None:
            return nsm.MODES[self.precision]
        if not torch.is_autocast_enabled():
            return nsm.MODE_FP32
        if torch.get_autocast_dtype("cuda") == torch.float16 and not self.training:
            return nsm.MODE_FP32
        return nsm.MODE_BF16

    @staticmethod
    def _out_dtype(mode):
        """dtype the reference would return: the autocast dtype under autocast (Unetmodel.py:148 runs inside it), else
        fp32; a forced bf16 mode outside autocast keeps returning bf16."""
        if torch.is_autocast_enabled():
            return torch.get_autocast_dtype("cuda")
        return torch.bfloat16 if mode == nsm.MODE_BF16 else torch.float32

    def _all_tensors(self):
        ts = []
        for name, _, _ in _BLOCKS:
            ts += getattr(self, name).tensors()
        return ts + [self.conv10.weight, self.conv10.bias]

    def _packed_blob(self, mode):
        ts = self._all_tensors()
        key = tuple((t.data_ptr(), t._version) for t in ts)
        hit = self._packed.get(mode)
        if hit is None or hit[0] != key:
            self._packed[mode] = (key, nsm.unet_pack(ts, mode))
        return self._packed[mode][1]

    def _workspace(self, B, H, W, mode, device):
        k = (B, H, W, mode, device)
        ws = self._ws.get(k)
        if ws is None:
            if len(self._ws) >= 4:
                self._ws.clear()
            ws = self._ws[k] = nsm.unet_workspace(B, H, W, mode, device)
        return ws

    def set_input_stats(self, means=None, stds=None):
        """Fuse MmapLiverDataset's (x - mean) / (std + 1e-8) into the first kernel (pass raw G-buffers then)."""
        if means is None:
            self.input_stats = None
        else:
            dev = self.conv10.weight.device
            self.input_stats = (torch.as_tensor(means, dtype=torch.float32, device=dev).contiguous(),
                                torch.as_tensor(stds, dtype=torch.float32, device=dev).contiguous())

    # -- forward ------------------------------------------------------------------------------------------------
    def forward(self, x):
        if x.dim() != 4 or x.shape[1] != 4:
            raise ValueError(f"expected [B,4,H,W], got {tuple(x.shape)}")
        nsm.require_device(x)
        mode = self._mode()
        if self.training:
            from nsm_train import unet_train_forward  # deferred: training kernels are a separate layer
            return unet_train_forward(self, x, mode)
        B, _, H, W = x.shape
        xin = x.detach().to(torch.float32).contiguous()
        blob = self._packed_blob(mode)
        ws = self._workspace(B, H, W, mode, x.device)
        y = torch.empty(B, 1, H - H % 2, W - W % 2, dtype=torch.float32, device=x.device)
        mean, std = self.input_stats if self.input_stats is not None else (None, None)
        nsm.unet_infer(blob, mode, xin, y, ws, mean, std)
        self.last_workspace = (ws, B, H, W, mode)
        dt = self._out_dtype(mode)
        return y if dt == torch.float32 else y.to(dt)

    @torch.no_grad()
    def infer_host(self, x_host, y_host=None):
        """End-to-end call with HOST buffers (what infer.py:46-68 does around the model): H2D copy, forward, D2H copy,
        synchronise -- one C-ABI call (nsm_unet_infer_host)."""
        dev = self.conv10.weight.device
        nsm.require_device(self.conv10.weight)
        mode = self._mode()
        B, _, H, W = x_host.shape
        if y_host is None:
            y_host = torch.empty(B, 1, H - H % 2, W - W % 2, dtype=torch.float32).pin_memory()
        blob = self._packed_blob(mode)
        ws = self._workspace(B, H, W, mode, dev)
        mean, std = self.input_stats if self.input_stats is not None else (None, None)
        nsm.unet_infer_host(blob, mode, x_host, y_host, ws, mean, std)
        self.last_workspace = (ws, B, H, W, mode)
        return y_host


def _infer_host_u8(self, x_host, y_host=None):
    """As infer_host, but the last kernel also applies infer.py:79's `(out * 255).astype(uint8)`: the result comes back as
    [B,1,H',W'] uint8 (a quarter of the D2H bytes)."""
    dev = self.conv10.weight.device
    nsm.require_device(self.conv10.weight)
    mode = self._mode()
    B, _, H, W = x_host.shape
    if y_host is None:
        y_host = torch.empty(B, 1, H - H % 2, W - W % 2, dtype=torch.uint8).pin_memory()
    blob = self._packed_blob(mode)
    ws = self._workspace(B, H, W, mode, dev)
    mean, std = self.input_stats if self.input_stats is not None else (None, None)
    with torch.no_grad():
        nsm.unet_infer_host_u8(blob, mode, x_host, y_host, ws, mean, std)
    return y_host


Unet.infer_host_u8 = _infer_host_u8


def _open_pipe(self, B, H, W):
    """Frame pipeline for sequences (the loop of infer.py / inference.py): `pipe.submit(x_pinned, y_pinned)` per frame,
    `pipe.sync()` at the end.  Host<->device copies of neighbouring frames overlap the kernels (nsm_unet_pipe_*).  The
    parameters are packed now: re-open the pipe after changing them."""
    dev = self.conv10.weight.device
    nsm.require_device(self.conv10.weight)
    mode = self._mode()
    mean, std = self.input_stats if self.input_stats is not None else (None, None)
    with torch.no_grad():
        return nsm.FramePipe(self._packed_blob(mode), mode, B, H, W, dev, mean, std)


Unet.open_pipe = _open_pipe


def makefilepath(folder_path):
    """Kept for interface parity with the reference helper (Unetmodel.py:152-154)."""
    os.makedirs(folder_path, exist_ok=True)
