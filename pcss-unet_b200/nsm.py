"""ctypes binding of libnsm_b200.so (C ABI declared in include/nsm_b200.h).

PyTorch is used only for device memory, streams and (in the trainer) torch.distributed; all arithmetic of the hot
path runs in the hand-written sm_100a kernels behind this boundary.  There is NO fallback: if the library is missing
or the device is not a B200 the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnsm_b200.so")

MODE_BF16 = 0
MODE_FP32 = 1
MODE_FP32_TRAIN = 2
# storage format of the OPERANDS of the decoder's 3x3 convolutions in fp32 mode (fp16 hi plane + 8-bit cross plane, see
# csrc/nsm_common.cuh); whole-network calls never take it, only the stage-level entry points
FMT_F16_X8 = 3
MODES = {"bf16": MODE_BF16, "fp32": MODE_FP32, "fp32_train": MODE_FP32_TRAIN}
NUM_TENSORS = 98

_lib = None
_lock = threading.Lock()


class NsmError(RuntimeError):
    pass


class ConvArgs(Structure):
    _fields_ = [
        ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int), ("ksize", c_int),
        ("mode", c_int),
        ("inp", c_void_p * 2), ("weight", c_void_p * 2),
        ("bias", c_void_p), ("bn_scale", c_void_p), ("bn_shift", c_void_p),
        ("lrelu", c_int),
        ("out", c_void_p * 2), ("residual", c_void_p * 2), ("pool", c_void_p * 2),
        ("out_f32", c_void_p),
        ("stats", c_void_p),
    ]


class UpBlockArgs(Structure):
    _fields_ = [
        ("mode", c_int), ("N", c_int), ("Hs", c_int), ("Ws", c_int), ("H", c_int), ("W", c_int), ("Cmid", c_int),
        ("Cout", c_int),
        ("src", c_void_p * 2), ("weight3", c_void_p * 2), ("weight1", c_void_p * 2),
        ("bias3", c_void_p), ("bn_scale3", c_void_p), ("bn_shift3", c_void_p),
        ("bias1", c_void_p), ("bn_scale1", c_void_p), ("bn_shift1", c_void_p),
        ("residual", c_void_p * 2), ("out", c_void_p * 2),
        ("tail", c_int),
        ("w10", c_void_p), ("b10", c_void_p),
        ("y", c_void_p), ("y_u8", c_void_p),
    ]


def _declare(lib):
    lib.nsm_last_error.restype = c_char_p
    lib.nsm_version.restype = c_int
    lib.nsm_launch_count.restype = c_longlong
    lib.nsm_tmap_cache_stats.restype = None
    lib.nsm_tmap_cache_stats.argtypes = [POINTER(c_longlong), POINTER(c_longlong)]
    lib.nsm_check_device.restype = c_int
    lib.nsm_unet_packed_bytes.restype = c_size_t
    lib.nsm_unet_packed_bytes.argtypes = [c_int]
    lib.nsm_unet_pack.argtypes = [POINTER(c_void_p), c_int, c_void_p, c_void_p]
    lib.nsm_unet_workspace_bytes.restype = c_size_t
    lib.nsm_unet_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    infer_args = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                  c_size_t, c_void_p]
    lib.nsm_unet_infer.argtypes = infer_args
    lib.nsm_unet_infer_host.argtypes = infer_args
    lib.nsm_unet_infer_u8.argtypes = infer_args
    lib.nsm_unet_infer_host_u8.argtypes = infer_args
    lib.nsm_unet_tap.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_char_p, c_void_p, POINTER(c_int),
                                 POINTER(c_int), POINTER(c_int), c_void_p]
    lib.nsm_nchw_to_planes.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.nsm_planes_to_nchw.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.nsm_pack_conv_weight.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.nsm_conv_fwd.argtypes = [POINTER(ConvArgs), c_void_p]
    lib.nsm_upblock.argtypes = [POINTER(UpBlockArgs), c_void_p]
    lib.nsm_unet_fused_decoder.restype = c_int
    lib.nsm_unet_set_fused_decoder.argtypes = [c_int]
    lib.nsm_upsample_match.argtypes = [POINTER(c_void_p), c_int, c_int, c_int, c_int, POINTER(c_void_p), c_int,
                                       c_int, c_int, c_void_p]
    lib.nsm_add_noise_clamp.argtypes = [c_void_p, c_void_p, c_longlong, c_float, c_float, c_float, c_void_p, c_void_p]
    lib.nsm_mse_loss_fwd_bwd.argtypes = [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]
    lib.nsm_l1_loss_fwd_bwd.argtypes = [c_void_p, c_void_p, POINTER(c_void_p), c_int, c_longlong, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p]
    lib.nsm_channel_sums.argtypes = [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p, c_void_p]
    lib.nsm_standardize.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p, c_void_p]
    lib.nsm_perturb.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_int, c_longlong, c_void_p,
                                c_float, c_void_p]
    lib.nsm_unet_pipe_workspace_bytes.restype = c_size_t
    lib.nsm_unet_pipe_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.nsm_unet_pipe_create.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                         POINTER(c_void_p)]
    lib.nsm_unet_pipe_submit.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.nsm_unet_pipe_sync.argtypes = [c_void_p]
    lib.nsm_unet_pipe_destroy.argtypes = [c_void_p]
    for name in ("nsm_unet_pipe_create", "nsm_unet_pipe_submit", "nsm_unet_pipe_sync", "nsm_unet_pipe_destroy"):
        getattr(lib, name).restype = c_int
    lib.nsm_profile_enable.argtypes = [c_int]
    lib.nsm_profile_read.argtypes = [c_char_p, c_size_t]
    vp, ll, fp = c_void_p, c_longlong, c_float
    lib.nsm_bn_stats.argtypes = [vp, vp, ll, c_int, c_int, vp, vp]
    lib.nsm_bn_finalize.argtypes = [vp, ll, c_int, vp, vp, fp, fp, c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.nsm_bn_act.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.nsm_bn_bwd.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, c_int, vp, vp, vp,
                               vp, vp, vp, vp]
    lib.nsm_pool_bwd_add.argtypes = [vp, vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, vp]
    lib.nsm_planes_add.argtypes = [vp, vp, vp, vp, vp, vp, ll, c_int, vp]
    lib.nsm_bilinear_bwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_int, c_int, c_int, vp]
    lib.nsm_upsample_match_bwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_int, c_int, c_int, vp]
    lib.nsm_train_input_prep.argtypes = [vp, c_int, c_int, c_int, vp, vp, c_int, vp]
    lib.nsm_train_input_grad.argtypes = [vp, vp, c_int, c_int, c_int, vp, c_int, vp]
    lib.nsm_sigmoid_shuffle_fwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp]
    lib.nsm_sigmoid_shuffle_bwd.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_pack_conv_weight_padded.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_pad_vector.argtypes = [vp, c_int, c_int, fp, c_int, vp, vp]
    lib.nsm_wgrad_workspace_bytes.restype = c_size_t
    lib.nsm_wgrad_workspace_bytes.argtypes = [c_int] * 7
    lib.nsm_wgrad.argtypes = [vp, vp, vp, vp] + [c_int] * 9 + [vp, c_size_t, vp, vp]
    lib.nsm_adamw_clip_step.argtypes = [c_int, POINTER(vp), POINTER(vp), POINTER(vp), POINTER(vp), POINTER(ll), fp, fp,
                                        fp, fp, fp, fp, c_int, vp, vp]
    lib.nsm_train_input_prep_c16.argtypes = [vp, c_int, c_int, c_int, vp, vp, c_int, vp]
    lib.nsm_train_input_grad_c16.argtypes = [vp, vp, c_int, c_int, c_int, vp, c_int, vp]
    lib.nsm_sigmoid_shuffle_fwd_px4.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp]
    lib.nsm_sigmoid_shuffle_bwd_px4.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_pack_conv_weight_px4.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_px4_reduce_dw.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, vp, vp]
    lib.nsm_fold_channel_sums.argtypes = [vp, c_int, c_int, c_int, c_int, vp, vp]
    lib.nsm_tile_vector.argtypes = [vp, c_int, c_int, c_int, fp, c_int, vp, vp]
    for name in PX4_EXPORTS:
        getattr(lib, name).restype = c_int
    lib.nsm_vgg_input_prep.argtypes = [vp, vp, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_relu_maxpool.argtypes = [vp, vp, c_int, c_int, c_int, c_int, c_int, c_int, vp, vp, vp]
    lib.nsm_feature_l1.argtypes = [vp, vp, ll, c_int, vp, vp]
    lib.nsm_acc_to_double.argtypes = [vp, ll, vp, vp]
    for name in ("nsm_vgg_input_prep", "nsm_relu_maxpool", "nsm_feature_l1", "nsm_acc_to_double"):
        getattr(lib, name).restype = c_int
    for name in TRAIN_EXPORTS:
        if name != "nsm_wgrad_workspace_bytes":
            getattr(lib, name).restype = c_int
    for name in ("nsm_unet_infer_u8", "nsm_unet_infer_host_u8", "nsm_profile_enable", "nsm_profile_read", "nsm_unet_pack", "nsm_unet_infer", "nsm_unet_infer_host", "nsm_unet_tap", "nsm_nchw_to_planes",
                 "nsm_planes_to_nchw", "nsm_pack_conv_weight", "nsm_conv_fwd", "nsm_upsample_match",
                 "nsm_l1_loss_fwd_bwd", "nsm_channel_sums", "nsm_standardize", "nsm_perturb", "nsm_add_noise_clamp",
                 "nsm_mse_loss_fwd_bwd"):
        getattr(lib, name).restype = c_int


PX4_EXPORTS = [
    "nsm_train_input_prep_c16", "nsm_train_input_grad_c16", "nsm_sigmoid_shuffle_fwd_px4", "nsm_sigmoid_shuffle_bwd_px4",
    "nsm_pack_conv_weight_px4", "nsm_px4_reduce_dw", "nsm_fold_channel_sums", "nsm_tile_vector",
]

TRAIN_EXPORTS = [
    "nsm_bn_stats", "nsm_bn_finalize", "nsm_bn_act", "nsm_bn_bwd", "nsm_pool_bwd_add", "nsm_planes_add",
    "nsm_bilinear_bwd", "nsm_upsample_match_bwd", "nsm_train_input_prep", "nsm_train_input_grad", "nsm_sigmoid_shuffle_fwd",
    "nsm_sigmoid_shuffle_bwd", "nsm_pack_conv_weight_padded", "nsm_pad_vector", "nsm_wgrad_workspace_bytes",
    "nsm_wgrad", "nsm_adamw_clip_step",
]

EXPORTS = TRAIN_EXPORTS + [
    "nsm_launch_count", "nsm_tmap_cache_stats", "nsm_add_noise_clamp", "nsm_mse_loss_fwd_bwd", "nsm_unet_infer_u8", "nsm_unet_infer_host_u8", "nsm_unet_pipe_workspace_bytes",
    "nsm_unet_pipe_create", "nsm_unet_pipe_submit", "nsm_unet_pipe_sync", "nsm_unet_pipe_destroy",
    "nsm_last_error", "nsm_version", "nsm_check_device", "nsm_unet_packed_bytes", "nsm_unet_pack",
    "nsm_unet_workspace_bytes", "nsm_unet_infer", "nsm_unet_infer_host", "nsm_unet_tap", "nsm_nchw_to_planes",
    "nsm_planes_to_nchw", "nsm_pack_conv_weight", "nsm_conv_fwd", "nsm_upsample_match", "nsm_l1_loss_fwd_bwd",
    "nsm_channel_sums", "nsm_standardize", "nsm_perturb", "nsm_profile_enable", "nsm_profile_read",
    "nsm_vgg_input_prep", "nsm_relu_maxpool", "nsm_feature_l1", "nsm_upblock", "nsm_unet_fused_decoder", "nsm_unet_set_fused_decoder", "nsm_upblock_prof",
    "nsm_acc_to_double",
] + PX4_EXPORTS


def lib():
    """Load the shared library once.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NsmError(f"{LIB_PATH} is missing: build it with `python pcss-unet_b200/build.py` "
                                   "(there is no CPU / PyTorch fallback for the hot path)")
                l = ctypes.CDLL(LIB_PATH)
                _declare(l)
                _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().nsm_last_error().decode(errors="replace")
        raise NsmError(f"{what or 'nsm call'} failed (code {rc}): {msg}")


def require_device(t: torch.Tensor = None):
    if not torch.cuda.is_available():
        raise NsmError("CUDA device required: the B200 hot path has no CPU fallback "
                       "(use the reference classes on CPU)")
    if t is not None and not t.is_cuda:
        raise NsmError("tensor must live on the CUDA device")
    check(lib().nsm_check_device(), "nsm_check_device")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def mode_planes(mode: int) -> int:
    return 1 if mode == MODE_BF16 else 2


# ---------------------------------------------------------------------------------------------------------------
# thin typed wrappers
# ---------------------------------------------------------------------------------------------------------------

def launch_count() -> int:
    return int(lib().nsm_launch_count())


def tmap_cache_stats():
    """(hits, misses) of the TMA-descriptor table since the library was loaded."""
    h, m = c_longlong(0), c_longlong(0)
    lib().nsm_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
    return int(h.value), int(m.value)


def profile_enable(on: bool):
    check(lib().nsm_profile_enable(int(on)), "nsm_profile_enable")


def profile_read():
    """[(name, ms, flops, bytes)] for every launch since profile_enable(True) / the previous read."""
    buf = ctypes.create_string_buffer(1 << 20)
    check(lib().nsm_profile_read(buf, len(buf)), "nsm_profile_read")
    rows = []
    for line in buf.value.decode().splitlines():
        name, ms, fl, by = line.rsplit(",", 3)
        rows.append((name, float(ms), float(fl), float(by)))
    return rows


def unet_pack(tensors, mode: int) -> torch.Tensor:
    """tensors: the 98 fp32 CUDA tensors in the order of include/nsm_b200.h -> packed uint8 blob."""
    assert len(tensors) == NUM_TENSORS
    dev = tensors[0].device
    keep = [t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in tensors]
    arr = (c_void_p * NUM_TENSORS)(*[t.data_ptr() for t in keep])
    blob = torch.empty(lib().nsm_unet_packed_bytes(mode), dtype=torch.uint8, device=dev)
    check(lib().nsm_unet_pack(arr, mode, blob.data_ptr(), stream_ptr()), "nsm_unet_pack")
    return blob


def unet_workspace(B, H, W, mode, device) -> torch.Tensor:
    n = lib().nsm_unet_workspace_bytes(B, H, W, mode)
    if n == 0:
        raise NsmError(f"input {B}x4x{H}x{W} not supported (need H, W >= 16)")
    return torch.empty(n, dtype=torch.uint8, device=device)


def unet_infer(blob, mode, x, y, ws, mean=None, std=None):
    B, C, H, W = x.shape
    assert C == 4 and x.dtype == torch.float32 and x.is_contiguous()
    check(lib().nsm_unet_infer(blob.data_ptr(), mode, x.data_ptr(), B, H, W, ptr(mean), ptr(std), y.data_ptr(),
                               ws.data_ptr(), ws.numel(), stream_ptr()), "nsm_unet_infer")


def unet_infer_host(blob, mode, x_host, y_host, ws, mean=None, std=None):
    B, C, H, W = x_host.shape
    assert C == 4 and x_host.dtype == torch.float32 and x_host.is_contiguous() and not x_host.is_cuda
    check(lib().nsm_unet_infer_host(blob.data_ptr(), mode, x_host.data_ptr(), B, H, W, ptr(mean), ptr(std),
                                    y_host.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "nsm_unet_infer_host")


def unet_infer_host_u8(blob, mode, x_host, y_host, ws, mean=None, std=None):
    B, C, H, W = x_host.shape
    assert C == 4 and x_host.dtype == torch.float32 and x_host.is_contiguous() and not x_host.is_cuda
    assert y_host.dtype == torch.uint8 and y_host.is_contiguous() and not y_host.is_cuda
    check(lib().nsm_unet_infer_host_u8(blob.data_ptr(), mode, x_host.data_ptr(), B, H, W, ptr(mean), ptr(std),
                                       y_host.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "nsm_unet_infer_host_u8")


class FramePipe:
    """Frame pipeline (nsm_unet_pipe_*): copies of neighbouring frames overlap the kernels of the current one.
    submit(x_host, y_host) takes pinned [B,4,H,W] fp32 and a pinned [B,1,H',W'] fp32 or uint8 result buffer; when the
    call for frame k returns, result k-2 is complete; sync() completes everything submitted."""

    def __init__(self, blob, mode, B, H, W, device, mean=None, std=None):
        n = lib().nsm_unet_pipe_workspace_bytes(B, H, W, mode)
        if n == 0:
            raise NsmError(f"input {B}x4x{H}x{W} not supported (need H, W >= 16)")
        self.shape = (B, 4, H, W)
        self.out_shape = (B, 1, H - H % 2, W - W % 2)
        self._keep = (blob, mean, std, torch.empty(n, dtype=torch.uint8, device=device))
        torch.cuda.current_stream(device).synchronize()   # the blob / statistics were produced on torch's stream
        h = c_void_p()
        check(lib().nsm_unet_pipe_create(blob.data_ptr(), mode, B, H, W, ptr(mean), ptr(std), self._keep[3].data_ptr(),
                                         n, byref(h)), "nsm_unet_pipe_create")
        self._h = h

    def submit(self, x_host, y_host):
        assert tuple(x_host.shape) == self.shape and x_host.dtype == torch.float32 and x_host.is_contiguous()
        assert tuple(y_host.shape) == self.out_shape and y_host.is_contiguous()
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise NsmError("FramePipe.submit needs pinned host tensors (the copies are asynchronous)")
        if y_host.dtype == torch.float32:
            rc = lib().nsm_unet_pipe_submit(self._h, x_host.data_ptr(), y_host.data_ptr(), None)
        elif y_host.dtype == torch.uint8:
            rc = lib().nsm_unet_pipe_submit(self._h, x_host.data_ptr(), None, y_host.data_ptr())
        else:
            raise NsmError("FramePipe.submit: result buffer must be float32 or uint8")
        check(rc, "nsm_unet_pipe_submit")

    def sync(self):
        check(lib().nsm_unet_pipe_sync(self._h), "nsm_unet_pipe_sync")

    def close(self):
        if self._h is not None:
            lib().nsm_unet_pipe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def unet_tap(ws, B, H, W, mode, name) -> torch.Tensor:
    C, h, w = c_int(), c_int(), c_int()
    check(lib().nsm_unet_tap(ws.data_ptr(), B, H, W, mode, name.encode(), None, byref(C), byref(h), byref(w),
                             stream_ptr()), "nsm_unet_tap")
    out = torch.empty(B, C.value, h.value, w.value, dtype=torch.float32, device=ws.device)
    check(lib().nsm_unet_tap(ws.data_ptr(), B, H, W, mode, name.encode(), out.data_ptr(), None, None, None,
                             stream_ptr()), "nsm_unet_tap")
    return out


class PlaneTensor:
    """NHWC bf16 planes of one activation (plane 1 only in fp32 mode)."""

    def __init__(self, N, C, H, W, mode, device):
        self.shape = (N, C, H, W)
        self.mode = mode
        self.p0 = torch.empty(N, H, W, C, dtype=torch.bfloat16, device=device)
        self.p1 = torch.empty(N, H, W, C, dtype=torch.bfloat16, device=device) if mode != MODE_BF16 else None

    @classmethod
    def from_nchw(cls, x, mode):
        N, C, H, W = x.shape
        x = x.to(torch.float32).contiguous()
        t = cls(N, C, H, W, mode, x.device)
        check(lib().nsm_nchw_to_planes(x.data_ptr(), N, C, H, W, mode, ptr(t.p0), ptr(t.p1), stream_ptr()),
              "nsm_nchw_to_planes")
        return t

    def to_nchw(self):
        N, C, H, W = self.shape
        y = torch.empty(N, C, H, W, dtype=torch.float32, device=self.p0.device)
        check(lib().nsm_planes_to_nchw(ptr(self.p0), ptr(self.p1), N, C, H, W, self.mode, y.data_ptr(),
                                       stream_ptr()), "nsm_planes_to_nchw")
        return y

    def pair(self):
        return (c_void_p * 2)(ptr(self.p0), ptr(self.p1))

    def px_view(self, px):
        """The same bytes addressed as [N, H, W/px, px*C]: px horizontally adjacent pixels = one pixel of px*C virtual
        channels (px > 0), or the inverse view (px < 0: [N, H, W*|px|, C/|px|])."""
        N, C, H, W = self.shape
        v = object.__new__(PlaneTensor)
        if px > 0:
            assert W % px == 0
            v.shape = (N, C * px, H, W // px)
        else:
            assert C % (-px) == 0
            v.shape = (N, C // (-px), H, W * (-px))
        v.mode = self.mode
        n, c, h, w = v.shape
        v.p0 = self.p0.view(n, h, w, c)
        v.p1 = None if self.p1 is None else self.p1.view(n, h, w, c)
        return v


def pack_conv_weight(w, mode, dgrad=False):
    """nn.Conv2d weight [Cout,Cin,k,k] fp32 -> (plane0, plane1|None) bf16 [rows][k*k][inner]."""
    Cout, Cin, k, _ = w.shape
    w = w.detach().to(torch.float32).contiguous()
    rows, inner = (Cin, Cout) if dgrad else (Cout, Cin)
    p0 = torch.empty(rows, k * k, inner, dtype=torch.bfloat16, device=w.device)
    p1 = torch.empty_like(p0) if mode != MODE_BF16 else None
    check(lib().nsm_pack_conv_weight(w.data_ptr(), Cout, Cin, k, int(dgrad), mode, ptr(p0), ptr(p1), stream_ptr()),
          "nsm_pack_conv_weight")
    return p0, p1


def conv_fwd(x: PlaneTensor, wp, ksize, Cout, mode, bias=None, bn_scale=None, bn_shift=None, lrelu=False,
             residual: PlaneTensor = None, pool=False, want_f32=False, want_out=True, stats=None):
    """One fused conv stage on the tensor cores.  Returns (out PlaneTensor|None, pooled|None, raw fp32 NHWC|None)."""
    N, Cin, H, W = x.shape
    dev = x.p0.device
    omode = MODE_FP32 if mode == FMT_F16_X8 else mode      # 8-bit cross operands, plain fp16 hi+lo results
    out = PlaneTensor(N, Cout, H, W, omode, dev) if want_out else None
    pl = PlaneTensor(N, Cout, H // 2, W // 2, omode, dev) if pool else None
    raw = torch.empty(N, H, W, Cout, dtype=torch.float32, device=dev) if want_f32 else None
    a = ConvArgs()
    a.N, a.H, a.W, a.Cin, a.Cout, a.ksize, a.mode = N, H, W, Cin, Cout, ksize, mode
    a.inp = x.pair()
    a.weight = (c_void_p * 2)(ptr(wp[0]), ptr(wp[1]))
    a.bias, a.bn_scale, a.bn_shift = ptr(bias) or None, ptr(bn_scale) or None, ptr(bn_shift) or None
    a.lrelu = int(lrelu)
    a.out = out.pair() if out is not None else (c_void_p * 2)(None, None)
    a.residual = residual.pair() if residual is not None else (c_void_p * 2)(None, None)
    a.pool = pl.pair() if pl is not None else (c_void_p * 2)(None, None)
    a.out_f32 = ptr(raw) or None
    a.stats = ptr(stats) or None
    check(lib().nsm_conv_fwd(byref(a), stream_ptr()), "nsm_conv_fwd")
    return out, pl, raw


def upblock(src: PlaneTensor, H, W, w3p, w1p, Cout, v3, v1, residual: PlaneTensor = None, tail=None, want_u8=False):
    """Fused decoder block (nsm_upblock): composite up-sample of `src` to (H, W) on the operand path -> 3x3 + BN + LReLU
    -> 1x1 + BN + LReLU -> + residual, or (tail = (w10 [4,16], b10 [4])) conv10 + sigmoid + pixel_shuffle.
    v3 / v1 = (bias, bn_scale, bn_shift) of the two stages.  Returns the output PlaneTensor, or (y, y_u8 | None)."""
    N, Cmid, Hs, Ws = src.shape
    dev = src.p0.device
    a = UpBlockArgs()
    a.mode, a.N, a.Hs, a.Ws, a.H, a.W, a.Cmid, a.Cout = src.mode, N, Hs, Ws, H, W, Cmid, Cout
    a.src = src.pair()
    a.weight3 = (c_void_p * 2)(ptr(w3p[0]), ptr(w3p[1]))
    a.weight1 = (c_void_p * 2)(ptr(w1p[0]), ptr(w1p[1]))
    a.bias3, a.bn_scale3, a.bn_shift3 = (ptr(t) for t in v3)
    a.bias1, a.bn_scale1, a.bn_shift1 = (ptr(t) for t in v1)
    a.residual = residual.pair() if residual is not None else (c_void_p * 2)(None, None)
    out = y = y8 = None
    if tail is None:
        out = PlaneTensor(N, Cout, H, W, src.mode, dev)
        a.out, a.tail = out.pair(), 0
    else:
        a.out, a.tail = (c_void_p * 2)(None, None), 1
        a.w10, a.b10 = ptr(tail[0]), ptr(tail[1])
        y = torch.empty(N, 1, 2 * H, 2 * W, dtype=torch.float32, device=dev)
        a.y = y.data_ptr()
        if want_u8:
            y8 = torch.empty(N, 1, 2 * H, 2 * W, dtype=torch.uint8, device=dev)
            a.y_u8 = y8.data_ptr()
    check(lib().nsm_upblock(byref(a), stream_ptr()), "nsm_upblock")
    return out if tail is None else (y, y8)


def upsample_match(x: PlaneTensor, hd, wd, out_x8=False):
    """out_x8: fp32-mode source, result in the 8-bit cross operand format of the decoder's 3x3 convolutions."""
    N, C, hs, ws = x.shape
    omode = FMT_F16_X8 if out_x8 else x.mode
    assert not out_x8 or x.mode == MODE_FP32
    out = PlaneTensor(N, C, hd, wd, omode, x.p0.device)
    check(lib().nsm_upsample_match(x.pair(), N, hs, ws, C, out.pair(), hd, wd, omode, stream_ptr()),
          "nsm_upsample_match")
    return out


def vgg_input_prep(output, target, mode):
    """[B,1,H,W] fp32 output / target -> PlaneTensor [2B,64,H,W] (3 real channels), customLoss.py:44-61."""
    B, _, H, W = output.shape
    x = PlaneTensor(2 * B, 64, H, W, mode, output.device)
    check(lib().nsm_vgg_input_prep(output.data_ptr(), target.data_ptr(), B, H, W, mode, *_pp(x), stream_ptr()),
          "nsm_vgg_input_prep")
    return x


def relu_maxpool(x: PlaneTensor, pool: bool):
    N, C, H, W = x.shape
    out = PlaneTensor(N, C, H // 2 if pool else H, W // 2 if pool else W, x.mode, x.p0.device)
    check(lib().nsm_relu_maxpool(*_pp(x), N, H, W, C, int(pool), x.mode, *_pp(out), stream_ptr()), "nsm_relu_maxpool")
    return out


def feature_l1(f: PlaneTensor, acc):
    """acc (ONE accumulator slot: a [1, 4] row view of nsm.acc_zeros) += sum |a - b| over the two halves of the batch."""
    N, C, H, W = f.shape
    check(lib().nsm_feature_l1(*_pp(f), (N // 2) * C * H * W, f.mode, acc.data_ptr(), stream_ptr()), "nsm_feature_l1")


def acc_zeros(n, device):
    """n zeroed order-independent accumulator slots (include/nsm_b200.h: nsm_acc = four 64-bit words each)."""
    return torch.zeros(n, 4, dtype=torch.int64, device=device)


def acc_to_double(acc):
    """fp64 values of accumulator slots [n, 4] -> float64 [n] (device)."""
    n = acc.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=acc.device)
    check(lib().nsm_acc_to_double(acc.data_ptr(), n, out.data_ptr(), stream_ptr()), "nsm_acc_to_double")
    return out


def l1_loss_fwd_bwd(out, target=None, perturbed=(), coef_l1=0.0, coef_pert=0.0, want_grad=True):
    """Returns (acc[3] float64 device tensor: sum|o-t|, sum_i sum|o-y_i|, #out-of-range; grad or None)."""
    def aligned(t):
        # (the kernel takes its 16-byte vector path only when every pointer is 16-byte aligned, so a contiguous view at an
        # odd element offset needs no copy)
        return t.contiguous()
    out = aligned(out)
    assert out.dtype == torch.float32
    n = out.numel()
    acc = acc_zeros(3, out.device)
    grad = torch.empty_like(out) if want_grad else None
    pert = [aligned(p) for p in perturbed]
    arr = (c_void_p * max(1, len(pert)))(*[p.data_ptr() for p in pert])
    tgt = None if target is None else aligned(target.to(torch.float32))
    check(lib().nsm_l1_loss_fwd_bwd(out.data_ptr(), ptr(tgt), arr, len(pert), n, coef_l1, coef_pert, ptr(grad),
                                    acc.data_ptr(), stream_ptr()), "nsm_l1_loss_fwd_bwd")
    return acc_to_double(acc), grad


def add_noise_clamp(x, noise, eps, lo, hi):
    """clamp(x + eps * noise, lo, hi) in one pass (customLoss.py:226-231)."""
    x = x.contiguous()
    noise = noise.contiguous()
    assert x.dtype == torch.float32 and noise.dtype == torch.float32 and x.shape == noise.shape
    out = torch.empty_like(x)
    check(lib().nsm_add_noise_clamp(x.data_ptr(), noise.data_ptr(), x.numel(), eps, lo, hi, out.data_ptr(), stream_ptr()),
          "nsm_add_noise_clamp")
    return out


def mse_loss_fwd_bwd(out, ref, want_diff=True):
    """Returns (float64 device scalar sum (out-ref)^2, out - ref or None)."""
    out = out.contiguous()
    ref = ref.to(torch.float32).contiguous()
    assert out.dtype == torch.float32 and out.shape == ref.shape
    acc = acc_zeros(1, out.device)
    diff = torch.empty_like(out) if want_diff else None
    check(lib().nsm_mse_loss_fwd_bwd(out.data_ptr(), ref.data_ptr(), out.numel(), ptr(diff), acc.data_ptr(), stream_ptr()),
          "nsm_mse_loss_fwd_bwd")
    return acc_to_double(acc)[0], diff


def channel_sums(x, means=None):
    """x: [S,C,...] fp32 CUDA.  Returns float64 [C]: sum x (means None) or sum (x-mean_c)^2."""
    x = x.contiguous()
    S, C = x.shape[0], x.shape[1]
    HW = x.numel() // (S * C)
    sums = acc_zeros(C, x.device)
    m = None if means is None else means.to(device=x.device, dtype=torch.float64).contiguous()
    check(lib().nsm_channel_sums(x.data_ptr(), S, C, HW, ptr(m), sums.data_ptr(), stream_ptr()),
          "nsm_channel_sums")
    return acc_to_double(sums)


def standardize(x, mean, std):
    x = x.contiguous()
    C = mean.numel()
    if x.dim() == 3:
        S, HW = 1, x.shape[1] * x.shape[2]
    else:
        S, HW = x.shape[0], x.numel() // (x.shape[0] * C)
    y = torch.empty_like(x)
    check(lib().nsm_standardize(x.data_ptr(), y.data_ptr(), S, C, HW, mean.data_ptr(), std.data_ptr(),
                                stream_ptr()), "nsm_standardize")
    return y


def perturb(x, noise, stds, std_factor):
    """x [B,C,H,W], noise [count,C,B,H,W] (one block per reference randn_like draw), stds [C] -> [count,B,C,H,W]."""
    x = x.contiguous()
    noise = noise.contiguous()
    count = noise.shape[0]
    B, C = x.shape[0], x.shape[1]
    HW = x.numel() // (B * C)
    out = torch.empty((count,) + tuple(x.shape), dtype=torch.float32, device=x.device)
    check(lib().nsm_perturb(x.data_ptr(), noise.data_ptr(), out.data_ptr(), count, B, C, HW, stds.data_ptr(),
                            float(std_factor), stream_ptr()), "nsm_perturb")
    return out


# ---------------------------------------------------------------------------------------------------------------
# training-stage wrappers (all arithmetic in libnsm_b200.so; torch only allocates)
# ---------------------------------------------------------------------------------------------------------------

def _pp(t):
    """(plane0 ptr, plane1 ptr) of a PlaneTensor or (0, 0) for None."""
    return (0, 0) if t is None else (ptr(t.p0), ptr(t.p1))


def bn_stats(z: PlaneTensor):
    N, C, H, W = z.shape
    sums = acc_zeros(2 * C, z.p0.device)
    check(lib().nsm_bn_stats(*_pp(z), N * H * W, C, z.mode, sums.data_ptr(), stream_ptr()), "nsm_bn_stats")
    return sums


def bn_finalize(sums, P, gamma, beta, running_mean, running_var, updates=1, eps=1e-5, momentum=0.1):
    C = gamma.numel()
    dev = gamma.device
    out = torch.empty(4, C, dtype=torch.float32, device=dev)   # scale, shift, mean, invstd
    check(lib().nsm_bn_finalize(sums.data_ptr(), P, C, gamma.data_ptr(), beta.data_ptr(), eps, momentum, updates,
                                ptr(running_mean), ptr(running_var), out[0].data_ptr(), out[1].data_ptr(),
                                out[2].data_ptr(), out[3].data_ptr(), stream_ptr()), "nsm_bn_finalize")
    return out


def bn_act(z: PlaneTensor, scale, shift, mask=None, lrelu=True, residual: PlaneTensor = None, pool=False):
    N, C, H, W = z.shape
    dev = z.p0.device
    out = PlaneTensor(N, C, H, W, z.mode, dev)
    pl = PlaneTensor(N, C, H // 2, W // 2, z.mode, dev) if pool else None
    check(lib().nsm_bn_act(*_pp(z), N, H, W, C, z.mode, scale.data_ptr(), shift.data_ptr(), ptr(mask), int(lrelu),
                           *_pp(residual), *_pp(out), *_pp(pl), stream_ptr()), "nsm_bn_act")
    return out, pl


def bn_bwd(dy: PlaneTensor, z: PlaneTensor, scale, shift, mean, invstd, mask=None, lrelu=True, want_dbias=True):
    """Returns (dz planes, dgamma[C], dbeta[C], dbias[C]|None)."""
    N, C, H, W = z.shape
    dev = z.p0.device
    dz = PlaneTensor(N, C, H, W, z.mode, dev)
    sums = acc_zeros(3 * C, dev)
    g = torch.empty(3, C, dtype=torch.float32, device=dev)
    check(lib().nsm_bn_bwd(*_pp(dy), *_pp(z), N, H, W, C, z.mode, scale.data_ptr(), shift.data_ptr(), ptr(mask),
                           mean.data_ptr(), invstd.data_ptr(), int(lrelu), sums.data_ptr(), *_pp(dz),
                           g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr() if want_dbias else 0, stream_ptr()),
          "nsm_bn_bwd")
    return dz, g[0], g[1], (g[2] if want_dbias else None)


def pool_bwd_add(a: PlaneTensor, dpool: PlaneTensor, shape):
    N, C, H, W = shape
    out = PlaneTensor(N, C, H, W, dpool.mode, dpool.p0.device)
    check(lib().nsm_pool_bwd_add(*_pp(a), *_pp(dpool), *_pp(out), N, H, W, C, dpool.mode, stream_ptr()),
          "nsm_pool_bwd_add")
    return out


def planes_add(a: PlaneTensor, b: PlaneTensor):
    out = PlaneTensor(*a.shape, a.mode, a.p0.device)
    check(lib().nsm_planes_add(*_pp(a), *_pp(b), *_pp(out), a.p0.numel(), a.mode, stream_ptr()), "nsm_planes_add")
    return out


def bilinear_bwd(dout: PlaneTensor, hi, wi):
    N, C, ho, wo = dout.shape
    din = PlaneTensor(N, C, hi, wi, dout.mode, dout.p0.device)
    check(lib().nsm_bilinear_bwd(*_pp(dout), N, ho, wo, C, *_pp(din), hi, wi, dout.mode, stream_ptr()),
          "nsm_bilinear_bwd")
    return din


def upsample_match_bwd(dout: PlaneTensor, hs, ws):
    """Adjoint of upsample_match: (hd, wd) -> [(2hs, 2ws) ->] (hs, ws)."""
    N, C, hd, wd = dout.shape
    din = PlaneTensor(N, C, hs, ws, dout.mode, dout.p0.device)
    check(lib().nsm_upsample_match_bwd(*_pp(dout), N, hd, wd, C, *_pp(din), hs, ws, dout.mode, stream_ptr()),
          "nsm_upsample_match_bwd")
    return din


def pack_conv_weight_px4(w, mode, CoutV, CinV, dgrad=False):
    """Pixel-packed (4 pixels = one virtual pixel) operand of a thin convolution, see nsm_b200.h."""
    Cout, Cin, k, _ = w.shape
    w = w.detach().to(torch.float32).contiguous()
    rows, inner = (CinV, CoutV) if dgrad else (CoutV, CinV)
    p0 = torch.empty(rows, k * k, inner, dtype=torch.bfloat16, device=w.device)
    p1 = torch.empty_like(p0) if mode != MODE_BF16 else None
    check(lib().nsm_pack_conv_weight_px4(w.data_ptr(), Cout, Cin, k, CoutV, CinV, int(dgrad), mode, ptr(p0), ptr(p1),
                                         stream_ptr()), "nsm_pack_conv_weight_px4")
    return p0, p1


def px4_reduce_dw(dwv, Cout, Cin, ksize):
    CoutV, CinV = dwv.shape[0], dwv.shape[1]
    dw = torch.empty(Cout, Cin, ksize, ksize, dtype=torch.float32, device=dwv.device)
    check(lib().nsm_px4_reduce_dw(dwv.data_ptr(), Cout, Cin, ksize, CoutV, CinV, dw.data_ptr(), stream_ptr()),
          "nsm_px4_reduce_dw")
    return dw


def fold_channel_sums(sums, nvec, groups, C):
    """[nvec * CV] per-virtual-channel accumulator slots -> [nvec * C]: out[v][c] = sum_g in[v][g*C + c]."""
    CV = sums.shape[0] // nvec
    out = torch.empty(nvec * C, 4, dtype=torch.int64, device=sums.device)
    check(lib().nsm_fold_channel_sums(sums.data_ptr(), nvec, CV, groups, C, out.data_ptr(), stream_ptr()),
          "nsm_fold_channel_sums")
    return out


def tile_vector(v, rep, npad, fill=0.0, round_bf16=False):
    v = v.detach().to(torch.float32).contiguous()
    out = torch.empty(npad, dtype=torch.float32, device=v.device)
    check(lib().nsm_tile_vector(v.data_ptr(), v.numel(), rep, npad, fill, int(round_bf16), out.data_ptr(), stream_ptr()),
          "nsm_tile_vector")
    return out


def train_input_prep(x, mode, c16=False):
    """x [N,4,Hin,Win] -> un-shuffled NHWC planes: 64 zero-padded channels, or (c16) the 16 real ones."""
    N, _, Hin, Win = x.shape
    out = PlaneTensor(N, 16 if c16 else 64, (Hin - Hin % 2) // 2, (Win - Win % 2) // 2, mode, x.device)
    fn = lib().nsm_train_input_prep_c16 if c16 else lib().nsm_train_input_prep
    check(fn(x.data_ptr(), N, Hin, Win, *_pp(out), mode, stream_ptr()), "nsm_train_input_prep")
    return out


def train_input_grad(d: PlaneTensor, H, W):
    N = d.shape[0]
    dx = torch.empty(N, 4, H, W, dtype=torch.float32, device=d.p0.device)
    fn = lib().nsm_train_input_grad_c16 if d.shape[1] == 16 else lib().nsm_train_input_grad
    check(fn(*_pp(d), N, H, W, dx.data_ptr(), d.mode, stream_ptr()), "nsm_train_input_grad")
    return dx


def sigmoid_shuffle_fwd(c10: PlaneTensor, px4=False):
    """c10: [N,64,h,w] (4 real channels), or px4: the pixel-packed [N,64,h,w/4] (see nsm_b200.h)."""
    N, _, h, w = c10.shape
    if px4:
        w *= 4
    y = torch.empty(N, 1, 2 * h, 2 * w, dtype=torch.float32, device=c10.p0.device)
    fn = lib().nsm_sigmoid_shuffle_fwd_px4 if px4 else lib().nsm_sigmoid_shuffle_fwd
    check(fn(*_pp(c10), N, h, w, c10.mode, y.data_ptr(), stream_ptr()), "nsm_sigmoid_shuffle_fwd")
    return y


def sigmoid_shuffle_bwd(dy, y, mode, px4=False):
    N, _, H, W = y.shape
    d = PlaneTensor(N, 64, H // 2, W // 8 if px4 else W // 2, mode, y.device)
    dy = dy.to(torch.float32).contiguous()
    fn = lib().nsm_sigmoid_shuffle_bwd_px4 if px4 else lib().nsm_sigmoid_shuffle_bwd
    check(fn(dy.data_ptr(), y.data_ptr(), N, H // 2, W // 2, mode, *_pp(d), stream_ptr()), "nsm_sigmoid_shuffle_bwd")
    return d


def pack_conv_weight_padded(w, mode, CoutP, CinP, dgrad=False):
    Cout, Cin, k, _ = w.shape
    w = w.detach().to(torch.float32).contiguous()
    rows, inner = (CinP, CoutP) if dgrad else (CoutP, CinP)
    p0 = torch.empty(rows, k * k, inner, dtype=torch.bfloat16, device=w.device)
    p1 = torch.empty_like(p0) if mode != MODE_BF16 else None
    check(lib().nsm_pack_conv_weight_padded(w.data_ptr(), Cout, Cin, k, CoutP, CinP, int(dgrad), mode, ptr(p0), ptr(p1),
                                            stream_ptr()), "nsm_pack_conv_weight_padded")
    return p0, p1


def pad_vector(v, npad, fill=0.0, round_bf16=False):
    v = v.detach().to(torch.float32).contiguous()
    out = torch.empty(npad, dtype=torch.float32, device=v.device)
    check(lib().nsm_pad_vector(v.data_ptr(), v.numel(), npad, fill, int(round_bf16), out.data_ptr(), stream_ptr()),
          "nsm_pad_vector")
    return out


_wgrad_ws = {}


def wgrad(dz: PlaneTensor, x: PlaneTensor, ksize, Cout_real, Cin_real):
    """dW [Cout_real, Cin_real, k, k] fp32 = sum over pixels of dz (x) x (tap-shifted)."""
    N, Cout, H, W = dz.shape
    Cin = x.shape[1]
    dev = dz.p0.device
    need = lib().nsm_wgrad_workspace_bytes(N, H, W, Cout, Cin, ksize, dz.mode)
    ws = _wgrad_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = _wgrad_ws[dev] = torch.empty(need, dtype=torch.uint8, device=dev)
    dw = torch.empty(Cout_real, Cin_real, ksize, ksize, dtype=torch.float32, device=dev)
    check(lib().nsm_wgrad(*_pp(dz), *_pp(x), N, H, W, Cout, Cin, ksize, dz.mode, Cout_real, Cin_real, ws.data_ptr(),
                          ws.numel(), dw.data_ptr(), stream_ptr()), "nsm_wgrad")
    return dw
