#!/usr/bin/env python
"""Times single conv stages (training path shapes) in isolation: python tools/thin_conv_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pcss-unet_b200"), ROOT]
import torch, nsm
nsm.require_device()
mode = nsm.MODE_BF16
CASES = [("k1 64->256 @256x64 (conv2.1x1 px4)", 32, 64, 256, 64, 256, 1), ("k1 128->64 @256x256 (conv8.1x1)", 32, 128, 256, 256, 64, 1),
         ("k1 128->512 @64x64 (conv4.1x1)", 32, 128, 64, 64, 512, 1), ("k3 64->64 @256x256 (conv9.3x3)", 32, 64, 256, 256, 64, 3),
         ("k3 128->128 @256x256 (conv8.3x3)", 32, 128, 256, 256, 128, 3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
only = os.environ.get("PROBE_CASE")
for name, N, Cin, H, W, Cout, k in CASES:
    if only and only not in name:
        continue
    x = nsm.PlaneTensor(N, Cin, H, W, mode, "cuda"); x.p0.normal_()
    w = nsm.pack_conv_weight(torch.randn(Cout, Cin, k, k, device="cuda") * 0.05, mode)
    b = torch.zeros(Cout, device="cuda")
    for stats in (False, True):
        ms = []
        R = 8    # launches queued back to back: the host side (tensor-map encoding, ctypes) hides behind the kernels
        sums = nsm.acc_zeros(2 * Cout, "cuda") if stats else None
        for it in range(4):
            flush.zero_()
            torch.cuda._sleep(400000)      # let the host run ahead
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(R):
                nsm.conv_fwd(x, w, k, Cout, mode, bias=b, stats=sums)
            e.record(); torch.cuda.synchronize()
            ms.append(s.elapsed_time(e) / R)
        t = sorted(ms[1:])[len(ms[1:]) // 2]
        by = N * H * W * (Cin + Cout) * 2
        print(f"{name:40s} stats={int(stats)} {t:.3f} ms  {by / t / 1e6:7.0f} GB/s  out {N*H*W*Cout*2/t/1e6:7.0f} GB/s  {2.0*N*H*W*Cin*Cout*k*k/t/1e9:7.0f} TF/s")
