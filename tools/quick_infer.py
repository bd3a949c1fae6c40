#!/usr/bin/env python
"""Per-layer timing of the cfg1 inference step without the bench's result checks (kernel experiments).
    python tools/quick_infer.py [fp32|bf16] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pcss-unet_b200"), ROOT]
import torch, nsm, bench
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
net = bench.make_model(prec, "cuda")
x = torch.randn(1, 4, 1080, 1920, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = net(x)
    ms = 0.0
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); y = net(x); e.record(); torch.cuda.synchronize()
        ms += s.elapsed_time(e)
    nsm.profile_enable(True)
    for _ in range(steps):
        flush.zero_(); net(x)
    torch.cuda.synchronize()
    rows = nsm.profile_read()
    nsm.profile_enable(False)
print(f"step {ms / steps:.3f} ms; finite={bool(torch.isfinite(y.float()).all())}")
agg = {}
for name, t, fl, by in rows:
    agg.setdefault(name, []).append(t)
for k, v in agg.items():
    print(f"{k:28s} {sum(v) / len(v):.4f} ms")
