#!/usr/bin/env python
"""Worker-warp cycle breakdown of the fused decoder blocks (NSM_UB_DBG=64):  NSM_UB_DBG=64 python tools/ub_prof.py fp32"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "pcss-unet_b200"), ROOT]
import torch, nsm, bench
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
net = bench.make_model(prec, "cuda")
x = torch.randn(1, 4, 1080, 1920, device="cuda")
names = ["job wait", "job math", "job publish", "mid wait", "mid math", "final prefetch", "final wait", "final math"]
buf = (ctypes.c_ulonglong * 32)()
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    nsm.lib().nsm_upblock_prof(buf)
    net(x)
    torch.cuda.synchronize()
    nsm.lib().nsm_upblock_prof(buf)
mnames = ["wait acc1 free", "wait halo", "wait weights", "issue 3x3 tap", "wait 1x1 operand", "wait acc2 free", "issue 1x1",
          "other"]
for blk, off in (("conv8 block", 0), ("conv9 block", 16)):
    for who, nm, o in (("worker warp 0", names, off), ("MMA thread", mnames, off + 8)):
        vals = list(buf[o:o + 8])
        tot = sum(vals)
        print(f"{blk}: cycles of CTA 0 / {who} over one launch: {tot}")
        for n, v in zip(nm, vals):
            print(f"  {n:16s} {v:9d}  {100 * v / max(tot, 1):5.1f} %")
