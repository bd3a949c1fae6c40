"""Times nsm.bn_bwd / nsm.bn_act on the cfg2 training shapes (batch 32 x 512^2 crops), per-launch events.
    python tools/bn_probe.py [bf16|fp32_train] [reps]"""
import sys
sys.path.insert(0, 'pcss-unet_b200'); sys.path.insert(0, '.')
import torch, nsm
mode = nsm.MODES[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
shapes = [(32, 128, 256, 256), (32, 64, 256, 256), (32, 512, 128, 128), (32, 1024, 64, 64), (32, 512, 32, 32), (32, 64, 128, 128)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for N, C, H, W in shapes:
    z = nsm.PlaneTensor(N, C, H, W, mode, "cuda"); dy = nsm.PlaneTensor(N, C, H, W, mode, "cuda")
    for t in (z, dy):
        t.p0.view(torch.int16).random_(-20000, 20000)
        if t.p1 is not None: t.p1.view(torch.int16).random_(-2000, 2000)
    gamma = torch.rand(C, device="cuda") + 0.5; beta = torch.randn(C, device="cuda") * 0.1
    rm = torch.zeros(C, device="cuda"); rv = torch.ones(C, device="cuda")
    st = nsm.bn_finalize(nsm.bn_stats(z), N * H * W, gamma, beta, rm, rv)
    mask = (torch.rand(N, C, device="cuda") > 0.2).float() / 0.8
    tb, ta = [], []
    for r in range(reps):
        flush.zero_()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); nsm.bn_bwd(dy, z, st[0], st[1], st[2], st[3], mask=mask); e[1].record()
        nsm.bn_act(z, st[0], st[1], mask=mask); e[2].record()
        torch.cuda.synchronize()
        tb.append(e[0].elapsed_time(e[1])); ta.append(e[1].elapsed_time(e[2]))
    by = N * C * H * W * 2 * (2 if mode != 0 else 1)
    tb, ta = min(tb), min(ta)
    print(f"{N}x{C}x{H}x{W}: bn_bwd {tb:.3f} ms {5 * by / tb / 1e6:7.0f} GB/s   bn_act {ta:.3f} ms {2 * by / ta / 1e6:7.0f} GB/s")
