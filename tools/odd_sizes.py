import sys, os
sys.path[:0] = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pcss-unet_b200"), os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import torch, nsm, oracle
from Unetmodel import Unet
torch.manual_seed(0)
P = oracle.init_params(42)
g = torch.Generator().manual_seed(1)
oracle.calibrate_bn(P, torch.randn(1, 4, 64, 64, generator=g), generator=g)
worst = 0
for shape in [(1,4,16,16),(1,4,17,33),(2,4,18,34),(1,4,24,100),(3,4,50,70),(1,4,130,66),(1,4,257,259)]:
    x = torch.randn(*shape, generator=g)
    ref = oracle.unet_forward(x, P, training=False)
    for prec, tol in (("fp32", 1e-4), ("bf16", 6e-2)):
        net = Unet(precision=prec); net.load_state_dict(P); net = net.cuda().eval()
        with torch.no_grad():
            y = net(x.cuda()).float().cpu()
        err = (y - ref).abs().max().item()
        print(shape, prec, f"{err:.3e}", "fused" if nsm.lib().nsm_unet_fused_decoder() else "staged")
        assert y.shape == ref.shape and err <= tol, (shape, prec, err)
print("ok")
