import sys, os
sys.path.insert(0,'pcss-unet_b200'); sys.path.insert(0,'.')
import torch, nsm
from Unetmodel import Unet
from customLoss import CustomLoss
net = Unet(precision="bf16").cuda().train()
crit = CustomLoss("cuda")
x = torch.randn(32,4,512,512,device="cuda"); t = torch.rand(32,1,512,512,device="cuda")
for i in range(3):
    out = net(x); loss = crit(out, t, x); loss.backward()
torch.cuda.synchronize()
nsm.profile_enable(True)
out = net(x); loss = crit(out, t, x); loss.backward()
torch.cuda.synchronize()
rows = nsm.profile_read()
for name, ms, fl, by in rows:
    if fl > 0: print(f"{name:40s} {ms:8.3f} ms {fl/ms/1e9:8.1f} TF/s")
    else: print(f"{name:40s} {ms:8.3f} ms {by/ms/1e6:8.1f} GB/s")
