import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/convnet-approximater_b200")
import torch
from approx import backend as B
from approx.layers import LowRankExpConvV1
torch.manual_seed(0)
for (C, N, d, pad, H) in [(64, 192, 5, 2, 27), (192, 384, 3, 1, 13), (384, 256, 3, 1, 13)]:
    for dec in (False, True):
        layer = LowRankExpConvV1(C, N, d, 1, pad, 8).cuda().eval()
        if dec: layer.decomp()
        layer = layer.bfloat16()
        x = torch.randn(512, H, H, C, device="cuda").bfloat16()
        with torch.no_grad():
            for _ in range(2): layer.forward_nhwc(x)
        torch.cuda.synchronize()
