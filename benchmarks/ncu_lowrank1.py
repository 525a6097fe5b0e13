import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/convnet-approximater_b200")
import torch
from approx.layers import LowRankExpConvV1
torch.manual_seed(0)
layer = LowRankExpConvV1(192, 384, 3, 1, 1, 8).cuda().eval().bfloat16()
x = torch.randn(512, 13, 13, 192, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(3): layer.forward_nhwc(x)
torch.cuda.synchronize()
