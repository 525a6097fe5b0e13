#!/usr/bin/env python
"""Golden vectors of one MSCAN block's attention half, made by IMPORTING the reference (run in the build container only):

    python tests/golden/make_golden_mscan_block.py        ->  tests/golden/mscan_block.npz

The reference's `MultiScaleConvAttnModule` (src/approx/models/mscan.py:83-112) in eval mode with random BatchNorm statistics and a
layer scale large enough for the branch to matter:   y_attn = x + layer_scale_1 * attn(norm1(x))   (mscan.py:108-109), fp32 and bf16."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402


def main():
    G._install_stubs()
    sys.path.insert(0, G.REF)
    from approx.models.mscan import MultiScaleConvAttnModule
    torch.manual_seed(11)
    C = 32
    blk = MultiScaleConvAttnModule(num_channel=C, hidden_channel=2 * C, drop=0., drop_path=0.).eval()
    with torch.no_grad():
        for bn in (blk.norm1, blk.norm2):
            bn.running_mean.copy_(0.3 * torch.randn(C))
            bn.running_var.copy_(0.5 + torch.rand(C))
            bn.weight.copy_(1.0 + 0.2 * torch.randn(C))
            bn.bias.copy_(0.1 * torch.randn(C))
        blk.layer_scale_1.copy_(0.25 + 0.5 * torch.rand(C))
        blk.layer_scale_2.copy_(0.25 + 0.5 * torch.rand(C))
    x = torch.randn(2, C, 20, 22)
    out = {"x": x.numpy()}
    out.update({k: v.detach().cpu().float().numpy() for k, v in blk.state_dict().items()})
    with torch.no_grad():
        out["y_attn"] = (x + blk.layer_scale_1.unsqueeze(-1).unsqueeze(-1) * blk.attn(blk.norm1(x))).numpy()
        out["y_block"] = blk(x).numpy()
        bb = blk.to(torch.bfloat16)
        xb = x.to(torch.bfloat16)
        out["y_attn_bf16"] = (xb + bb.layer_scale_1.unsqueeze(-1).unsqueeze(-1) * bb.attn(bb.norm1(xb))).float().numpy()
    np.savez_compressed(os.path.join(HERE, "mscan_block.npz"), **out)
    print("written", os.path.join(HERE, "mscan_block.npz"), {k: v.shape for k, v in out.items() if k.startswith(("x", "y_"))})


if __name__ == "__main__":
    main()
