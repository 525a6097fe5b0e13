"""the FFN half of an MSCAN block (approx.layers.MSCANFFN: fc1 slabs, depthwise 3x3 + GELU, fc2 with the layer scale and
the residual in its epilogue) and the whole block (approx.layers.MSCANBlock) against the oracle restatement of
src/approx/models/mscan.py:56-62, 108-112 and the golden output of the reference's own MultiScaleConvAttnModule."""
import pytest
import torch

from oracle import approx_oracle as O
from test_mscan_attention_gpu import _oracle_from_golden

T = torch.tensor


def _oracle_ffn(sd, x, eps=1e-5):
    return O.mscan_ffn_forward(x, sd["norm2.weight"], sd["norm2.bias"], sd["norm2.running_mean"], sd["norm2.running_var"], eps,
                               sd["mlp.fc1.weight"], sd["mlp.fc1.bias"], sd["mlp.dconv.weight"], sd["mlp.dconv.bias"],
                               sd["mlp.fc2.weight"], sd["mlp.fc2.bias"], sd["layer_scale_2"])


def test_oracle_mscan_block_matches_reference_golden(golden):
    """CPU: attention half followed by the FFN half reproduces the reference module's block output (fp32)."""
    g = golden("mscan_block")
    sd = {k: T(g[k]) for k in g if k not in ("x", "y_attn", "y_block", "y_attn_bf16")}
    y1 = _oracle_from_golden(g, torch.float32, T(g["x"]))
    assert O.compare4d(T(g["y_block"]), _oracle_ffn(sd, y1)) < 1e-5


def _block(C, Hd):
    from torch import nn
    from approx.layers import MSCA, MSCANBlock

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.proj_1 = nn.Conv2d(C, C, 1)
            self.spatial_gating_unit = MSCA(C, 5, (7, 11, 21))
            self.proj_2 = nn.Conv2d(C, C, 1)

    class FFN(nn.Module):             # same attribute names as the reference's FFN (mscan.py:27-54)
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Conv2d(C, Hd, 1)
            self.dconv = nn.Conv2d(Hd, Hd, 3, 1, 1, groups=Hd)
            self.fc2 = nn.Conv2d(Hd, C, 1)

    return MSCANBlock(nn.BatchNorm2d(C), Attn(), nn.Parameter(torch.ones(C)), nn.BatchNorm2d(C), FFN(), nn.Parameter(torch.ones(C)))


@pytest.mark.gpu
def test_fused_mscan_block_vs_reference_golden(golden):
    """bf16 fused block (8 kernels + 2 layout kernels) vs the reference's fp32 block output.  Tolerances as for the attention half: the
    output is x + scale * branch twice over, every kernel rounds to bf16 once: block < 5e-3 rel-L2, the two branches together < 1.5e-2."""
    from approx import backend as B
    g = golden("mscan_block")
    x = T(g["x"])
    C = x.shape[1]
    m = _block(C, 2 * C)
    sd = {k: T(g[k]) for k in g if k not in ("x", "y_attn", "y_block", "y_attn_bf16")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing), (missing, unexpected)
    m = m.cuda().bfloat16().eval()
    m(x.cuda().bfloat16())
    n0 = B.launch_count()
    y = m(x.cuda().bfloat16()).float().cpu()
    assert B.launch_count() - n0 <= 10
    xb = x.bfloat16().float()
    assert O.compare4d(T(g["y_block"]), y) < 5e-3
    assert O.compare4d(T(g["y_block"]) - xb, y - xb) < 1.5e-2


@pytest.mark.gpu
@pytest.mark.parametrize("C,Hd,H", [(64, 512, 19), (160, 640, 12), (256, 1024, 9), (32, 256, 33)])
def test_fused_mscan_ffn_slabs_vs_oracle(C, Hd, H):
    """The FFN half alone at the four MSCAN-T (C, hidden) pairs on small images: hidden > 256 takes several fc1 slabs, C >= 160 several
    K slabs of fc2 accumulated in place.  Against the fp32 oracle on the same bf16-rounded parameters."""
    from approx.layers import MSCANFFN
    torch.manual_seed(C)
    blk = _block(C, Hd)
    with torch.no_grad():
        bn = blk.norm2
        bn.running_mean.copy_(0.3 * torch.randn(C)); bn.running_var.copy_(0.5 + torch.rand(C))
        bn.weight.copy_(1 + 0.2 * torch.randn(C)); bn.bias.copy_(0.1 * torch.randn(C))
        blk.layer_scale_2.copy_(4.0 + 2.0 * torch.rand(C))       # with torch's default init the mlp output is ~0.02 x the input: a
    m = MSCANFFN.from_block(blk).eval()                          # scale of ~5 makes the branch big enough to see past y's bf16 rounding
    x = torch.randn(2, C, H, H + 2)
    sd = {k: v.detach().bfloat16().float() for k, v in blk.state_dict().items()}
    ref = _oracle_ffn(sd, x.bfloat16().float(), blk.norm2.eps)
    y = m.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu()
    xb = x.bfloat16().float()
    assert O.compare4d(ref, y) < 4e-3
    assert O.compare4d(ref - xb, y - xb) < 1.2e-2
    with pytest.raises(RuntimeError):
        m.train()(x.cuda().bfloat16())
