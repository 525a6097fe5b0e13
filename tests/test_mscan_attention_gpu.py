"""the attention half of an MSCAN block fused around MSCA (approx.layers.MSCANAttention, four kernels) against the
oracle restatement of src/approx/models/mscan.py:72-80, 108-109 and the golden vectors of the reference's own module."""
import pytest
import torch

from oracle import approx_oracle as O

T = torch.tensor


def _oracle_from_golden(g, dt, x):
    f = lambda k: T(g[k]).to(dt)
    br = [tuple(f(f"attn.spatial_gating_unit.sd_convs.branches.{i}.{n}") for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    msca = lambda a: O.msca_forward(a, f("attn.spatial_gating_unit.conv0.weight"), f("attn.spatial_gating_unit.conv0.bias"),
                                    lambda t: O.parallel_forward(t, br, True), f("attn.spatial_gating_unit.channel_mix.weight"),
                                    f("attn.spatial_gating_unit.channel_mix.bias"))
    return O.mscan_attention_forward(x.to(dt), f("norm1.weight"), f("norm1.bias"), f("norm1.running_mean"), f("norm1.running_var"), 1e-5,
                                     f("attn.proj_1.weight"), f("attn.proj_1.bias"), msca, f("attn.proj_2.weight"), f("attn.proj_2.bias"),
                                     f("layer_scale_1"))


def test_oracle_mscan_attention_matches_reference_golden(golden):
    """CPU: the restatement reproduces the reference module's attention half (fp32)."""
    g = golden("mscan_block")
    y = _oracle_from_golden(g, torch.float32, T(g["x"]))
    assert O.compare4d(T(g["y_attn"]), y) < 1e-5


def _build_block(g, C):
    """The attention half assembled from this backend's layers with the reference block's state_dict keys."""
    from torch import nn
    from approx.layers import MSCA, MSCANAttention

    class Attn(nn.Module):            # same attribute names as SpatialAttention (mscan.py:65-70)
        def __init__(self):
            super().__init__()
            self.proj_1 = nn.Conv2d(C, C, 1)
            self.activation = nn.GELU()
            self.spatial_gating_unit = MSCA(C, 5, (7, 11, 21))
            self.proj_2 = nn.Conv2d(C, C, 1)

    m = MSCANAttention(nn.BatchNorm2d(C), Attn(), nn.Parameter(torch.ones(C)))
    sd = {k: T(v) for k, v in g.items() if k.startswith(("norm1.", "attn.", "layer_scale_1"))}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("num_batches_tracked" in k for k in missing), (missing, unexpected)
    return m


@pytest.mark.gpu
def test_fused_mscan_attention_vs_oracle_and_reference_golden(golden):
    """bf16 fused block vs (a) the fp32 oracle and (b) the reference's bf16 module output.  Tolerances: the block output is
    x + layer_scale * branch; the fused kernels round the BRANCH at different points than the module graph (BatchNorm folded into
    proj_1, one rounding for proj_2 + shortcut + scale + residual), so the branch agrees to bf16 resolution (rel-L2 < 1e-2) and the
    block output to < 4e-3 (one bf16 rounding of an O(1) tensor is 2e-3 rel-L2 at most)."""
    from approx import backend as B
    g = golden("mscan_block")
    x = T(g["x"])
    C = x.shape[1]
    m = _build_block(g, C).cuda().bfloat16().eval()
    n0 = B.launch_count()
    y = m(x.cuda().bfloat16()).float().cpu()
    first = B.launch_count() - n0
    n0 = B.launch_count()
    m(x.cuda().bfloat16())
    assert B.launch_count() - n0 <= 6, (first, B.launch_count() - n0)      # 2 layout kernels + proj_1 | sd | mix+gate | proj_2
    ref32 = _oracle_from_golden(g, torch.float32, x)
    xb = x.bfloat16().float()
    assert O.compare4d(ref32, y) < 4e-3
    assert O.compare4d(T(g["y_attn_bf16"]), y) < 4e-3
    assert O.compare4d(ref32 - xb, y - xb) < 1e-2                          # the branch itself


@pytest.mark.gpu
@pytest.mark.parametrize("C,H", [(64, 33), (160, 16)])
def test_fused_mscan_attention_random_shapes_vs_oracle(C, H):
    from torch import nn
    from approx.layers import MSCA, MSCANAttention
    torch.manual_seed(C)

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.proj_1 = nn.Conv2d(C, C, 1)
            self.spatial_gating_unit = MSCA(C, 5, (7, 11, 21))
            self.proj_2 = nn.Conv2d(C, C, 1)

    bn = nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.running_mean.copy_(0.3 * torch.randn(C)); bn.running_var.copy_(0.5 + torch.rand(C))
        bn.weight.copy_(1 + 0.2 * torch.randn(C)); bn.bias.copy_(0.1 * torch.randn(C))
    m = MSCANAttention(bn, Attn(), nn.Parameter(0.25 + 0.5 * torch.rand(C))).eval()
    x = torch.randn(2, C, H, H + 3)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    br = [tuple(sd[f"attn.spatial_gating_unit.sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    msca = lambda a: O.msca_forward(a, sd["attn.spatial_gating_unit.conv0.weight"], sd["attn.spatial_gating_unit.conv0.bias"],
                                    lambda t: O.parallel_forward(t, br, True), sd["attn.spatial_gating_unit.channel_mix.weight"],
                                    sd["attn.spatial_gating_unit.channel_mix.bias"])
    ref = O.mscan_attention_forward(x, sd["norm1.weight"], sd["norm1.bias"], sd["norm1.running_mean"], sd["norm1.running_var"], bn.eps,
                                    sd["attn.proj_1.weight"], sd["attn.proj_1.bias"], msca, sd["attn.proj_2.weight"],
                                    sd["attn.proj_2.bias"], sd["layer_scale_1"])
    y = m.cuda().bfloat16()(x.cuda().bfloat16()).float().cpu()
    xb = x.bfloat16().float()
    assert O.compare4d(ref, y) < 4e-3
    assert O.compare4d(ref - xb, y - xb) < 1e-2
    with pytest.raises(RuntimeError):
        m.train()(x.cuda().bfloat16())
