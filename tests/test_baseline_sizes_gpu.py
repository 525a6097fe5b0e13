"""Parity at the sizes BASELINE.json / SURVEY §8(d) name, CUDA path against the CPU ORACLE (never against this repo's own
generic kernels):

  config 1  one 3x3 s1 p1 scheme-2 layer, B = 8, 56 x 56, (C, N) in {(64, 64), (256, 256)}, K in {4, 8, 16, 32, 64}:
            fp32 compare4d < 1e-5 (metric of the reference's tests/test_low_rank_exp.py:5-14) and bf16 < 1e-3
  config 2  each of the 13 VGG-16 layer shapes at full H (224 ... 14), B = 2, K = 64, with the ReLU / 2x2 max-pool epilogue
            exactly as the trunk runs that layer
  config 4  MSCA original / MscaRep d0+fix / d1+fix at the four MSCAN-T (C, H) = (32,128) (64,64) (160,32) (256,16), B = 2
  scheme 1  the reference's AlexNet conv2-5 shapes (configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py: M = 8) at its
            profiling batch 64, grouped and separable (after decomp()) forms

bf16 tolerance 1e-3 is against the oracle run in bf16 on identical bf16 inputs and weights."""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu


def _factors(C, K, N, seed, dt=torch.float32):
    g = torch.Generator().manual_seed(seed)
    v_w = (torch.randn(K, C, 3, 1, generator=g) / (3 * C) ** 0.5).to(dt)
    h_w = (torch.randn(N, K, 1, 3, generator=g) / (3 * K) ** 0.5).to(dt)
    b = torch.randn(N, generator=g).to(dt)
    return v_w, h_w, b


# ------------------------------------------------------------------------------------------------ config 1
@pytest.mark.parametrize("K", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("C,N", [(64, 64), (256, 256)])
def test_config1_fp32_module_vs_oracle(C, N, K):
    """fp32: the module's exact path (two FFMA tap-GEMM launches) against two stock fp32 convs, compare4d < 1e-5."""
    from approx.layers import LowRankExpConvV2
    v_w, h_w, b = _factors(C, K, N, 100 + K)
    x = torch.randn(8, C, 56, 56, generator=torch.Generator().manual_seed(42))
    ref = O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))
    m = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N)
    m.v_conv.weight.data = v_w
    m.h_conv.weight.data = h_w
    m.h_conv.bias.data = b
    with torch.no_grad():
        y = m.cuda()(x.cuda())
    assert y.dtype == torch.float32 and y.shape == ref.shape
    assert O.compare4d(ref, y.cpu()) < 1e-5


@pytest.mark.parametrize("K", [4, 8, 16, 32, 64])
@pytest.mark.parametrize("C,N", [(64, 64), (256, 256)])
def test_config1_bf16_fused_vs_bf16_oracle(C, N, K):
    """bf16: the fused tcgen05 kernel (module path) against the oracle in bf16 on identical inputs, compare4d < 1e-3."""
    from approx import backend as B
    from approx.layers import LowRankExpConvV2
    v_w, h_w, b = _factors(C, K, N, 200 + K, torch.bfloat16)
    x = torch.randn(8, C, 56, 56, generator=torch.Generator().manual_seed(42)).bfloat16()
    ref = O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))
    m = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).bfloat16()
    m.v_conv.weight.data = v_w
    m.h_conv.weight.data = h_w
    m.h_conv.bias.data = b
    assert B.lowrank2_fused_supported(56, 56, C, K, N, 3)
    with torch.no_grad():
        y = m.cuda()(x.cuda())
    assert O.compare4d(ref.float(), y.float().cpu()) < 1e-3


# ------------------------------------------------------------------------------------------------ config 2
VGG = [(3, 64, 224), (64, 64, 224), (64, 128, 112), (128, 128, 112), (128, 256, 56), (256, 256, 56), (256, 256, 56),
       (256, 512, 28), (512, 512, 28), (512, 512, 28), (512, 512, 14), (512, 512, 14), (512, 512, 14)]
POOLED = {1, 3, 6, 9, 12}


@pytest.mark.parametrize("i", range(13))
def test_vgg_layer_shape_full_size_vs_bf16_oracle(i):
    """Layer i of the trunk at its full BASELINE spatial size, rank 64, ReLU (+ max-pool where VGG has one) in the epilogue."""
    from approx import backend as B
    C, N, H = VGG[i]
    v_w, h_w, b = _factors(C, 64, N, 300 + i, torch.bfloat16)
    x = torch.randn(2, C, H, H, generator=torch.Generator().manual_seed(7 + i)).bfloat16()
    ref = torch.relu(O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1)))
    if i in POOLED:
        ref = F.max_pool2d(ref.float(), 2)
    wv, wh = B.lowrank2_pack(v_w.cuda().float(), h_w.cuda().float())
    xn = B.to_nhwc(x.cuda(), B.round_up(C, 16))
    y = B.lowrank2_fused(xn, wv, wh, b.float().cuda(), relu=True, pool=i in POOLED)
    got = y.permute(0, 3, 1, 2).float().cpu()
    assert got.shape == ref.shape
    assert O.compare4d(ref.float(), got) < 1e-3


def test_vgg_trunk_layers_chain_at_224_vs_oracle():
    """The whole 13-layer trunk on ONE 224 x 224 image, every layer's input taken from the ORACLE's previous output so rounding
    differences do not compound: each layer of the module chain (layout kernel, fused convs, fused pools) at < 1e-3."""
    from approx import backend as B
    from approx.models import VGG16Scheme2Trunk
    torch.manual_seed(3)
    trunk = VGG16Scheme2Trunk(64).bfloat16()
    with torch.no_grad():
        for m in trunk.conv_layers():
            m.v_conv.weight.normal_(0, (1.0 / (3 * m.in_channels)) ** 0.5)
            m.h_conv.weight.normal_(0, (2.0 / (3 * 64)) ** 0.5)
    x = torch.randn(1, 3, 224, 224).bfloat16()
    cpu = [(m.v_conv.weight.detach().clone(), m.h_conv.weight.detach().clone(), m.h_conv.bias.detach().clone())
           for m in trunk.conv_layers()]
    trunk = trunk.cuda().eval()
    ref = x
    with torch.no_grad():
        for i, m in enumerate(trunk.features):
            xin = B.to_nhwc(ref.cuda(), B.round_up(ref.shape[1], 16))
            got = m.forward_nhwc(xin).permute(0, 3, 1, 2).float().cpu()
            ref = torch.relu(O.lowrank2_forward(ref, *cpu[i], (1, 1), (1, 1)))
            if i in POOLED:
                ref = F.max_pool2d(ref.float(), 2).bfloat16()
            assert O.compare4d(ref.float(), got) < 1e-3, i


# ------------------------------------------------------------------------------------------------ config 4
def _oracle_msca_original(m, x, dt):
    sd = {k: v.detach().cpu().to(dt) for k, v in m.state_dict().items()}
    br = [tuple(sd[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    return O.msca_forward(x.to(dt), sd["conv0.weight"], sd["conv0.bias"], lambda a: O.parallel_forward(a, br, True),
                          sd["channel_mix.weight"], sd["channel_mix.bias"]), sd, br


def _oracle_msca_rep(sd, br, x, C, decomp, dt):
    """Oracle forward of the re-parameterised block: the reference's kernel merge / SVD split (fp32, as MscaRep runs it), then
    the block in dtype dt."""
    w, b, r = O.msca_equivalent_kernel([tuple(t.float() for t in q) for q in br], C)
    if decomp == 0:
        w, b, r = w.to(dt), b.to(dt), r.to(dt)
        sdf = lambda a: O.fix_padding_bias(F.conv2d(a, w, b, padding=10, groups=C), r, 10)
    else:
        fac, _ = O.msca_rep_factors(w, decomp)
        brs = [(c1.to(dt), None, c2.to(dt), b.to(dt) if j == decomp - 1 else None) for j, (c1, c2) in enumerate(fac)]
        r = r.to(dt)
        sdf = lambda a: O.fix_padding_bias(O.parallel_forward(a, brs, False), r, 10)
    return O.msca_forward(x.to(dt), sd["conv0.weight"], sd["conv0.bias"], sdf, sd["channel_mix.weight"], sd["channel_mix.bias"])


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,H", [(32, 128), (64, 64), (160, 32), (256, 16)])
def test_mscan_t_shapes_original_and_rep_vs_oracle(C, H, dt):
    """MSCA at each MSCAN-T stage shape (512 x 512 input, stem stride 4), B = 2: original block, MscaRep(decomp=0, fix) and
    MscaRep(decomp=1, fix), each against the CPU oracle in the same dtype."""
    from approx.core import MscaRep
    from approx.layers import MSCA
    torch.manual_seed(C + H)
    src = MSCA(C, 5, (7, 11, 21))
    x = torch.randn(2, C, H, H)
    ref, sd, br = _oracle_msca_original(src, x, dt)
    tol = 1e-5 if dt == torch.float32 else 1e-3
    with torch.no_grad():
        y = src.cuda().to(dt).eval()(x.cuda().to(dt))
    assert O.compare4d(ref.float(), y.float().cpu()) < tol
    src = src.cpu().float()
    for decomp in (0, 1):
        app = MscaRep(decomp=decomp, fix=True)
        sub = app.initialize(src)
        with contextlib.redirect_stdout(io.StringIO()):
            app.optimize(sub)
        tgt = app.postprocess(sub)
        # the oracle re-parameterises the fp32 source itself (reference algebra on the CPU), then runs in dt
        sd32 = {k: v.detach().cpu() for k, v in src.state_dict().items()}
        br32 = [tuple(sd32[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
                for i in range(3)]
        sdd = {k: v.to(dt) for k, v in sd32.items()}
        ref_r = _oracle_msca_rep(sdd, br32, x, C, decomp, dt)
        with torch.no_grad():
            y_r = tgt.cuda().to(dt).eval()(x.cuda().to(dt))
        # fp32: 5e-5 (the device SVD's factors differ from LAPACK's in the last bits; the reconstruction, not the factors, is pinned)
        assert O.compare4d(ref_r.float(), y_r.float().cpu()) < (5e-5 if dt == torch.float32 else 2e-3), (decomp, dt)


# ------------------------------------------------------------------------------------------------ scheme 1 (AlexNet conv2-5)
ALEX = [(64, 192, 5, 27), (192, 384, 3, 13), (384, 256, 3, 13), (256, 256, 3, 13)]


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("C,N,d,H", ALEX)
def test_alexnet_scheme1_layers_at_profiling_batch_vs_bf16_oracle(C, N, d, H, sep):
    """LowRankExpConvV1 (M = 8) at the reference's profiling batch 64 (scripts/low_rank_exp_spr.py:35-63): fused tcgen05 kernel
    through the module against the bf16 oracle; `sep`: after decomp() (SeparableConv form)."""
    from approx.layers import LowRankExpConvV1
    torch.manual_seed(C + d)
    m = LowRankExpConvV1(C, N, d, 1, d // 2, num_base=8)
    if sep:
        m = m.cuda()
        m.decomp()
        m = m.cpu()
    m = m.bfloat16()
    x = torch.randn(64, C, H, H).bfloat16()
    with torch.no_grad():
        if sep:
            ref = O.lowrank1_forward_decomp(x, m.s_conv.v_conv.weight, m.s_conv.h_conv.weight, m.d_conv.weight, m.d_conv.bias,
                                            (1, 1), (d // 2, d // 2))
        else:
            ref = O.lowrank1_forward(x, m.s_conv.weight, m.s_conv.bias, m.d_conv.weight, m.d_conv.bias, (1, 1), (d // 2, d // 2))
        y = m.cuda().eval()(x.cuda())
    assert O.compare4d(ref.float(), y.float().cpu()) < 1e-3
