"""Parity of the fused tcgen05 scheme-2 kernel: against the oracle run in bf16 on identical bf16
inputs/weights (tolerance compare4d <= 1e-3, BASELINE.json) and against the generic CUDA path."""
import pytest
import torch

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # B, C, H, W, K, N, relu
    (2, 64, 56, 56, 16, 64, False),     # config-1 style single layer
    (2, 64, 56, 56, 64, 64, True),
    (2, 3, 32, 30, 8, 64, False),       # first layer: C=3 padded to 16, K=8 padded to 16
    (1, 128, 20, 37, 32, 128, True),    # ragged H/W: edge tiles clip
    (2, 256, 14, 14, 64, 256, False),   # W=14: single column tile, 4 channel chunks
    (1, 512, 14, 14, 64, 512, True),    # weights do not fit: streamed mode
    (3, 96, 9, 5, 4, 192, False),       # C=96 -> CK=32, tiny image
    (1, 48, 17, 15, 16, 64, False),     # C=48 -> CK=16 x 3 chunks
]


def make(C, K, N, seed):
    g = torch.Generator().manual_seed(seed)
    v_w = (torch.randn(K, C, 3, 1, generator=g) / (3 * C) ** 0.5).bfloat16()
    h_w = (torch.randn(N, K, 1, 3, generator=g) / (3 * K) ** 0.5).bfloat16()
    b = torch.randn(N, generator=g).bfloat16()
    return v_w, h_w, b


@pytest.mark.parametrize("case", CASES)
def test_fused_matches_bf16_oracle(case):
    from approx import backend as B
    Bn, C, H, W, K, N, relu = case
    v_w, h_w, b = make(C, K, N, 11)
    x = torch.randn(Bn, C, H, W, generator=torch.Generator().manual_seed(5)).bfloat16()
    ref = O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))
    if relu:
        ref = torch.relu(ref)
    assert B.lowrank2_fused_supported(H, W, C, K, N, 3)
    wv, wh = B.lowrank2_pack(v_w.cuda().float(), h_w.cuda().float())
    xn = B.to_nhwc(x.cuda(), B.round_up(C, 16))
    y = B.lowrank2_fused(xn, wv, wh, b.float().cuda(), relu=relu)
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float().cpu()
    err = O.compare4d(ref.float(), got)
    assert err < 1e-3, f"rel err {err}"
    # generic CUDA path (two tap-GEMM launches, T through HBM) must agree too
    wv_g = B.pack_taps(v_w.cuda().float(), torch.bfloat16)
    wh_g = B.pack_taps(h_w.cuda().float(), torch.bfloat16)
    t = B.conv_taps(B.to_nhwc(x.cuda()), wv_g, None, K, 3, 0, (1, 1), 1)
    y2 = B.conv_taps(t, wh_g, b.float().cuda(), N, 3, 1, (1, 1), 1, relu=relu)
    assert O.compare4d(y2.permute(0, 3, 1, 2).float().cpu(), got) < 1e-3


def test_module_takes_fused_path_and_matches_oracle():
    from approx import backend as B
    from approx.layers import LowRankExpConvV2
    torch.manual_seed(0)
    m = LowRankExpConvV2(64, 16, (3, 3), (1, 1), (1, 1), out_channels=128).cuda().bfloat16()
    x = torch.randn(8, 64, 56, 56).bfloat16()
    before = B.launch_count()
    y = m(x.cuda())
    assert B.launch_count() - before <= 3          # (layout) + pack + ONE fused conv kernel
    ref = O.lowrank2_forward(x, m.v_conv.weight.cpu(), m.h_conv.weight.cpu(), m.h_conv.bias.cpu(), (1, 1), (1, 1))
    assert y.shape == ref.shape
    assert O.compare4d(ref.float(), y.float().cpu()) < 1e-3


def test_linearity_at_full_size():
    """Size-independent property at a BASELINE-sized layer: f(2x) - bias == 2(f(x) - bias) (exact in bf16
    up to the final rounding: powers of two commute with every rounding step)."""
    from approx import backend as B
    C, K, N = 64, 64, 64
    v_w, h_w, _ = make(C, K, N, 3)
    wv, wh = B.lowrank2_pack(v_w.cuda().float(), h_w.cuda().float())
    x = torch.randn(16, 224, 224, C, device="cuda").bfloat16()
    y1 = B.lowrank2_fused(x, wv, wh, None)
    y2 = B.lowrank2_fused(x * 2, wv, wh, None)
    assert torch.equal(y2, y1 * 2)
    # and a tile-boundary check against the generic path on a slice
    wv_g = B.pack_taps(v_w.cuda().float(), torch.bfloat16)
    wh_g = B.pack_taps(h_w.cuda().float(), torch.bfloat16)
    t = B.conv_taps(x[:2].contiguous(), wv_g, None, K, 3, 0, (1, 1), 1)
    yg = B.conv_taps(t, wh_g, None, N, 3, 1, (1, 1), 1)
    assert O.compare4d(yg.permute(0, 3, 1, 2).float().cpu(), y1[:2].permute(0, 3, 1, 2).float().cpu()) < 1e-3


@pytest.mark.parametrize("case", [(2, 64, 28, 28, 32, 64), (1, 128, 14, 42, 64, 128), (2, 16, 16, 20, 16, 64)])
def test_fused_relu_maxpool_epilogue(case):
    """ReLU + 2x2 max-pool fused into the epilogue == max_pool2d(relu(layer(x))) of the bf16 oracle (max is exact)."""
    import torch.nn.functional as F
    from approx import backend as B
    Bn, C, H, W, K, N = case
    v_w, h_w, b = make(C, K, N, 21)
    x = torch.randn(Bn, C, H, W, generator=torch.Generator().manual_seed(9)).bfloat16()
    ref = F.max_pool2d(torch.relu(O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))).float(), 2)
    wv, wh = B.lowrank2_pack(v_w.cuda().float(), h_w.cuda().float())
    xn = B.to_nhwc(x.cuda(), B.round_up(C, 16))
    y = B.lowrank2_fused(xn, wv, wh, b.float().cuda(), relu=True, pool=True)
    assert y.shape == (Bn, H // 2, W // 2, N)
    got = y.permute(0, 3, 1, 2).float().cpu()
    assert O.compare4d(ref, got) < 1e-3
    full = B.lowrank2_fused(xn, wv, wh, b.float().cuda(), relu=True)
    assert torch.equal(B.maxpool2x2(full), y)          # bit-identical to pooling the un-fused output


def test_vgg_trunk_matches_oracle_small():
    """13-layer trunk (rank 16) on a 2 x 3 x 32 x 32 input against the oracle run in bf16."""
    import torch.nn.functional as F
    from approx.models import VGG16Scheme2Trunk, VGG16_CFG
    torch.manual_seed(0)
    trunk = VGG16Scheme2Trunk(16).bfloat16()
    with torch.no_grad():
        for m in trunk.conv_layers():
            m.v_conv.weight.normal_(0, (1.0 / (3 * m.in_channels)) ** 0.5)
            m.h_conv.weight.normal_(0, (2.0 / (3 * 16)) ** 0.5)
    x = torch.randn(2, 3, 32, 32).bfloat16()
    ref = x
    it = iter(trunk.conv_layers())
    for c in VGG16_CFG:
        if c == "M":
            ref = F.max_pool2d(ref.float(), 2).bfloat16()
        else:
            m = next(it)
            ref = torch.relu(O.lowrank2_forward(ref, m.v_conv.weight.detach(), m.h_conv.weight.detach(),
                                                m.h_conv.bias.detach(), (1, 1), (1, 1)))
    y = trunk.cuda()(x.cuda())
    assert y.shape == ref.shape == (2, 512, 1, 1)
    assert O.compare4d(ref.float(), y.float().cpu()) < 2e-2      # 13 chained bf16 layers: rounding differences compound


@pytest.mark.parametrize("case", [(2, 64, 20, 30, 64, 64), (1, 128, 9, 17, 32, 128), (2, 16, 28, 28, 16, 64)])
def test_fused_forward_also_writes_T(case):
    """cna_lowrank2_fused_forward_t: y identical to the plain fused call, T == v_conv(x) rounded to bf16."""
    import torch.nn.functional as F
    from approx import backend as B
    Bn, C, H, W, K, N = case
    v_w, h_w, b = make(C, K, N, 5)
    x = torch.randn(Bn, C, H, W, generator=torch.Generator().manual_seed(3)).bfloat16()
    wv, wh = B.lowrank2_pack(v_w.cuda().float(), h_w.cuda().float())
    xn = B.to_nhwc(x.cuda(), B.round_up(C, 16))
    y0 = B.lowrank2_fused(xn, wv, wh, b.float().cuda())
    y1, t = B.lowrank2_fused_with_t(xn, wv, wh, b.float().cuda())
    assert torch.equal(y0, y1)
    t_ref = F.conv2d(x, v_w, None, 1, (1, 0))
    assert O.compare4d(t_ref.float(), t[..., :K].permute(0, 3, 1, 2).float().cpu()) < 1e-3


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("case", [(2, 64, 20, 30, 64), (1, 256, 9, 17, 64), (2, 128, 28, 14, 32)])
def test_tapconv3_tcgen05(case, axis):
    """3-tap C -> K conv on tcgen05 along H (axis 0) and along W (axis 1, transposed tensor maps) vs torch bf16."""
    import torch.nn.functional as F
    from approx import backend as B
    Bn, C, H, W, K = case
    g = torch.Generator().manual_seed(17)
    x = torch.randn(Bn, C, H, W, generator=g).bfloat16()
    w = (torch.randn(K, C, 3, 1, generator=g) / (3 * C) ** 0.5).bfloat16()
    ref = F.conv2d(x, w if axis == 0 else w.permute(0, 1, 3, 2), None, 1, (1, 0) if axis == 0 else (0, 1))
    wv, _ = B.lowrank2_pack(w.cuda().float(), torch.zeros(64, K, 1, 3, device="cuda"))
    out = B.tapconv3_tc(B.to_nhwc(x.cuda()).contiguous(), wv, axis)
    assert O.compare4d(ref.float(), out[..., :K].permute(0, 3, 1, 2).float().cpu()) < 1e-3


@pytest.mark.parametrize("channels_last_out", [False, True])
def test_host_pipeline_matches_direct_call(channels_last_out):
    """approx.pipeline.HostPipeline (pinned host buffers, 3 streams, ring of device inputs): every queued batch comes back
    bit-identical to `module(x.cuda())`, for an NCHW-contiguous and for a channels_last host output buffer."""
    from approx.layers import LowRankExpConvV2
    from approx.pipeline import HostPipeline
    torch.manual_seed(3)
    m = LowRankExpConvV2(3, 16, (3, 3), (1, 1), (1, 1), out_channels=64).cuda().bfloat16().eval()
    xs = [torch.randn(4, 3, 40, 36).bfloat16().pin_memory() for _ in range(5)]
    fmt = torch.channels_last if channels_last_out else torch.contiguous_format
    ys = [torch.empty(4, 64, 40, 36, dtype=torch.bfloat16).contiguous(memory_format=fmt).pin_memory() for _ in xs]
    pipe = HostPipeline(m, "cuda:0", depth=2)
    for x, y in zip(xs, ys):
        pipe.submit(x, y)
    pipe.wait()
    with torch.no_grad():
        for x, y in zip(xs, ys):
            ref = m(x.cuda()).cpu()
            assert torch.equal(ref, y)


def test_uint8_image_input_is_normalised_in_the_layout_kernel():
    """uint8 NCHW images (what the reference's timm loaders ship to the GPU, hooks/fps.py:38-49) through
    cna_u8_nchw_to_nhwc_norm == the same images normalised with (x - 255 mean) / (255 std) in fp32 and rounded to bf16."""
    from approx import backend as B
    from approx.models import VGG16Scheme2Trunk
    torch.manual_seed(1)
    x = torch.randint(0, 256, (3, 3, 40, 36), dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    mean255 = torch.tensor([m * 255.0 for m in mean]); inv = torch.tensor([1.0 / (s * 255.0) for s in std])
    want = ((x.float() - mean255.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)).bfloat16()
    got = B.u8_to_nhwc_norm(x.cuda(), mean255.cuda(), inv.cuda(), 16)
    assert got.shape == (3, 40, 36, 16)
    assert torch.equal(got[..., :3].permute(0, 3, 1, 2).cpu(), want) and float(got[..., 3:].abs().max()) == 0.0
    trunk = VGG16Scheme2Trunk(16).cuda().bfloat16().eval()
    xi = torch.randint(0, 256, (2, 3, 32, 32), dtype=torch.uint8)
    with torch.no_grad():
        y_u8 = trunk(xi.cuda())
        y_bf = trunk(((xi.float() - mean255.view(1, 3, 1, 1)) * inv.view(1, 3, 1, 1)).bfloat16().cuda())
    assert torch.equal(y_u8, y_bf)
