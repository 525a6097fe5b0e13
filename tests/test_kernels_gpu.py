"""GPU parity tests of the generic CUDA kernels (through the C ABI via approx.backend) against the
CPU oracle / stock torch CPU convs on the same seeded inputs.  fp32: compare4d < 1e-5 (the reference's
threshold, tests/test_low_rank_exp.py:14); bf16: compare4d < 1e-3 against the oracle run in bf16."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu

TOL32, TOL16 = 1e-5, 1e-3


@pytest.fixture(scope="module")
def B():
    from approx import backend
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    backend.lib()
    return backend


def dev(t, dt=torch.float32):
    return t.to("cuda", dt)


def nchw(y_nhwc):
    return y_nhwc.permute(0, 3, 1, 2).float().cpu()


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,c_pad", [((2, 3, 5, 7), 16), ((1, 33, 9, 4), 33), ((3, 64, 8, 8), 64)])
def test_layout_roundtrip(B, dt, shape, c_pad):
    torch.manual_seed(0)
    x = torch.randn(*shape).to(dt)
    xn = B.to_nhwc(x.cuda(), c_pad)
    assert xn.shape == (shape[0], shape[2], shape[3], c_pad)
    ref = torch.zeros(shape[0], shape[2], shape[3], c_pad, dtype=dt)
    ref[..., :shape[1]] = x.permute(0, 2, 3, 1)
    assert torch.equal(xn.cpu(), ref)
    back = B.to_nchw_contiguous(xn, shape[1])
    assert torch.equal(back.cpu(), x)
    # channels_last input is taken zero-copy
    xcl = x.cuda().contiguous(memory_format=torch.channels_last)
    assert B.to_nhwc(xcl).data_ptr() == xcl.data_ptr()


@pytest.mark.parametrize("shape,c_pad", [((4, 3, 224, 224), 16), ((2, 8, 9, 11), 8), ((3, 1, 17, 5), 16), ((2, 5, 6, 6), 8)])
@pytest.mark.parametrize("src_dt", [torch.float32, torch.bfloat16])
def test_image_layout_kernel(B, shape, c_pad, src_dt):
    """Few-channel NCHW images -> zero-padded bf16 NHWC (the first kernel of the host-buffer pipeline): exact re-layout."""
    torch.manual_seed(1)
    x = torch.randn(*shape).to(src_dt)
    xn = B.to_nhwc(x.cuda(), c_pad, dtype=torch.bfloat16)
    ref = torch.zeros(shape[0], shape[2], shape[3], c_pad, dtype=torch.bfloat16)
    ref[..., :shape[1]] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert xn.dtype == torch.bfloat16 and torch.equal(xn.cpu(), ref)


CONV_CASES = [
    # B, C, H, W, N, taps, axis, sh, sw, pad
    (2, 16, 12, 10, 8, 3, 0, 1, 1, 1),
    (2, 8, 12, 10, 24, 3, 1, 1, 1, 1),
    (3, 6, 11, 9, 4, 5, 0, 2, 1, 2),
    (3, 4, 11, 9, 7, 3, 1, 1, 2, 1),
    (2, 70, 7, 9, 130, 1, 1, 1, 1, 0),
    (1, 3, 20, 20, 64, 3, 0, 1, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_conv_taps_vs_torch(B, case, dt):
    Bn, C, H, W, N, taps, axis, sh, sw, pad = case
    torch.manual_seed(1)
    x = torch.randn(Bn, C, H, W).to(dt)
    w = (torch.randn(N, C, taps, 1) if axis == 0 else torch.randn(N, C, 1, taps)).to(dt)
    b = torch.randn(N).to(dt)
    stride = (sh, 1) if axis == 0 else (1, sw)
    padding = (pad, 0) if axis == 0 else (0, pad)
    ref = F.conv2d(x, w, b, stride, padding)
    wpk = B.pack_taps(w.cuda().float(), dt)
    y = B.conv_taps(B.to_nhwc(x.cuda()), wpk, b.float().cuda(), N, taps, axis, stride, pad)
    assert O.compare4d(ref.float(), nchw(y)) < (TOL32 if dt == torch.float32 else TOL16)
    # relu + gate epilogue
    g = torch.randn_like(ref)
    y2 = B.conv_taps(B.to_nhwc(x.cuda()), wpk, b.float().cuda(), N, taps, axis, stride, pad, relu=True,
                     gate=B.to_nhwc(g.cuda()).contiguous())
    ref2 = (torch.relu(ref) * g)
    assert O.compare4d(ref2.float(), nchw(y2)) < (TOL32 if dt == torch.float32 else 3e-3)


DW_CASES = [
    # B, C, H, W, mult, kh, kw, sh, sw, ph, pw
    (2, 8, 9, 11, 4, 3, 3, 1, 1, 1, 1),
    (2, 6, 13, 12, 3, 5, 5, 2, 2, 2, 2),
    (2, 8, 24, 26, 1, 1, 21, 1, 1, 0, 10),
    (2, 8, 7, 9, 1, 21, 1, 1, 1, 10, 0),
    (1, 5, 7, 9, 1, 21, 21, 1, 1, 10, 10),
    (2, 6, 10, 8, 2, 3, 1, 2, 1, 1, 0),
    # register-tiled kernels (dwconv_tile.cu): C % 32 == 0, stride 1, 'same' padding, MSCA's kernel sizes, ragged tiles
    (2, 32, 19, 37, 1, 5, 5, 1, 1, 2, 2),
    (1, 64, 33, 18, 1, 21, 21, 1, 1, 10, 10),
    (2, 32, 7, 7, 1, 21, 21, 1, 1, 10, 10),
    (2, 32, 20, 35, 1, 1, 7, 1, 1, 0, 3),
    (2, 32, 35, 20, 1, 7, 1, 1, 1, 3, 0),
    (1, 96, 16, 40, 1, 1, 11, 1, 1, 0, 5),
    (1, 32, 40, 16, 1, 11, 1, 1, 1, 5, 0),
    (2, 32, 17, 50, 1, 1, 21, 1, 1, 0, 10),
    (2, 160, 32, 32, 1, 21, 1, 1, 1, 10, 0),
    (2, 32, 18, 18, 1, 3, 3, 1, 1, 1, 1),
    # TMA-pipelined strips (dwstrip_tma.cu, bf16): several tiles per axis, ragged edges, every strip length it is built for
    (3, 64, 70, 75, 1, 1, 7, 1, 1, 0, 3),
    (3, 64, 70, 75, 1, 21, 1, 1, 1, 10, 0),
    (2, 32, 33, 65, 1, 1, 3, 1, 1, 0, 1),
    (2, 32, 65, 33, 1, 5, 1, 1, 1, 2, 0),
    (2, 96, 40, 100, 1, 1, 11, 1, 1, 0, 5),
    (1, 32, 100, 40, 1, 11, 1, 1, 1, 5, 0),
    (2, 32, 5, 19, 1, 1, 21, 1, 1, 0, 10),
    (2, 32, 19, 5, 1, 7, 1, 1, 1, 3, 0),
]


@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_dwconv_vs_torch(B, case, dt):
    Bn, C, H, W, mult, kh, kw, sh, sw, ph, pw = case
    torch.manual_seed(2)
    x = torch.randn(Bn, C, H, W).to(dt)
    w = (torch.randn(C * mult, 1, kh, kw) / (kh * kw) ** 0.5).to(dt)
    b = torch.randn(C * mult).to(dt)
    ref = F.conv2d(x, w, b, (sh, sw), (ph, pw), groups=C)
    wpk = B.pack_dw(w.cuda().float(), dt)
    y = B.dwconv(B.to_nhwc(x.cuda()), wpk, b.float().cuda(), mult, kh, kw, (sh, sw), (ph, pw))
    tol = TOL32 if dt == torch.float32 else TOL16
    assert O.compare4d(ref.float(), nchw(y)) < tol
    # fused running-sum + FixPaddingBias residues
    add = torch.randn_like(ref)
    res = torch.randn(2, C * mult, 10)
    ref2 = O.fix_padding_bias(ref.clone(), res.to(dt), 10) + add
    y2 = B.dwconv(B.to_nhwc(x.cuda()), wpk, b.float().cuda(), mult, kh, kw, (sh, sw), (ph, pw),
                  res=res.to(dt).float().cuda(), add=B.to_nhwc(add.cuda()).contiguous())
    assert O.compare4d(ref2.float(), nchw(y2)) < (tol if dt == torch.float32 else 3e-3)


@pytest.mark.parametrize("k,along_w", [(3, True), (7, True), (7, False), (11, False), (21, True), (21, False)])
def test_tma_strip_kernel_is_bit_identical_to_the_register_window_kernel(B, k, along_w, monkeypatch):
    """The TMA-fed strip kernel (csrc/dwstrip_tma.cu) accumulates the taps in the same order as dw_strip_kernel, so the two agree
    bit for bit; CNA_DW_NO_TMA selects the old one.  Shape with ragged tiles on both axes and three channel chunks."""
    torch.manual_seed(k)
    kh, kw = (1, k) if along_w else (k, 1)
    x = torch.randn(3, 45, 77, 96, device="cuda").bfloat16()
    w = torch.randn(96, 1, kh, kw, device="cuda") / k ** 0.5
    wpk = B.pack_dw(w, torch.bfloat16)
    b = torch.randn(96, device="cuda")
    y_new = B.dwconv(x, wpk, b, 1, kh, kw, (1, 1), (kh // 2, kw // 2))
    monkeypatch.setenv("CNA_DW_NO_TMA", "1")
    y_old = B.dwconv(x, wpk, b, 1, kh, kw, (1, 1), (kh // 2, kw // 2))
    monkeypatch.delenv("CNA_DW_NO_TMA")
    assert torch.equal(y_new, y_old)


@pytest.mark.parametrize("shape", [(2, 64, 37, 45), (1, 256, 16, 16), (3, 32, 70, 9)])
@pytest.mark.parametrize("act", [None, "gelu"])
def test_dwconv3x3_gelu_tma_vs_torch(B, shape, act):
    """FFN.dconv + FFN.act of an MSCAN block (reference models/mscan.py:38-45, 58-59) in one pass; bf16 vs the fp32 math on the
    same bf16 inputs, tolerance = one bf16 rounding of the output."""
    Bn, C, H, W = shape
    torch.manual_seed(7)
    x = torch.randn(Bn, C, H, W).bfloat16()
    w = (torch.randn(C, 1, 3, 3) / 3).bfloat16()
    b = torch.randn(C).bfloat16()
    ref = F.conv2d(x.float(), w.float(), b.float(), 1, 1, groups=C)
    if act == "gelu":
        ref = F.gelu(ref)
    y = B.dwconv3x3_act(B.to_nhwc(x.cuda()).contiguous(), B.pack_dw(w.cuda().float(), torch.bfloat16), b.float().cuda(), act)
    assert O.compare4d(ref, nchw(y)) < 3e-3
    assert O.compare4d(ref.bfloat16().float(), nchw(y)) < TOL16


@pytest.mark.parametrize("H", [24, 7, 20])
def test_fix_padding_bias_inplace(B, H):
    torch.manual_seed(3)
    x = torch.randn(2, 6, H, 5)
    res = torch.randn(2, 6, 10)
    ref = O.fix_padding_bias(x.clone(), res, 10)
    xn = B.to_nhwc(x.cuda()).contiguous()
    B.fix_padding_bias_(xn, res.cuda())
    assert torch.allclose(nchw(xn), ref, atol=1e-6)


def test_maxpool_and_add(B):
    torch.manual_seed(4)
    for dt, C in [(torch.float32, 5), (torch.bfloat16, 16), (torch.bfloat16, 6)]:
        x = torch.randn(2, C, 8, 6).to(dt)
        y = B.maxpool2x2(B.to_nhwc(x.cuda()).contiguous())
        assert torch.equal(nchw(y), F.max_pool2d(x.float(), 2))
        a, b = torch.randn(3, 4, 5, C).to(dt), torch.randn(3, 4, 5, C).to(dt)
        assert torch.equal(B.add(a.cuda(), b.cuda()).cpu(), a + b)


def test_errors_are_loud(B):
    from approx.layers import MSCA
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MSCA(8, 5, (7, 11, 21))(torch.randn(1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="cna_b200"):
        B._check(B.lib().cna_conv_taps_forward(0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0, 0))


@pytest.mark.parametrize("case", [(2, 32, 9, 13, 32, True), (1, 160, 16, 16, 160, True), (2, 64, 7, 5, 48, False),
                                  (1, 256, 12, 11, 256, True), (3, 96, 5, 5, 16, False)])
def test_conv1x1_tcgen05(B, case):
    """tcgen05 1x1 conv (+bias, +gate / ReLU) against the bf16 torch reference of channel_mix / d_conv."""
    Bn, C, H, W, N, gated = case
    torch.manual_seed(11)
    x = torch.randn(Bn, C, H, W).bfloat16()
    w = (torch.randn(N, C, 1, 1) / C ** 0.5).bfloat16()
    b = torch.randn(N).bfloat16()
    assert B.conv1x1_tc_supported(C, N)
    xn = B.to_nhwc(x.cuda()).contiguous()
    w_nc = B.conv1x1_pack(w.cuda().float())
    if gated:
        g = torch.randn(Bn, N, H, W).bfloat16()
        ref = g * F.conv2d(x, w, b)
        y = B.conv1x1_tc(xn, w_nc, b.float().cuda(), gate=B.to_nhwc(g.cuda()).contiguous())
    else:
        ref = torch.relu(F.conv2d(x, w, b))
        y = B.conv1x1_tc(xn, w_nc, b.float().cuda(), relu=True)
    assert O.compare4d(ref.float(), nchw(y)) < TOL16
