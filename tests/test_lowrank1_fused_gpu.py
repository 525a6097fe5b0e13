"""Parity of the fused tcgen05 scheme-1 kernel (csrc/lowrank1_tc.cu): against the oracle run in bf16 on identical bf16
inputs / weights (tolerance compare4d <= 1e-3, BASELINE.json) and against this library's unfused CUDA path."""
import pytest
import torch

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # B, C, N, d, M, H, W, separable
    (2, 64, 192, 5, 8, 27, 27, False),    # AlexNet conv2 (configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py)
    (2, 192, 384, 3, 8, 13, 13, False),   # AlexNet conv3: N = 384 -> two N-tiles on one A slab, tiles straddle images
    (3, 32, 48, 3, 8, 13, 13, True),      # separable form, tiles straddle images (R = 9 rows of H = 13)
    (2, 16, 64, 3, 4, 9, 20, False),      # M = 4: two input channels behind one 8-wide unit
    (1, 8, 16, 5, 8, 7, 100, True),       # W = 100: one image row per tile
    (2, 24, 288, 3, 8, 6, 11, True),      # N = 288 = 2 x 144
    (2, 16, 32, 5, 4, 30, 17, True),      # M = 4, separable 5 + 5
    (1, 8, 16, 3, 8, 3, 3, False),        # tiny image: R = H = 3, mostly empty tile
    (2, 24, 32, 3, 8, 90, 100, False),    # 180 one-row tiles on 148 SMs, 3 K-chunks: the two producer groups swap parity per tile
    (2, 24, 32, 3, 8, 90, 100, True),
    (2, 16, 64, 3, 8, 20, 16, False),     # W % 4 == 0: row-run mapping at no padding cost (3x3, d x d form)
    (3, 8, 32, 5, 8, 9, 30, True),        # row runs, separable 5 + 5, 4-row tiles straddling the 9-row images
]


def build(C, N, d, M, sep, seed):
    from approx.layers import LowRankExpConvV1
    torch.manual_seed(seed)
    m = LowRankExpConvV1(C, N, d, 1, d // 2, num_base=M, decomp=sep)
    return m.bfloat16()


def oracle(m, x, sep):
    if sep:
        return O.lowrank1_forward_decomp(x, m.s_conv.v_conv.weight, m.s_conv.h_conv.weight, m.d_conv.weight, m.d_conv.bias,
                                         (1, 1), m.padding)
    return O.lowrank1_forward(x, m.s_conv.weight, m.s_conv.bias, m.d_conv.weight, m.d_conv.bias, (1, 1), m.padding)


@pytest.mark.parametrize("case", CASES)
def test_fused_matches_bf16_oracle(case):
    from approx import backend as B
    Bn, C, N, d, M, H, W, sep = case
    m = build(C, N, d, M, sep, 3)
    x = torch.randn(Bn, C, H, W, generator=torch.Generator().manual_seed(9)).bfloat16()
    with torch.no_grad():
        ref = oracle(m, x, sep).float()
    assert B.lowrank1_fused_supported(Bn, H, W, C, N, M, d, sep)
    mg = m.cuda().eval()
    xn = B.to_nhwc(x.cuda())
    with torch.no_grad():
        before = B.launch_count()
        y = mg.forward_nhwc(xn)
        assert B.launch_count() - before <= 2          # (d_conv weight pack) + ONE fused kernel
        y2 = mg.forward_nhwc(xn, fused=False)
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float().cpu()
    assert got.shape == ref.shape
    err = O.compare4d(ref, got)
    assert err < 1e-3, f"rel err vs bf16 oracle {err}"
    err2 = O.compare4d(y2.permute(0, 3, 1, 2).float().cpu(), got)
    assert err2 < 1e-3, f"rel err vs unfused CUDA path {err2}"


def test_module_forward_takes_fused_path_after_decomp():
    """LowRankExpConvV1 -> decomp() -> forward on NCHW input: the module switches to the separable fused kernel."""
    import copy
    m = build(64, 192, 5, 8, False, 5).float().cuda().eval()
    x = torch.randn(4, 64, 27, 27, generator=torch.Generator().manual_seed(1)).bfloat16()
    with torch.no_grad():
        y0 = m.bfloat16()(x.cuda())
        m = m.float()
        m.decomp()                       # batched rank-1 SVD on the device
        m = m.bfloat16()
        ref = oracle(copy.deepcopy(m).cpu(), x, True).float()
        y1 = m(x.cuda())
    assert O.compare4d(ref, y1.float().cpu()) < 1e-3
    # rank-1 split of random 5x5 bases is lossy: the two forms differ, the kernel must not have reused the d x d filters
    assert O.compare4d(y0.float().cpu(), y1.float().cpu()) > 1e-2


def test_full_size_agrees_with_unfused_path():
    """AlexNet conv4 at the reference's profiling batch (B = 64): the tcgen05 kernel against the FFMA kernels (two independent
    implementations of the same rounding points) — size-independent check at the full shape."""
    from approx import backend as B
    m = build(384, 256, 3, 8, False, 7).cuda().eval()
    x = torch.randn(64, 13, 13, 384, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)).bfloat16()
    with torch.no_grad():
        y = m.forward_nhwc(x)
        y2 = m.forward_nhwc(x, fused=False)
    assert torch.isfinite(y.float()).all()
    num = (y.float() - y2.float()).norm() / y2.float().norm()
    assert num < 1e-3, float(num)


def test_random_shapes_agree_with_unfused_path():
    """Seeded sweep over ragged shapes (both slot mappings, both forms, 1-3 tiles per SM, 1..6 K-chunks): the tcgen05 kernel
    against the FFMA kernels on the same bf16 inputs."""
    import random
    from approx import backend as B
    rng = random.Random(1234)
    done = 0
    while done < 16:
        M = rng.choice([4, 8])
        d = rng.choice([3, 5])
        C = (64 // M) * rng.randint(1, 6)
        if C % 8:
            continue
        N = 16 * rng.randint(1, 24)
        if N > 256 and N % 32:
            continue
        H, W = rng.randint(d, 40), rng.randint(d, 60)
        Bn = rng.randint(1, 5)
        sep = rng.random() < 0.5
        if not B.lowrank1_fused_supported(Bn, H, W, C, N, M, d, sep):
            continue
        m = build(C, N, d, M, sep, 100 + done).cuda().eval()
        x = torch.randn(Bn, H, W, C, device="cuda", generator=torch.Generator(device="cuda").manual_seed(done)).bfloat16()
        with torch.no_grad():
            y = m.forward_nhwc(x)
            y2 = m.forward_nhwc(x, fused=False)
        err = float((y.float() - y2.float()).norm() / y2.float().norm())
        assert err < 1e-3, (Bn, C, N, d, M, H, W, sep, err)
        done += 1


@pytest.mark.parametrize("sep", [False, True])
@pytest.mark.parametrize("M", [2, 3, 5, 6, 7, 9, 10, 11, 16])
def test_num_base_outside_4_and_8_runs_fused_by_padding(M, sep):
    """The reference's own sweep (scripts/low_rank_exp_spr.py:35-63) walks num_base = 2..11: every M <= 16 stays on the fused kernel
    (padded with zero bases to the 2-, 4-, 8- or 16-basis instance) and still matches the bf16 oracle."""
    from approx import backend as B
    C, N, d, H = 64, 96, 3, 13
    m = build(C, N, d, M, sep, 11 + M)
    x = torch.randn(4, C, H, H, generator=torch.Generator().manual_seed(2)).bfloat16()
    with torch.no_grad():
        ref = oracle(m, x, sep).float()
    mg = m.cuda().eval()
    xn = B.to_nhwc(x.cuda())
    with torch.no_grad():
        mg.forward_nhwc(xn)                          # first call packs
        before = B.launch_count()
        y = mg.forward_nhwc(xn)
        assert B.launch_count() - before == 1, "expected exactly the fused kernel"
    assert O.compare4d(ref, y.permute(0, 3, 1, 2).float().cpu()) < 1e-3


@pytest.mark.parametrize("Bn,C,N,d,M,H,W,sep", [
    (3, 8, 32, 3, 16, 13, 13, False),        # one channel octet = two chunks, tiles straddle images
    (2, 72, 64, 3, 12, 27, 27, True),        # 18 chunks, 12 bases padded to 16, decomposed form
    (2, 24, 192, 5, 9, 27, 27, False),       # 5x5 taps (AlexNet conv2 geometry), 9 -> 16
    (2, 40, 48, 5, 16, 9, 20, True),         # 5x5 separable, W != H, ragged last tile
    (5, 192, 384, 3, 11, 13, 13, False),     # AlexNet conv3 at num_base 11: two N-tiles share every A slab
    (2, 64, 192, 5, 2, 27, 27, False),       # 2-basis instance: 32 input channels per chunk, AlexNet conv2 geometry
    (3, 96, 64, 5, 2, 9, 20, True),          # 2 bases, 3 chunks, separable 5 + 5, W != H
    (4, 256, 256, 3, 2, 7, 7, False),        # the sweep's conv5 point at num_base 2 (7 x 7 images: 49 of 128 tile rows)
])
def test_two_and_sixteen_basis_instances_vs_oracle(Bn, C, N, d, M, H, W, sep):
    """num_base 9..16 run on the 16-basis instance (4 input channels per 64-wide chunk; the staged x box is the enclosing channel
    octet, of which a chunk reads one half), num_base 2 on the 2-basis instance (32 input channels per chunk) when C % 32 == 0:
    one launch, bf16 oracle within 1e-3 (compare4d, reference tests/test_low_rank_exp.py:5-14)."""
    from approx import backend as B
    m = build(C, N, d, M, sep, 100 + M + C)
    x = torch.randn(Bn, C, H, W, generator=torch.Generator().manual_seed(5)).bfloat16()
    with torch.no_grad():
        ref = oracle(m, x, sep).float()
    mg = m.cuda().eval()
    xn = B.to_nhwc(x.cuda())
    with torch.no_grad():
        mg.forward_nhwc(xn)
        before = B.launch_count()
        y = mg.forward_nhwc(xn)
        assert B.launch_count() - before == 1, "expected exactly the fused kernel"
    err = O.compare4d(ref, y.permute(0, 3, 1, 2).float().cpu())
    assert err < 1e-3, (Bn, C, N, d, M, H, W, sep, err)


def test_plan_record_follows_weight_updates_decomp_and_invalidate():
    """The fused path keeps the last call's decision and packed operands in one record (the small layers are host-bound at the
    reference's profiling batches).  It must notice `.weight.data = ...` assignments (the reference's write-back,
    core/low_rank_exp.py:181-184), `decomp()` replacing `s_conv` (layers/low_rank_conv.py:58-66), `_ops.invalidate` after an in-place
    edit through `.data`, a different input shape, and the train()/no-grad switch — each time against the bf16 oracle."""
    from approx import backend as B
    from approx.layers import _ops
    C, N, d, M, H = 64, 96, 3, 4, 13
    m = build(C, N, d, M, False, 41)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, C, H, H, generator=g).bfloat16()
    mg = m.cuda().eval()
    xg = x.cuda().contiguous(memory_format=torch.channels_last)

    def check(xc, xd, sep):
        twin = build(C, N, d, M, sep, 0)                   # CPU module holding mg's current parameters, for the oracle
        twin.load_state_dict({k: v.cpu() for k, v in mg.state_dict().items()})
        with torch.no_grad():
            ref = oracle(twin, xc, sep).float()
            before = B.launch_count()
            y = mg(xd)
            assert B.launch_count() - before <= 3          # (re-pack after a change) + the fused kernel
            y2 = mg(xd)                                    # second call: the record is reused
        assert torch.equal(y, y2)
        err = O.compare4d(ref, y.float().cpu())
        assert err < 1e-3, err

    check(x, xg, False)
    mg.d_conv.weight.data = (torch.randn(N, C * M, 1, 1, generator=g) * 0.05).bfloat16().cuda()      # write-back style assignment
    check(x, xg, False)
    mg.s_conv.weight.data = (torch.randn(C * M, 1, d, d, generator=g) * 0.3).bfloat16().cuda()
    check(x, xg, False)
    mg.d_conv.bias.data.mul_(3.0)                       # in place through .data: no version bump, the documented case for invalidate()
    _ops.invalidate(mg)
    check(x, xg, False)
    x5 = torch.randn(5, C, H + 1, H, generator=g).bfloat16()                                          # another input shape
    check(x5, x5.cuda().contiguous(memory_format=torch.channels_last), False)
    check(x, xg, False)
    mg.decomp()                                         # s_conv becomes a SeparableConv
    check(x, xg, True)
    mg.load_state_dict({k: (v * 0.5 if k == "d_conv.weight" else v) for k, v in mg.state_dict().items()})
    check(x, xg, True)
