"""Drop-in modules and apps on the GPU against golden vectors produced by the reference's own modules
(tests/golden/make_golden.py).  Weights are loaded through `load_state_dict` with the reference's keys,
which also proves state_dict compatibility.  fp32 tolerance: compare4d < 1e-5."""
import io
import contextlib

import numpy as np
import pytest
import torch
from torch import nn

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def T(a):
    return torch.from_numpy(np.asarray(a)).float()


def sd_from(g, prefix):
    return {k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix) and "." in k[len(prefix):]}


def test_lowrank2_shipped_golden(golden):
    from approx.layers import LowRankExpConvV2
    g = golden("lowrank2_shipped")
    for tag in "ab":
        C, K, kh, kw, sh, sw, ph, pw = (int(v) for v in g[f"{tag}_cfg"])
        m = LowRankExpConvV2(C, K, (kh, kw), (sh, sw), (ph, pw))
        m.load_state_dict(sd_from(g, f"{tag}_"), strict=True)
        y = m.cuda()(T(g[f"{tag}_x"]).cuda())
        assert O.compare4d(T(g[f"{tag}_y"]), y.cpu()) < TOL


def test_lowrank1_golden_forward_and_decomp(golden):
    from approx.layers import LowRankExpConvV1
    g = golden("lowrank1")
    for tag in "ab":
        C, N, k, s, p, M = (int(v) for v in g[f"{tag}_cfg"])
        m = LowRankExpConvV1(C, N, k, s, p, num_base=M)
        sd = {kk: vv for kk, vv in sd_from(g, f"{tag}_").items() if not kk.startswith("dec_")}
        m.load_state_dict(sd, strict=True)
        m = m.cuda()
        x = T(g[f"{tag}_x"]).cuda()
        assert O.compare4d(T(g[f"{tag}_y"]), m(x).cpu()) < TOL
        m.decomp()          # batched rank-1 SVD on the device
        m.decomp()          # idempotent
        assert sorted(m.state_dict()) == sorted(k[len(f"{tag}_dec_"):] for k in g if k.startswith(f"{tag}_dec_"))
        assert O.compare4d(T(g[f"{tag}_y_decomp"]), m(x).cpu()) < TOL


def test_lowrank_exp_v1_app_golden(golden):
    from approx.core import LowRankExpV1
    g = golden("lowrank1_app")
    src = nn.Conv2d(6, 10, 3, 1, 1)
    src.weight.data = T(g["src_weight"]); src.bias.data = T(g["src_bias"])
    app = LowRankExpV1(num_bases=(4,), max_iter=0, lmda_length=1, min_lmda=0.001, max_lmda=0.001,
                       init_method="svd", do_decomp=True)
    sub = app.initialize(src)              # weights still on the CPU, as in Runner.run
    app.optimize(sub)                      # SVD fit runs on the device, results land back on the CPU
    assert sub.new.s_conv.weight.device.type == "cpu"
    C, M = 6, 4
    recon = lambda sw, dw: dw.reshape(-1, M) @ sw.reshape(C, M, 9)[0]
    assert torch.allclose(recon(sub.new.s_conv.weight.data, sub.new.d_conv.weight.data),
                          recon(T(g["fit_s_conv.weight"]), T(g["fit_d_conv.weight"])), atol=2e-5)
    tgt = app.postprocess(sub).cuda()
    assert sorted(tgt.state_dict()) == sorted(k[len("post_"):] for k in g if k.startswith("post_"))
    y = tgt(T(g["x"]).cuda())
    assert O.compare4d(T(g["y_post"]), y.cpu()) < 2e-5


def _msca_from_golden(g):
    from approx.layers import MSCA
    m = MSCA(8, 5, (7, 11, 21))
    m.load_state_dict(sd_from(g, "src_"), strict=True)
    return m


def test_msca_original_golden(golden):
    g = golden("msca")
    m = _msca_from_golden(g).cuda()
    for t in ("big", "small"):
        assert O.compare4d(T(g[f"y_src_{t}"]), m(T(g[f"x_{t}"]).cuda()).cpu()) < TOL


def test_msca_equivalent_kernel_golden(golden):
    from approx.core import MscaRep
    g = golden("msca")
    m = _msca_from_golden(g).cuda()
    w, b, r = MscaRep.get_equivalent_kernel(m.sd_convs)
    assert torch.allclose(w.cpu(), T(g["eq_weight"]), atol=1e-6)
    assert torch.allclose(b.cpu(), T(g["eq_bias"]), atol=1e-6)
    assert torch.allclose(r.cpu(), T(g["eq_res"]), atol=1e-6)


@pytest.mark.parametrize("decomp,fix", [(0, True), (0, False), (1, True), (2, True), (4, True), (3, False)])
def test_msca_rep_variants_golden(golden, decomp, fix):
    from approx.core import MscaRep
    g = golden("msca")
    tag = f"d{decomp}{'f' if fix else 'n'}"
    src = _msca_from_golden(g)                       # CPU weights, as in Runner.run
    app = MscaRep(decomp=decomp, fix=fix)
    sub = app.initialize(src)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        app.optimize(sub)
    if decomp > 0:
        assert abs(float(buf.getvalue().strip().rstrip(",")) - float(g[f"{tag}_pce"])) < 1e-4
    tgt = app.postprocess(sub)
    want_keys = sorted(k[len(tag) + 1:] for k in g if k.startswith(tag + "_") and "." in k)
    assert sorted(tgt.state_dict()) == want_keys
    tgt = tgt.cuda()
    for t in ("big", "small"):
        y = tgt(T(g[f"x_{t}"]).cuda())
        assert O.compare4d(T(g[f"{tag}_y_{t}"]), y.cpu()) < 3e-5, (tag, t)
    if (decomp, fix) in ((0, True), (4, True)):       # exact re-parameterisations reproduce the source block
        for t in ("big", "small"):
            assert O.compare4d(T(g[f"y_src_{t}"]), tgt(T(g[f"x_{t}"]).cuda()).cpu()) < 3e-5


def test_msca_profile_app_and_substitution(golden):
    from approx.core import MscaProfile
    from approx.layers import MSCAProfile, Substitution
    g = golden("msca")
    src = _msca_from_golden(g)
    sub = MscaProfile(deploy=False).initialize(src)
    assert isinstance(sub, Substitution) and isinstance(sub.new, MSCAProfile)
    sub = sub.cuda()
    x = T(g["x_big"]).cuda()
    y_old = sub(x)
    sub.switch_new(remove_old=True)
    assert not hasattr(sub, "old")
    assert O.compare4d(y_old.cpu(), sub(x).cpu()) < TOL
    assert O.compare4d(T(g["y_src_big"]), sub(x).cpu()) < TOL


def test_svd_small_reconstruction():
    from approx import backend as B
    torch.manual_seed(0)
    for shape in [(37, 1, 3, 3), (5, 21, 21), (3, 8, 5), (2, 5, 8), (4, 25, 25)]:
        A = torch.randn(*shape)
        U, S, Vh = B.svd_small(A.cuda())
        rec = (U * S.unsqueeze(-2)) @ Vh
        assert torch.allclose(rec.cpu(), A, atol=2e-5)
        assert torch.allclose(S.cpu(), torch.linalg.svdvals(A), atol=2e-5)
        assert bool((S[..., :-1] >= S[..., 1:] - 1e-6).all())


def test_bf16_msca_against_bf16_oracle(golden):
    """bf16 mode: the kernels round where the reference's bf16 modules materialise tensors."""
    g = golden("msca")
    m = _msca_from_golden(g)
    x = T(g["x_big"]).bfloat16()
    mb = _msca_from_golden(g).bfloat16()
    sd = {k: v for k, v in mb.state_dict().items()}
    br = [tuple(sd[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    ref = O.msca_forward(x, sd["conv0.weight"], sd["conv0.bias"], lambda a: O.parallel_forward(a, br, True),
                         sd["channel_mix.weight"], sd["channel_mix.bias"])
    y = mb.cuda().eval()(x.cuda())
    assert O.compare4d(ref.float(), y.float().cpu()) < 1e-3


def _oracle_msca(m, x, dt):
    """Oracle forward of an (original) MSCA module `m` in dtype dt on CPU."""
    sd = {k: v.detach().cpu().to(dt) for k, v in m.state_dict().items()}
    br = [tuple(sd[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    return O.msca_forward(x.to(dt), sd["conv0.weight"], sd["conv0.bias"], lambda a: O.parallel_forward(a, br, True),
                          sd["channel_mix.weight"], sd["channel_mix.bias"])


@pytest.mark.parametrize("C,H,W", [(32, 40, 50), (16, 7, 9), (64, 33, 32), (48, 21, 70)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_msca_fused_sd_kernel_vs_oracle(C, H, W, dt, monkeypatch):
    """The fused marching kernel (conv0 -> 3 strip cascades + identity) on ragged sizes, fp32 and bf16."""
    from approx import backend as B
    from approx.layers import MSCA
    monkeypatch.setenv("CNA_MSCA_FUSED", "1")        # force the one-pass kernel
    torch.manual_seed(C + H)
    m = MSCA(C, 5, (7, 11, 21))
    x = torch.randn(2, C, H, W)
    ref = _oracle_msca(m, x, dt)
    md = m.cuda().to(dt).eval()          # inference, as the reference's hooks run it (model.eval()); train() mode takes the autograd path
    n0 = B.launch_count()
    y = md(x.cuda().to(dt))
    assert B.launch_count() - n0 <= 12      # packing + layout + 2 compute kernels (fused sd, mix+gate)
    assert O.compare4d(ref.float(), y.float().cpu()) < (TOL if dt == torch.float32 else 1e-3)
    n0 = B.launch_count()
    md(x.cuda().to(dt))
    assert B.launch_count() - n0 <= 3       # cached weights: layout + fused sd + mix/gate


@pytest.mark.parametrize("C,H,W", [(32, 40, 50), (64, 33, 32), (160, 16, 17), (96, 7, 9)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_msca_tiled_chain_vs_oracle(C, H, W, dt, monkeypatch):
    """cna_msca_chain_forward (register-tiled 5x5 + strip kernels enqueued by one C call; the default for single-cascade forms)."""
    from approx import backend as B
    from approx.layers import MSCA
    monkeypatch.setenv("CNA_MSCA_FUSED", "chain")
    assert B.msca_chain_supported(C, 5, (7, 11, 21), dt)
    torch.manual_seed(C + H + 1)
    m = MSCA(C, 5, (7, 11, 21))
    x = torch.randn(2, C, H, W)
    ref = _oracle_msca(m, x, dt)
    y = m.cuda().to(dt).eval()(x.cuda().to(dt))
    assert O.compare4d(ref.float(), y.float().cpu()) < (TOL if dt == torch.float32 else 1e-3)


@pytest.mark.parametrize("decomp,fix", [(1, True), (2, True), (3, False), (0, True)])
def test_msca_rep_fused_matches_unfused(decomp, fix):
    """Re-parameterised forms through the fused kernel agree with the reference algebra on a 32-channel block."""
    from approx.core import MscaRep
    from approx.layers import MSCA
    torch.manual_seed(5)
    src = MSCA(32, 5, (7, 11, 21))
    x = torch.randn(2, 32, 26, 45)
    app = MscaRep(decomp=decomp, fix=fix)
    sub = app.initialize(src)
    with contextlib.redirect_stdout(io.StringIO()):
        app.optimize(sub)
    tgt = app.postprocess(sub)
    # oracle: same re-parameterisation applied with the CPU oracle functions
    sd = {k: v.detach().cpu() for k, v in src.state_dict().items()}
    br = [tuple(sd[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"))
          for i in range(3)]
    w, b, r = O.msca_equivalent_kernel(br, 32)
    import torch.nn.functional as F
    if decomp == 0:
        sdf = lambda a: O.fix_padding_bias(F.conv2d(a, w, b, padding=10, groups=32), r, 10)
    else:
        fac, _ = O.msca_rep_factors(w, decomp)
        def sdf(a):
            brs = [(c1, None, c2, b if j == decomp - 1 else None) for j, (c1, c2) in enumerate(fac)]
            yy = O.parallel_forward(a, brs, False)
            return O.fix_padding_bias(yy, r, 10) if fix else yy
    ref = O.msca_forward(x, sd["conv0.weight"], sd["conv0.bias"], sdf, sd["channel_mix.weight"], sd["channel_mix.bias"])
    y = tgt.cuda().eval()(x.cuda())
    assert O.compare4d(ref, y.cpu()) < 5e-5


def _torch_msca_forward(m, x):
    """Stock-torch restatement of MSCA.forward for any sd_convs form (the reference's modules, layers/msca.py:25-26,
    depth_seperable_conv.py:12-16,45-47,57-61) — the autograd reference for the training path."""
    import torch.nn.functional as F
    from approx.layers.depth_seperable_conv import CascadeConv, FixPaddingBias, ParallelConv
    C = m.num_channel
    conv = lambda c, t: F.conv2d(t, c.weight, c.bias, 1, c.padding, 1, C)
    a = conv(m.conv0, x)
    sd, fix = m.sd_convs, None
    if isinstance(sd, nn.Sequential):
        sd, fix = sd[0], sd[1]
    if isinstance(sd, ParallelConv):
        s = sum([conv(br.conv2, conv(br.conv1, a)) for br in sd.cascades()] + ([a] if sd.has_identity else []))
    elif isinstance(sd, CascadeConv):
        s = conv(sd.conv2, conv(sd.conv1, a))
    else:
        s = conv(sd, a)
    if fix is not None:
        p2 = min(s.shape[2], fix.p)
        top = torch.zeros_like(s); bot = torch.zeros_like(s)
        top[:, :, :p2, :] = fix.res[0][None, :, :p2, None]
        bot[:, :, -p2:, :] = fix.res[1][None, :, -p2:, None]
        s = s + top + bot
    return x * F.conv2d(s, m.channel_mix.weight, m.channel_mix.bias)


@pytest.mark.parametrize("form", ["original", "d1_fix", "d0_fix", "d3"])
def test_msca_training_path_gradients_vs_autograd(form):
    """model.train(): forward and every parameter / input gradient of the op-by-op CUDA path against torch autograd (fp32)."""
    import copy
    from approx.core import MscaRep
    from approx.layers import MSCA
    torch.manual_seed(13)
    C, H, W = 32, 14, 19
    m = MSCA(C, 5, (7, 11, 21))
    if form != "original":
        d, f = {"d1_fix": (1, True), "d0_fix": (0, True), "d3": (3, False)}[form]
        app = MscaRep(decomp=d, fix=f)
        sub = app.initialize(m)
        with contextlib.redirect_stdout(io.StringIO()):
            app.optimize(sub)
        m = app.postprocess(sub)
    ref = copy.deepcopy(m)                              # torch reference stays on the CPU: exact fp32 (cuDNN would use TF32)
    m = m.cuda().train()
    x = torch.randn(2, C, H, W)
    gy = torch.randn(2, C, H, W)
    x1 = x.clone().cuda().requires_grad_(True)
    y1 = m(x1)
    y1.backward(gy.cuda())
    x2 = x.clone().requires_grad_(True)
    y2 = _torch_msca_forward(ref, x2)
    y2.backward(gy)
    assert O.compare4d(y2.detach(), y1.detach().cpu()) < 1e-5
    assert (x1.grad.cpu() - x2.grad).norm() / x2.grad.norm() < 1e-4
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert n1 == n2
        assert p1.grad is not None, n1
        denom = p2.grad.norm().clamp_min(1e-6)
        assert (p1.grad.cpu() - p2.grad).norm() / denom < 2e-4, (n1, ((p1.grad.cpu() - p2.grad).norm() / denom).item())


def test_msca_rep_exactness_at_full_mscan_t_size():
    """BASELINE config 4 at full size (MSCAN-T stage 1: 32 channels, 128 x 128, batch 32): the exact re-parameterisations
    (d0 + fix: merged 21x21 kernel; d4 + fix: four SVD cascades) reproduce the original block — a size-independent property that
    needs no CPU oracle — in fp32, and the bf16 forms agree with the fp32 result to bf16 accuracy."""
    from approx.core import MscaRep
    from approx.layers import MSCA
    torch.manual_seed(21)
    src = MSCA(32, 5, (7, 11, 21))
    x = torch.randn(32, 32, 128, 128, device="cuda")
    with torch.no_grad():
        y0 = src.cuda().eval()(x)
        for d in (0, 4):
            app = MscaRep(decomp=d, fix=True)
            sub = app.initialize(src.cpu())
            with contextlib.redirect_stdout(io.StringIO()):
                app.optimize(sub)
            tgt = app.postprocess(sub).cuda().eval()
            assert O.compare4d(y0, tgt(x)) < 2e-5, d
            yb = tgt.bfloat16()(x.bfloat16())
            assert O.compare4d(y0, yb.float()) < 1e-2, d
        src.cuda()


@pytest.mark.parametrize("decomp", [False, True])
def test_lowrank1_training_path_gradients_vs_autograd(decomp):
    """LowRankExpConvV1 in model.train(): forward, input and parameter gradients of the CUDA path (grouped conv with channel
    multiplier or its separable form, then the 1x1 mix) against torch autograd through stock convs on the CPU (exact fp32)."""
    import copy
    import torch.nn.functional as F
    from approx.layers import LowRankExpConvV1
    torch.manual_seed(17)
    C, N, M, d = 12, 20, 4, 3
    m = LowRankExpConvV1(C, N, d, 1, 1, num_base=M, decomp=decomp)
    ref = copy.deepcopy(m)
    m = m.cuda().train()
    x = torch.randn(2, C, 11, 9)
    gy = torch.randn(2, N, 11, 9)
    x1 = x.clone().cuda().requires_grad_(True)
    y1 = m(x1)
    y1.backward(gy.cuda())
    x2 = x.clone().requires_grad_(True)
    if decomp:
        t = F.conv2d(x2, ref.s_conv.v_conv.weight, None, 1, (1, 0), 1, C)
        t = F.conv2d(t, ref.s_conv.h_conv.weight, None, 1, (0, 1), 1, C * M)
    else:
        t = F.conv2d(x2, ref.s_conv.weight, ref.s_conv.bias, 1, 1, 1, C)
    y2 = F.conv2d(t, ref.d_conv.weight, ref.d_conv.bias)
    y2.backward(gy)
    assert O.compare4d(y2.detach(), y1.detach().cpu()) < 1e-5
    assert (x1.grad.cpu() - x2.grad).norm() / x2.grad.norm() < 1e-4
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), ref.named_parameters()):
        assert n1 == n2 and p1.grad is not None, n1
        assert (p1.grad.cpu() - p2.grad).norm() / p2.grad.norm().clamp_min(1e-6) < 2e-4, n1


@pytest.mark.parametrize("d,M", [(3, 4), (5, 8)])
def test_alternating_nuclear_norm_fit_vs_fp64_restatement(d, M):
    """cna_lowrank1_alt_fit (the device replacement of the reference's cvxpy loop, core/low_rank_exp.py:160-176): the objective
    never increases along the alternation, both block updates reach the optimum an fp64 numpy restatement of the same objective
    finds from the same starting point.  Parity with cvxpy itself is unpinned (it cannot run here)."""
    from approx import backend as B
    from approx.core import fit
    torch.manual_seed(d)
    N, C = 12, 10
    Wc = torch.randn(N, C, d, d)
    W2, A0, B0 = fit.scheme1_init(Wc, M, 'svd')
    lm = [0.05, 0.2]
    A, Bs = A0.clone(), B0.clone()
    trace = B.lowrank1_alt_fit(W2, d, A, Bs, lm, 4, 1e-6, mm_rounds=8, fista_iters=60)
    assert len(trace) >= 4
    f_dev = O.lowrank1_alt_objective(W2.cpu(), A.cpu(), Bs.cpu(), lm[-1], d)
    assert abs(f_dev - trace[-1]) / f_dev < 1e-5                       # the kernel's own objective is the objective
    # within one lambda the alternation never increases f (block-coordinate descent on a convex-in-each-block function)
    n0 = None
    An, Bn, tr_np = O.lowrank1_alt_fit_np(W2.cpu(), A0.cpu(), B0.cpu(), lm, 4, 1e-6, d)
    f_np = tr_np[-1]
    assert f_dev <= f_np * (1 + 2e-3), (f_dev, f_np)
    assert f_dev >= f_np * (1 - 2e-2), (f_dev, f_np)                   # and it is not "better" by violating the model
    # single lambda: monotone
    A1, B1 = A0.clone(), B0.clone()
    t1 = B.lowrank1_alt_fit(W2, d, A1, B1, [0.1], 5, 0.0, mm_rounds=8, fista_iters=60)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(t1, t1[1:])), t1
    # lambda = 0 leaves the SVD starting point's fit quality intact (the SVD is optimal for the squared error, close for this one)
    A2, B2 = A0.clone(), B0.clone()
    B.lowrank1_alt_fit(W2, d, A2, B2, [0.0], 3, 0.0)
    e0 = (W2 - A0 @ B0).norm().item(); e2 = (W2 - A2 @ B2).norm().item()
    assert e2 <= e0 * 1.05


def test_lowrank_exp_v1_app_alternating_and_init_methods():
    """LowRankExpV1 with max_iter > 0 (the shipped low-rank-exp-v1_l2_alexnet.py schedule, shortened) and the 'standard' / 'random'
    starting points of the reference (core/low_rank_exp.py:100-125)."""
    from approx.core import LowRankExpV1
    torch.manual_seed(3)
    src = nn.Conv2d(6, 10, 3, 1, 1)
    x = torch.randn(2, 6, 9, 8)
    app = LowRankExpV1(num_bases=(4,), max_iter=3, lmda_length=3, min_lmda=0.0001, max_lmda=0.01, init_method='svd')
    def run(a, s_):
        t = a.postprocess(s_)
        t.s_conv.bias.data.zero_()      # the reference never sets s_conv's (randomly initialised) bias: zero it to compare fits
        return t.cuda()(x.cuda()).cpu()
    sub = app.initialize(src)
    app.optimize(sub)
    y = run(app, sub)
    ref = src(x)
    svd_app = LowRankExpV1(num_bases=(4,), max_iter=0, lmda_length=1, min_lmda=0.0, max_lmda=0.0)
    sub2 = svd_app.initialize(src); svd_app.optimize(sub2)
    y_svd = run(svd_app, sub2)
    e_alt = O.compare4d(ref.detach(), y.detach()); e_svd = O.compare4d(ref.detach(), y_svd.detach())
    assert e_alt < 1.3 * e_svd + 1e-3, (e_alt, e_svd)                  # a small penalty costs little accuracy
    # 'standard' with M = d*d is an exact re-expression of the conv: bases = unit filters, weights = the filters themselves
    app9 = LowRankExpV1(num_bases=(9,), max_iter=0, lmda_length=1, min_lmda=0.0, max_lmda=0.0, init_method='standard')
    sub9 = app9.initialize(src); app9.optimize(sub9)
    y9 = run(app9, sub9)
    assert O.compare4d(ref.detach(), y9.detach()) < 1e-5
    appr = LowRankExpV1(num_bases=(4,), max_iter=2, lmda_length=1, min_lmda=0.001, max_lmda=0.001, init_method='random')
    subr = appr.initialize(src); appr.optimize(subr)
    assert torch.isfinite(subr.new.d_conv.weight).all() and torch.isfinite(subr.new.s_conv.weight).all()


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"CNA_MSCA_WS": "0"}, {"CNA_MSCA_NIO": "8"}, {"CNA_MSCA_NIO": "4"}, {"CNA_MSCA_SD": "f"}, {"CNA_MSCA_SD": "o"}])
def test_msca_kernel_generations_agree_with_the_oracle(env):
    """The kernel generations behind the diagnostic switches (barrier-synchronised mma.sync kernel, 4 / 8 IO warps, FFMA2 16-row
    march, first-generation kernel) stay correct: each runs the original block and MscaRep d1+fix in its own process (the switches
    are read once per process) against the bf16 oracle."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, io, contextlib, torch
sys.path.insert(0, "convnet-approximater_b200"); sys.path.insert(0, ".")
from oracle import approx_oracle as O
from approx.layers import MSCA
from approx.core import MscaRep
torch.manual_seed(3)
C, H, W = 32, 37, 45
src = MSCA(C, 5, (7, 11, 21))
x = torch.randn(2, C, H, W)
dt = torch.bfloat16
sd = {k: v.detach().to(dt) for k, v in src.state_dict().items()}
br = [tuple(sd[f"sd_convs.branches.{i}.{n}"] for n in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")) for i in range(3)]
ref = O.msca_forward(x.to(dt), sd["conv0.weight"], sd["conv0.bias"], lambda a: O.parallel_forward(a, br, True), sd["channel_mix.weight"], sd["channel_mix.bias"])
import copy
y = copy.deepcopy(src).cuda().to(dt).eval()(x.cuda().to(dt))
e0 = O.compare4d(ref.float(), y.float().cpu())
app = MscaRep(decomp=1, fix=True)
sub = app.initialize(copy.deepcopy(src))
with contextlib.redirect_stdout(io.StringIO()):
    app.optimize(sub)
tgt = app.postprocess(sub)
# d1+fix: the bf16 kernel under test against this backend's fp32 kernels (a different code path, itself pinned by the golden tests)
y1 = copy.deepcopy(tgt).cuda().to(dt).eval()(x.cuda().to(dt))
y1f = copy.deepcopy(tgt).cuda().float().eval()(x.cuda())
e1 = O.compare4d(y1f.float().cpu(), y1.float().cpu())
print("ERR", e0, e1)
assert e0 < 1e-3, e0
assert e1 < 1e-2, e1
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, **env},
                       capture_output=True, text=True, timeout=200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_empty_batch_returns_empty_outputs_like_the_reference_layers():
    """A batch of zero images: the reference's layers (stacks of nn.Conv2d) return a batch of zero outputs; so do these, with the
    shape and dtype the stock convs give on the same empty input."""
    from approx.layers import LowRankExpConvV1, LowRankExpConvV2, MSCA, CascadeConv, ParallelConv, FixPaddingBias
    from torch import nn
    torch.manual_seed(0)
    cases = [
        (LowRankExpConvV2(16, 8, (3, 3), (1, 1), (1, 1), out_channels=32), (0, 16, 12, 10),
         nn.Sequential(nn.Conv2d(16, 8, (3, 1), 1, (1, 0), bias=False), nn.Conv2d(8, 32, (1, 3), 1, (0, 1)))),
        (LowRankExpConvV2(16, 8, (3, 3), (2, 2), (1, 1)), (0, 16, 12, 10),
         nn.Sequential(nn.Conv2d(16, 8, (3, 1), (2, 1), (1, 0), bias=False), nn.Conv2d(8, 8, (1, 3), (1, 2), (0, 1), groups=8))),
        (LowRankExpConvV1(8, 16, 3, 1, 1, num_base=4), (0, 8, 9, 9), nn.Conv2d(8, 16, 3, 1, 1)),
        (LowRankExpConvV1(8, 16, 5, 2, 2, num_base=3, decomp=True), (0, 8, 9, 9), nn.Conv2d(8, 16, 5, 2, 2)),
        (MSCA(32, 5, (7, 11, 21)), (0, 32, 16, 16), nn.Identity()),
        (CascadeConv(8, 7, 3, True, False), (0, 8, 6, 6), nn.Identity()),
        (ParallelConv(8, (7, 11), (3, 5), 2, True, True), (0, 8, 6, 6), nn.Identity()),
        (FixPaddingBias(8, 10), (0, 8, 6, 6), nn.Identity()),
    ]
    for dtype in (torch.float32, torch.bfloat16):
        for layer, shape, stock in cases:
            x = torch.zeros(shape, dtype=dtype)
            want = stock.to(dtype)(x)
            with torch.no_grad():
                got = layer.to("cuda", dtype).eval()(x.cuda())
            assert got.shape == want.shape and got.dtype == want.dtype and got.is_cuda, (type(layer).__name__, got.shape, want.shape)
