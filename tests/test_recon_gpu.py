"""Reconstruction step and scheme-2 fit on the GPU against torch autograd / the oracle on the same seeded inputs."""
import copy

import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import approx_oracle as O

pytestmark = pytest.mark.gpu


def _torch_scheme2(C, K, N, seed):
    torch.manual_seed(seed)
    v = nn.Conv2d(C, K, (3, 1), 1, (1, 0), bias=False)
    h = nn.Conv2d(K, N, (1, 3), 1, (0, 1))
    return v, h


def test_l2_recon_loss_and_grad():
    from approx import backend as B
    torch.manual_seed(0)
    y_new = torch.randn(5, 7, 6, 9, requires_grad=True)
    y_old = torch.randn(5, 7, 6, 9)
    y_old[2] = y_new[2].detach()                       # zero-norm sample: gradient must be 0, not NaN
    norms_ref = O.l2_recon_norm(y_new, y_old)
    (0.3 * norms_ref.sum()).backward()
    norms, g = B.l2_recon(y_new.detach().cuda(), y_old.cuda(), 0.3)
    assert torch.allclose(norms.cpu(), norms_ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(g.cpu(), y_new.grad, rtol=1e-4, atol=1e-7)
    assert torch.isfinite(g).all()


@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("dims", [(2, 5, 9, 11, 7, 3), (3, 33, 8, 6, 70, 5), (1, 64, 14, 14, 64, 3)])
def test_tap_conv_gradients_vs_autograd(axis, dims):
    from approx import backend as B
    Bn, C, H, W, N, taps = dims
    torch.manual_seed(1)
    x = torch.randn(Bn, C, H, W, requires_grad=True)
    w = torch.randn(N, C, *((taps, 1) if axis == 0 else (1, taps)), requires_grad=True)
    b = torch.randn(N, requires_grad=True)
    pad = taps // 2
    y = F.conv2d(x, w, b, 1, (pad, 0) if axis == 0 else (0, pad))
    g = torch.randn_like(y)
    y.backward(g)
    xn, gn = B.to_nhwc(x.detach().cuda()).contiguous(), B.to_nhwc(g.cuda()).contiguous()
    dw, db = B.conv_taps_wgrad(xn, gn, taps, axis, pad, want_bias=True)
    assert torch.allclose(dw.cpu().reshape(w.shape), w.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(db.cpu(), b.grad, rtol=1e-4, atol=1e-4)
    wt = B.pack_taps_dgrad(w.detach().cuda(), torch.float32)
    gx = B.conv_taps(gn, wt, None, C, taps, axis, (1, 1), taps - 1 - pad)
    assert O.compare4d(x.grad, gx.permute(0, 3, 1, 2).cpu()) < 1e-5


def test_adamw_matches_torch():
    from approx import backend as B
    torch.manual_seed(2)
    p0 = torch.randn(1000)
    p_ref = nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p = p0.clone().cuda(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(1000)
        p_ref.grad = g.clone()
        opt.step()
        B.adamw_step_(p, g.cuda(), m, v, step, 1e-2)
    assert torch.allclose(p.cpu(), p_ref.detach(), rtol=1e-5, atol=1e-6)


def _ref_steps(layer_state, C, K, N, xs, ys, lr, steps):
    v, h = _torch_scheme2(C, K, N, 0)
    v.weight.data = layer_state["v_conv.weight"].clone(); h.weight.data = layer_state["h_conv.weight"].clone()
    h.bias.data = layer_state["h_conv.bias"].clone()
    opt = torch.optim.AdamW(list(v.parameters()) + list(h.parameters()), lr=lr, weight_decay=0.01)
    losses = []
    for s in range(steps):
        y = h(v(xs[s]))
        loss = O.l2_recon_loss([O.l2_recon_norm(y, ys[s])])
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(loss.item())
    return v, h, losses


def test_reconstructor_matches_torch_training_fp32():
    from approx.layers import LowRankExpConvV2
    from approx.recon import Scheme2Reconstructor
    C, K, N = 12, 6, 20
    torch.manual_seed(3)
    layer = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N)
    state = {k: v.clone() for k, v in layer.state_dict().items()}
    teacher = nn.Conv2d(C, N, 3, 1, 1)
    xs = [torch.randn(4, C, 10, 9) for _ in range(3)]
    with torch.no_grad():
        ys = [teacher(x) for x in xs]
    v, h, losses = _ref_steps(state, C, K, N, xs, ys, 1e-2, 3)
    layer = layer.cuda()
    rec = Scheme2Reconstructor(layer, lr=1e-2)
    from approx import backend as B
    got = []
    for x, y in zip(xs, ys):
        got.append(rec.step(B.to_nhwc(x.cuda()).contiguous(), B.to_nhwc(y.cuda()).contiguous()).item())
    assert torch.allclose(torch.tensor(got), torch.tensor(losses), rtol=1e-4)
    assert torch.allclose(layer.v_conv.weight.cpu(), v.weight, rtol=1e-3, atol=2e-5)
    assert torch.allclose(layer.h_conv.weight.cpu(), h.weight, rtol=1e-3, atol=2e-5)
    assert torch.allclose(layer.h_conv.bias.cpu(), h.bias, rtol=1e-3, atol=2e-5)


def test_reconstruction_reduces_error_bf16_fused_path():
    """bf16 layer on the fused tcgen05 forward: the step must drive the per-sample output error down."""
    from approx import backend as B
    from approx.layers import LowRankExpConvV2
    from approx.recon import Scheme2Reconstructor
    torch.manual_seed(4)
    C, K, N = 64, 16, 64
    teacher = nn.Conv2d(C, N, 3, 1, 1).cuda().bfloat16()
    layer = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).cuda().bfloat16()
    rec = Scheme2Reconstructor(layer, lr=5e-3)
    x = torch.randn(8, C, 28, 28, device="cuda").bfloat16()
    with torch.no_grad():
        y = teacher(x)
    xn, yn = B.to_nhwc(x).contiguous(), B.to_nhwc(y).contiguous()
    first = rec.step(xn, yn).item()
    for _ in range(40):
        last = rec.step(xn, yn).item()
    assert last < 0.9 * first


def test_module_autograd_bridge():
    from approx.layers import LowRankExpConvV2
    torch.manual_seed(5)
    C, K, N = 8, 4, 16
    layer = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N)
    v, h = _torch_scheme2(C, K, N, 0)
    v.weight.data = layer.v_conv.weight.data.clone(); h.weight.data = layer.h_conv.weight.data.clone()
    h.bias.data = layer.h_conv.bias.data.clone()
    x = torch.randn(3, C, 9, 7)
    xr = x.clone().requires_grad_(True)
    tgt = torch.randn(3, N, 9, 7)
    loss_ref = torch.mean(O.l2_recon_norm(h(v(xr)), tgt)); loss_ref.backward()
    layer = layer.cuda()
    xg = x.cuda().requires_grad_(True)
    loss = torch.mean(torch.norm((layer(xg) - tgt.cuda()).reshape(3, -1), p=2, dim=1)); loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    assert torch.allclose(layer.v_conv.weight.grad.cpu(), v.weight.grad, rtol=1e-3, atol=1e-6)
    assert torch.allclose(layer.h_conv.weight.grad.cpu(), h.weight.grad, rtol=1e-3, atol=1e-6)
    assert torch.allclose(layer.h_conv.bias.grad.cpu(), h.bias.grad, rtol=1e-3, atol=1e-6)
    assert O.compare4d(xr.grad, xg.grad.cpu()) < 1e-5


def test_svd_topk_and_scheme2_fit():
    from approx import backend as B
    from approx.core import LowRankExpV2
    torch.manual_seed(6)
    A = torch.randn(90, 60)
    U, S, Vh = B.svd_topk(A.cuda(), 60)
    assert torch.allclose((U * S) @ Vh, A.cuda(), atol=1e-4)
    assert torch.allclose(S.cpu(), torch.linalg.svdvals(A), rtol=1e-5, atol=1e-5)
    U, S, Vh = B.svd_topk(A.t().contiguous().cuda(), 10)        # wide orientation, truncated
    best = torch.linalg.svd(A.t(), full_matrices=False)
    assert torch.allclose(((U * S) @ Vh).cpu(), (best[0][:, :10] * best[1][:10]) @ best[2][:10], atol=1e-4)
    # app flow: full-rank fit reproduces the source conv, truncated fit matches the oracle's closed form
    src = nn.Conv2d(6, 10, 3, 1, 1)
    x = torch.randn(2, 6, 9, 8)
    for K in (18, 5):
        app = LowRankExpV2(num_bases=(K,))
        sub = app.initialize(src)
        app.optimize(sub)
        tgt = app.postprocess(sub)
        v_o, h_o = O.lowrank2_svd_fit(src.weight.data, K)
        ref = O.lowrank2_forward(x, v_o, h_o, src.bias.data, (1, 1), (1, 1))
        y = tgt.cuda()(x.cuda())
        assert O.compare4d(ref, y.detach().cpu()) < 2e-5
        if K == 18:
            assert O.compare4d(src(x).detach(), y.detach().cpu()) < 2e-5


@pytest.mark.parametrize("case", [(2, 64, 20, 30, 64), (1, 256, 9, 17, 64), (2, 128, 28, 14, 32), (1, 512, 14, 14, 64)])
def test_wgrad3_tcgen05_matches_autograd(case):
    """tcgen05 weight gradients of v_conv (taps along H, wide = input) and h_conv (taps along W, wide = output gradient)
    against torch autograd on the same bf16 tensors (fp32 accumulation on both sides)."""
    import torch.nn.functional as F
    from approx import backend as B
    from oracle import approx_oracle as O
    Bn, C, H, W, K = case
    g = torch.Generator().manual_seed(23)
    x = torch.randn(Bn, C, H, W, generator=g).bfloat16()
    gt = torch.randn(Bn, K, H, W, generator=g).bfloat16()
    kp = 16 if K <= 16 else 32 if K <= 32 else 64
    # v_conv: dw[k][c][t] = sum x[h+t-1] * gt[h]
    wv = torch.zeros(K, C, 3, 1, requires_grad=True)
    F.conv2d(x.float(), wv, None, 1, (1, 0)).backward(gt.float())
    xn = B.to_nhwc(x.cuda()).contiguous()
    gtn = B.to_nhwc(gt.cuda(), kp).contiguous()
    dv = torch.zeros(K, C, 3, 1, device="cuda")
    B.wgrad3_tc(xn, gtn, dv, C, K, 0, 1)
    err = (dv.cpu() - wv.grad).norm() / wv.grad.norm()
    assert err < 2e-3, err
    # h_conv: N = C wide output channels, T narrow: dw[n][k][t] = sum T[w+t-1][k] * g[w][n]
    wh = torch.zeros(C, K, 1, 3, requires_grad=True)
    F.conv2d(gt.float(), wh, None, 1, (0, 1)).backward(x.float())
    dh = torch.zeros(C, K, 1, 3, device="cuda")
    B.wgrad3_tc(xn, gtn, dh, C, K, 1, 0)
    err = (dh.cpu() - wh.grad).norm() / wh.grad.norm()
    assert err < 2e-3, err


def test_graph_replay_matches_eager_steps():
    """A captured-and-replayed reconstruction step (device-side AdamW step counter, weights re-packed inside the graph)
    follows the eager step sequence."""
    import copy
    from approx.layers import LowRankExpConvV2
    from approx.recon import Scheme2Reconstructor
    torch.manual_seed(5)
    C, K, N = 64, 32, 128
    teacher = nn.Conv2d(C, N, 3, 1, 1).cuda().bfloat16()
    base = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).cuda().bfloat16()
    x = torch.randn(4, 28, 28, C, device="cuda").bfloat16()
    with torch.no_grad():
        y_old = teacher(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1).contiguous()
    eager = Scheme2Reconstructor(copy.deepcopy(base), lr=1e-3)
    for _ in range(4):
        eager.step(x, y_old)
    graphed = Scheme2Reconstructor(copy.deepcopy(base), lr=1e-3)
    graphed.capture(x, y_old)            # runs 2 warm-up steps, then captures
    for _ in range(2):
        loss = graphed.replay()
    torch.cuda.synchronize()
    assert int(graphed.step_dev.item()) == 4
    rel = (graphed.master - eager.master).norm() / eager.master.norm()
    assert rel < 2e-3, rel               # same kernels; eager packs from the bf16 parameters, the graph from the fp32 master
    assert torch.isfinite(loss)


@pytest.mark.parametrize("case", [(4, 64, 28, 28, 32, 128), (2, 128, 14, 30, 64, 256)])
def test_tensor_core_gradients_match_autograd_bf16(case):
    """Scheme2Reconstructor.gradients on the tcgen05 path (forward + T, fused loss/bias gradient, MN-major weight gradients,
    3-tap data gradient) against torch autograd through the same bf16 layer."""
    from approx.layers import LowRankExpConvV2
    from approx.recon import Scheme2Reconstructor
    from approx import backend as B
    Bn, C, H, W, K, N = case
    torch.manual_seed(31)
    layer = LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).cuda().bfloat16()
    x = torch.randn(Bn, C, H, W, device="cuda").bfloat16()
    y_old = torch.randn(Bn, N, H, W, device="cuda").bfloat16()
    # autograd reference with stock convs on the same parameters (bf16 forward like the reference's modules)
    v = layer.v_conv.weight.detach().clone().requires_grad_(True)
    h = layer.h_conv.weight.detach().clone().requires_grad_(True)
    b = layer.h_conv.bias.detach().clone().requires_grad_(True)
    y = F.conv2d(F.conv2d(x, v, None, 1, (1, 0)), h, b, 1, (0, 1))
    loss = (y.float() - y_old.float()).flatten(1).norm(dim=1).mean()
    loss.backward()
    rec = Scheme2Reconstructor(layer)
    norms = rec.gradients(B.to_nhwc(x, B.round_up(C, 16)).contiguous(), B.to_nhwc(y_old).contiguous())
    dv, dh, db = rec._views(rec.grad)
    assert abs(norms.mean().item() - loss.item()) / loss.item() < 5e-3
    for got, ref in ((dv, v.grad), (dh, h.grad), (db, b.grad)):
        ref = ref.float().reshape(got.shape)
        assert (got - ref).norm() / ref.norm() < 3e-2, ((got - ref).norm() / ref.norm()).item()


def _adamw_ref(p, g, lr, wd_mask, wd=0.01, betas=(0.9, 0.999), eps=1e-8, step=1, m=None, v=None):
    """One torch.optim.AdamW step on a flat tensor with per-element decay mask (biases: no decay, like timm's filter)."""
    m = torch.zeros_like(p) if m is None else m
    v = torch.zeros_like(p) if v is None else v
    p = p * (1 - lr * wd * wd_mask)
    m = betas[0] * m + (1 - betas[0]) * g
    v = betas[1] * v + (1 - betas[1]) * g * g
    denom = v.sqrt() / (1 - betas[1] ** step) ** 0.5 + eps
    return p - lr / (1 - betas[0] ** step) * m / denom, m, v


def _layerwise_setup(seed=41, Bn=4):
    from approx import backend as B
    from approx.layers import LowRankExpConvV2
    torch.manual_seed(seed)
    shapes = [(64, 32, 128, 28), (16, 16, 64, 20), (128, 64, 128, 14)]          # (C, K, N, H)
    layers = [LowRankExpConvV2(C, K, (3, 3), (1, 1), (1, 1), out_channels=N).cuda().bfloat16() for (C, K, N, H) in shapes]
    xs = [torch.randn(Bn, H, H, C, device="cuda").bfloat16() for (C, K, N, H) in shapes]
    ys = [torch.randn(Bn, H, H, N, device="cuda").bfloat16() for (C, K, N, H) in shapes]
    return shapes, layers, xs, ys


def test_layerwise_reconstructor_step_matches_autograd_and_adamw():
    """LayerwiseReconstructor (all layers in one flat buffer, one fused exchange + AdamW + re-pack kernel; world size 1 here):
    gradients against torch autograd through stock bf16 convs, the update against AdamW with no decay on biases, and the re-packed
    operands against a fresh pack of the updated parameters."""
    from approx import backend as B
    from approx.recon import LayerwiseReconstructor
    shapes, layers, xs, ys = _layerwise_setup()
    rec = LayerwiseReconstructor(layers, lr=1e-3, weight_decay=0.01, l2_weight=1.0)
    rec.bind(xs, ys)
    p0 = rec.param.clone()
    rec.step(parallel=True)
    torch.cuda.synchronize()
    rec.check()
    L, Bn = len(layers), xs[0].shape[0]
    ref_g = torch.zeros_like(p0)
    mask = torch.ones_like(p0)
    for l, ((C, K, N, H), lay, x, yo) in enumerate(zip(shapes, layers, xs, ys)):
        ov, oh, ob, end = rec.offsets[l]
        v = p0[ov:oh].view(K, C, 3, 1).bfloat16().requires_grad_(True)
        h = p0[oh:ob].view(N, K, 1, 3).bfloat16().requires_grad_(True)
        b = p0[ob:end].bfloat16().requires_grad_(True)
        y = F.conv2d(F.conv2d(x.permute(0, 3, 1, 2), v, None, 1, (1, 0)), h, b, 1, (0, 1))
        norms = (y.float() - yo.permute(0, 3, 1, 2).float()).flatten(1).norm(dim=1)
        (norms.sum() / (Bn * L)).backward()
        assert torch.allclose(rec.norms[l], norms.detach(), rtol=1e-2)
        for (lo_, hi_), t in (((ov, oh), v), ((oh, ob), h), ((ob, end), b)):
            ref_g[lo_:hi_] = t.grad.float().reshape(-1)
            rel = (rec.grad[lo_:hi_] - ref_g[lo_:hi_]).norm() / ref_g[lo_:hi_].norm()
            assert rel < 3e-2, (l, rel.item())
        mask[ob:end] = 0
    want, _, _ = _adamw_ref(p0, rec.grad.clone(), 1e-3, mask)
    assert torch.allclose(rec.param, want, rtol=1e-5, atol=1e-7)
    assert rec.steps_taken() == 1
    for l, ((C, K, N, H), (wv, wh, wd, bias)) in enumerate(zip(shapes, rec.packs)):      # operands re-packed by the same kernel
        ov, oh, ob, end = rec.offsets[l]
        wv2, wh2 = B.lowrank2_pack(rec.param[ov:oh].view(K, C, 3, 1), rec.param[oh:ob].view(N, K, 1, 3))
        assert torch.equal(wv, wv2) and torch.equal(wh, wh2)
        assert torch.equal(wd, B.lowrank2_pack(rec.param[oh:ob].view(N, K, 1, 3).permute(1, 0, 3, 2).flip(2).contiguous(),
                                               torch.zeros(64, K, 1, 3, device="cuda"))[0])
        assert torch.equal(bias, rec.param[ob:end])
    rec.sync_to_layers()
    assert torch.allclose(layers[0].h_conv.bias.float(), rec.param[rec.offsets[0][2]:rec.offsets[0][3]], rtol=1e-2)
    rec.close()


def test_layerwise_reconstructor_graph_replay_follows_eager():
    """Captured graph (one branch per layer, warm-up that does not advance the fit) == the eager multi-stream steps."""
    import copy
    from approx.recon import LayerwiseReconstructor
    shapes, layers, xs, ys = _layerwise_setup(seed=43)
    eager = LayerwiseReconstructor(copy.deepcopy(layers), lr=1e-3)
    eager.bind(xs, ys)
    for _ in range(3):
        eager.step(parallel=False)
    graphed = LayerwiseReconstructor(copy.deepcopy(layers), lr=1e-3)
    graphed.bind(xs, ys)
    graphed.capture(parallel=True)
    assert graphed.steps_taken() == 0
    for _ in range(3):
        graphed.replay()
    torch.cuda.synchronize()
    graphed.check()
    assert graphed.steps_taken() == eager.steps_taken() == 3
    rel = (graphed.param - eager.param).norm() / eager.param.norm()
    assert rel < 1e-4, rel.item()            # same kernels; split-K atomics make the gradients order-dependent in the last bits
    assert abs(float(graphed.loss()) - float(eager.loss())) / float(eager.loss()) < 1e-3
    first = float(graphed.loss())
    for _ in range(30):
        graphed.replay()
    assert float(graphed.loss()) < first     # the fit goes down
    eager.close(); graphed.close()
