"""Pins the CPU oracle (oracle/approx_oracle.py) against (a) the reference's own
test identities (tests/test_low_rank_exp.py:17-58 of the reference) and (b) the
golden vectors produced by the reference's modules (tests/golden/make_golden.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import approx_oracle as O

TOL = 1e-5  # the reference's fp32 threshold (tests/test_low_rank_exp.py:14)


def T(a):
    return torch.from_numpy(np.asarray(a)).float()


# ---- reference test identities, same shapes (reduced batch) ---------------
def test_ref_identity_linear_in_weight():
    torch.manual_seed(0)
    x = torch.randn(4, 256, 14, 14); w = torch.rand(128, 256, 3, 3)
    assert O.compare4d(F.conv2d(x, w) * 2, F.conv2d(x, w * 2)) < TOL


def test_ref_identity_rank1_depthwise_is_separable():
    torch.manual_seed(0)
    B, C, M, d = 4, 64, 30, 3
    x = torch.randn(B, C, 14, 14)
    v_w = torch.randn(M * C, 1, d, 1); h_w = torch.randn(M * C, 1, 1, d)
    y1 = F.conv2d(x, v_w @ h_w, groups=C)
    # the reference test applies h first, then v; SeparableConv applies v first: both equal y1
    y2 = O.separable_forward(x, v_w, h_w)
    assert O.compare4d(y1, y2) < TOL


def test_ref_identity_scheme1_layout():
    torch.manual_seed(0)
    N, C, M, d = 32, 64, 16, 5
    x = torch.randn(4, C, 14, 14)
    weights = torch.rand(N * C, M); bases = torch.rand(M, d * d)
    W = (weights @ bases).reshape(N, C, d, d)
    s_w = bases.reshape(M, d, d).unsqueeze(0).expand(C, -1, -1, -1).reshape(C * M, 1, d, d)
    d_w = weights.reshape(N, C * M, 1, 1)
    y2 = O.lowrank1_forward(x, s_w, None, d_w, None)
    assert O.compare4d(F.conv2d(x, W), y2) < TOL


def test_direct_numpy_conv_pins_torch_semantics():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 6, 9, 8)); w = rng.standard_normal((12, 2, 3, 5)); b = rng.standard_normal(12)
    y = O.conv2d_direct_np(x, w, b, (2, 1), (1, 2), groups=3)
    yt = F.conv2d(T(x).double(), T(w).double(), T(b).double(), (2, 1), (1, 2), groups=3)
    np.testing.assert_allclose(y, yt.numpy(), rtol=1e-5, atol=1e-5)


# ---- golden vectors from the reference's modules ---------------------------
def test_golden_lowrank2_shipped(golden):
    g = golden("lowrank2_shipped")
    for tag in "ab":
        C, K, kh, kw, sh, sw, ph, pw = g[f"{tag}_cfg"]
        y = O.lowrank2_forward(T(g[f"{tag}_x"]), T(g[f"{tag}_v_conv.weight"]), T(g[f"{tag}_h_conv.weight"]),
                               T(g[f"{tag}_h_conv.bias"]), (sh, sw), (ph, pw))
        assert O.compare4d(T(g[f"{tag}_y"]), y) < TOL


def test_golden_lowrank1_forward_and_decomp(golden):
    g = golden("lowrank1")
    for tag in "ab":
        C, N, k, s, p, M = g[f"{tag}_cfg"]
        x = T(g[f"{tag}_x"])
        y = O.lowrank1_forward(x, T(g[f"{tag}_s_conv.weight"]), T(g[f"{tag}_s_conv.bias"]),
                               T(g[f"{tag}_d_conv.weight"]), T(g[f"{tag}_d_conv.bias"]), (s, s), (p, p))
        assert O.compare4d(T(g[f"{tag}_y"]), y) < TOL
        v, h = O.lowrank1_decomp(T(g[f"{tag}_s_conv.weight"]))
        # factors are unique up to sign: compare the rank-1 products and the layer output
        ref = T(g[f"{tag}_dec_s_conv.v_conv.weight"]) @ T(g[f"{tag}_dec_s_conv.h_conv.weight"])
        assert torch.allclose(v @ h, ref, atol=1e-5)
        y2 = O.lowrank1_forward_decomp(x, v, h, T(g[f"{tag}_d_conv.weight"]), T(g[f"{tag}_d_conv.bias"]),
                                       (s, s), (p, p))
        assert O.compare4d(T(g[f"{tag}_y_decomp"]), y2) < TOL


def test_golden_lowrank1_app_svd_fit(golden):
    g = golden("lowrank1_app")
    s_w, d_w = O.lowrank1_svd_fit(T(g["src_weight"]), 4)
    # compare reconstructions weights@bases (sign/rotation ambiguity of SVD)
    C, M = 6, 4
    def recon(sw, dw):
        bases = sw.reshape(C, M, 9)[0]
        return dw.reshape(-1, M) @ bases
    assert torch.allclose(recon(s_w, d_w), recon(T(g["fit_s_conv.weight"]), T(g["fit_d_conv.weight"])), atol=1e-5)
    v, h = O.lowrank1_decomp(T(g["fit_s_conv.weight"]))
    y = O.lowrank1_forward_decomp(T(g["x"]), v, h, T(g["fit_d_conv.weight"]), T(g["src_bias"]), (1, 1), (1, 1))
    assert O.compare4d(T(g["y_post"]), y) < TOL


def _src_branches(g, prefix="src_"):
    br = []
    for i in range(3):
        br.append(tuple(T(g[f"{prefix}sd_convs.branches.{i}.{n}"]) for n in
                        ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias")))
    return br


def test_golden_msca_original(golden):
    g = golden("msca")
    br = _src_branches(g)
    for t in ("big", "small"):
        y = O.msca_forward(T(g[f"x_{t}"]), T(g["src_conv0.weight"]), T(g["src_conv0.bias"]),
                           lambda a: O.parallel_forward(a, br, True),
                           T(g["src_channel_mix.weight"]), T(g["src_channel_mix.bias"]))
        assert O.compare4d(T(g[f"y_src_{t}"]), y) < TOL


def test_golden_msca_equivalent_kernel(golden):
    g = golden("msca")
    w, b, r = O.msca_equivalent_kernel(_src_branches(g), 8)
    assert torch.allclose(w, T(g["eq_weight"]), atol=1e-6)
    assert torch.allclose(b, T(g["eq_bias"]), atol=1e-6)
    assert torch.allclose(r, T(g["eq_res"]), atol=1e-6)


def test_golden_msca_rep_variants(golden):
    g = golden("msca")
    w, b, r = O.msca_equivalent_kernel(_src_branches(g), 8)
    C = 8
    for decomp, fix in [(0, True), (0, False), (1, True), (2, True), (4, True), (3, False)]:
        tag = f"d{decomp}{'f' if fix else 'n'}"
        if decomp == 0:
            def sd(a):
                y = F.conv2d(a, w, b, padding=10, groups=C)
                return O.fix_padding_bias(y, r, 10) if fix else y
        else:
            fac, pce = O.msca_rep_factors(w, decomp)
            assert abs(pce - float(g[f"{tag}_pce"])) < 1e-5
            def sd(a):
                brs = [(c1, None, c2, b if j == decomp - 1 else None) for j, (c1, c2) in enumerate(fac)]
                y = O.parallel_forward(a, brs, False)
                return O.fix_padding_bias(y, r, 10) if fix else y
        for t in ("big", "small"):
            y = O.msca_forward(T(g[f"x_{t}"]), T(g["src_conv0.weight"]), T(g["src_conv0.bias"]), sd,
                               T(g["src_channel_mix.weight"]), T(g["src_channel_mix.bias"]))
            assert O.compare4d(T(g[f"{tag}_y_{t}"]), y) < 2e-5, tag
    # known-answer property: d0+fix and d4+fix reproduce the original block
    for tag in ("d0f", "d4f"):
        for t in ("big", "small"):
            assert O.compare4d(T(g[f"y_src_{t}"]), T(g[f"{tag}_y_{t}"])) < 2e-5


def test_golden_l2_recon(golden):
    g = golden("l2_recon")
    n = O.l2_recon_norm(T(g["y_new"]), T(g["y_old"]))
    assert torch.allclose(n, T(g["norm"]), rtol=1e-6)


def test_lowrank2_svd_fit_full_rank_is_exact():
    torch.manual_seed(1)
    W = torch.randn(10, 6, 3, 3)
    v_w, h_w = O.lowrank2_svd_fit(W, 18)  # rank of the (C*d) x (d*N) reshaping <= 18
    x = torch.randn(2, 6, 9, 8); b = torch.randn(10)
    y = O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))
    assert O.compare4d(F.conv2d(x, W, b, 1, 1), y) < TOL


def test_lowrank2_svd_fit_attains_the_eckart_young_optimum():
    """True scheme 2 is absent from the reference, so its fit has no reference output to match; it is pinned to the optimum of the
    problem instead (Jaderberg et al. 2014, section 3.2: min ||W - sum_k v_k * h_k||_F over K separable terms).  The effective d x d
    filters of the fitted pair, W~[n,c,i,j] = sum_k v[k,c,i] h[n,k,j], miss W by exactly the discarded singular values of the
    (C d) x (d N) reshaping (Eckart-Young), no rank-K pair of factors does better (random perturbations only increase the error),
    and the two-conv forward equals a dense conv with W~."""
    torch.manual_seed(3)
    N, C, d, K = 10, 6, 3, 7
    W = torch.randn(N, C, d, d)
    v_w, h_w = O.lowrank2_svd_fit(W, K)
    Wt = torch.einsum("kci,nkj->ncij", v_w[..., 0].double(), h_w[:, :, 0, :].double())
    s = torch.linalg.svdvals(W.permute(1, 2, 3, 0).reshape(C * d, d * N).double())
    err = (W.double() - Wt).norm()
    assert abs(err - s[K:].norm()) < 1e-5 * s.norm(), (float(err), float(s[K:].norm()))
    for t in range(20):
        dv, dh = 1e-2 * torch.randn_like(v_w), 1e-2 * torch.randn_like(h_w)
        Wp = torch.einsum("kci,nkj->ncij", (v_w + dv)[..., 0].double(), (h_w + dh)[:, :, 0, :].double())
        assert (W.double() - Wp).norm() >= err - 1e-9
    x = torch.randn(2, C, 9, 8); b = torch.randn(N)
    assert O.compare4d(F.conv2d(x, Wt.float(), b, 1, 1), O.lowrank2_forward(x, v_w, h_w, b, (1, 1), (1, 1))) < TOL


def test_golden_bf16_oracle_mode(golden):
    g = golden("bf16")
    bf = lambda a: T(a).to(torch.bfloat16)
    y = O.lowrank2_forward(bf(g["v2_x"]), bf(g["v2_v_conv.weight"]), bf(g["v2_h_conv.weight"]),
                           bf(g["v2_h_conv.bias"]), (1, 1), (1, 1))
    assert O.compare4d(T(g["v2_y"]), y.float()) < 1e-3


def _svd_start(W, M):
    """The reference's 'svd' starting point (core/low_rank_exp.py:110-120)."""
    u, s, vh = np.linalg.svd(W, full_matrices=False)
    k = min(M, s.shape[0])
    A0, B0 = np.zeros((W.shape[0], M)), np.zeros((M, W.shape[1]))
    A0[:, :k] = u[:, :k] * np.sqrt(s[:k])
    B0[:k] = vh[:k] * np.sqrt(s[:k])[:, None]
    return A0, B0


def test_alternating_fit_steps_satisfy_the_optimality_conditions_of_the_cvxpy_programs():
    """The reference solves `min_B sum_r ||W_r - A_r B||_2 + lmda sum_m ||B_m||_*` and `min_A ...` with cvxpy
    (core/low_rank_exp.py:126-135, 160-176), which cannot run here.  The oracle's two steps are pinned to those convex programs
    themselves: after a B-step (run to convergence) the bases satisfy the subgradient optimality conditions (nuclear-norm subdifferential)
    to 1e-9, after an
    A-step every row is a least-squares point — so each step returns an optimum of the program cvxpy is handed, whatever solver
    finds it.  Also: the objective never increases along the trace, and a larger lambda gives bases of lower rank."""
    rng = np.random.default_rng(0)
    for d, M, R in ((3, 4, 200), (5, 8, 300), (3, 6, 120)):
        W = rng.standard_normal((R, d * d))
        A0, B0 = _svd_start(W, M)
        ranks = []
        for lmda in (0.05, 1.0, 5.0):
            # one alternation, stopped after its B-step (epsilon = inf) ...
            _, B1, tr = O.lowrank1_alt_fit_np(W, A0, B0, [lmda], 1, np.inf, d, inner=2000)
            res_B, _ = O.lowrank1_alt_kkt_residual(W, A0, B1, lmda, d)
            assert res_B < 1e-9, (d, M, lmda, res_B)
            assert tr[0] <= O.lowrank1_alt_objective(W, A0, B0, lmda, d) + 1e-9 and tr[1] <= tr[0] + 1e-9
            # ... and the pair after the A-step
            A1, B1b, _ = O.lowrank1_alt_fit_np(W, A0, B0, [lmda], 1, np.inf, d, inner=2000)
            assert np.array_equal(B1, B1b)
            _, res_A = O.lowrank1_alt_kkt_residual(W, A1, B1, lmda, d)
            assert res_A < 1e-8, (d, M, lmda, res_A)
            # several alternations: monotone objective
            _, B5, tr5 = O.lowrank1_alt_fit_np(W, A0, B0, [lmda], 5, 0.0, d)
            assert all(b <= a + 1e-8 for a, b in zip(tr5, tr5[1:])), tr5
            ranks.append(sum(int((np.linalg.svd(B5[m].reshape(d, d), compute_uv=False) > 1e-8).sum()) for m in range(M)))
        assert ranks[0] >= ranks[1] >= ranks[2] and ranks[2] < ranks[0], ranks
    # a perturbed optimum is detected
    _, B1, _ = O.lowrank1_alt_fit_np(W, A0, B0, [1.0], 1, np.inf, d, inner=2000)
    bad, _ = O.lowrank1_alt_kkt_residual(W, A0, B1 + 1e-3 * rng.standard_normal(B1.shape), 1.0, d)
    assert bad > 1e-4
