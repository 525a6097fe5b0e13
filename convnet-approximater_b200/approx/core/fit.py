"""Device-side fitting primitives shared by the apps (thin wrappers over approx.backend).

The reference runs these on the CPU inside `Runner.run` (numpy / LAPACK); here the source weights are
moved to the current CUDA device, the factorisation kernels of csrc/fit.cu run there, and results are
returned on the device the source weights live on, so callers can keep assigning `.weight.data = ...`.
"""
import torch

from approx import backend as B


def _dev(t: torch.Tensor) -> torch.Tensor:
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("cna_b200: the fitting step needs a CUDA device — this backend has no CPU fallback")
    return t.cuda()


def rank1_split(s_w: torch.Tensor):
    """(C*M,1,d,d) basis filters -> v (C*M,1,d,1), h (C*M,1,1,d) with v@h the best rank-1 approximation
    (reference layers/low_rank_conv.py:61-65: u[...,0]*sqrt(s0), vh[...,0,:]*sqrt(s0))."""
    home = s_w.device
    u, s, vh = B.svd_small(_dev(s_w.detach()))
    r = torch.sqrt(s[..., 0])
    v = (u[..., 0] * r[..., None])[..., None]
    h = (vh[..., 0, :] * r[..., None])[..., None, :]
    return v.contiguous().to(home), h.contiguous().to(home)


def scheme1_svd(W: torch.Tensor, M: int):
    """Conv weight (N,C,d,d) -> s_conv.weight (C*M,1,d,d), d_conv.weight (N,C*M,1,1)
    (reference core/low_rank_exp.py:148-151, 110-120, 181-184)."""
    home = W.device
    N, C, d = W.shape[:3]
    assert W.shape[3] == d, "the reference assumes square kernels (core/low_rank_exp.py:149)"
    weights, bases = B.lowrank1_svd_fit(_dev(W.detach()).reshape(N * C, d * d), M)
    s_w = bases.reshape(1, M, d, d).expand(C, -1, -1, -1).reshape(C * M, 1, d, d).contiguous()
    d_w = weights.reshape(N, C * M, 1, 1).contiguous()
    return s_w.to(home), d_w.to(home)


def scheme1_init(W: torch.Tensor, M: int, init_method: str = 'svd', generator=None):
    """Starting point (weights (N*C, M), bases (M, d*d)) of the scheme-1 fit, on the device (reference core/low_rank_exp.py:98-125):
    'svd'      weights = U[:, :M] * sqrt(S[:M]),  bases = sqrt(S[:M]) * Vh[:M]     (truncated SVD of the (N*C) x d^2 filter matrix)
    'standard' weights = the first min(M, d^2) columns of the filter matrix, bases = the matching unit vectors
    'random'   both uniform in [0, 1)."""
    N, C, d = W.shape[:3]
    q = d * d
    W2 = _dev(W.detach()).to(torch.float32).reshape(N * C, q).contiguous()
    if init_method == 'svd':
        weights, bases = B.lowrank1_svd_fit(W2, M)
    elif init_method == 'standard':
        upb = min(M, q)
        weights = torch.zeros(N * C, M, device=W2.device)
        weights[:, :upb] = W2[:, :upb]
        bases = torch.zeros(M, q, device=W2.device)
        bases[torch.arange(upb), torch.arange(upb)] = 1.0
    elif init_method == 'random':
        weights = torch.rand(N * C, M, device=W2.device, generator=generator)
        bases = torch.rand(M, q, device=W2.device, generator=generator)
    else:
        raise AssertionError(f"Not supported init method: {init_method}")
    return W2, weights.contiguous(), bases.contiguous()


def scheme1_writeback(weights: torch.Tensor, bases: torch.Tensor, N: int, C: int, d: int, home):
    """(weights (N*C, M), bases (M, d*d)) -> s_conv.weight (C*M,1,d,d), d_conv.weight (N,C*M,1,1) (reference :181-184)."""
    M = bases.shape[0]
    s_w = bases.reshape(1, M, d, d).expand(C, -1, -1, -1).reshape(C * M, 1, d, d).contiguous()
    d_w = weights.reshape(N, C * M, 1, 1).contiguous()
    return s_w.to(home), d_w.to(home)


def pc_energy(bases: torch.Tensor) -> float:
    """Mean share of each basis' energy in its leading singular value (logged at low_rank_exp.py:177-180)."""
    _, s, _ = B.svd_small(_dev(bases.detach()))
    e = s ** 2
    return torch.mean(e[..., 0] / torch.sum(e, dim=-1)).item()


def msca_equivalent_kernel(par):
    """ParallelConv of strip cascades (+identity) -> merged (weight (C,1,K,K), bias (C,), res (2,C,K//2))."""
    cas = par.cascades()
    home = cas[0].conv1.weight.device
    w, b, r = B.msca_equiv_kernel([_dev(c.conv1.weight.detach()) for c in cas],
                                  [None if c.conv1.bias is None else _dev(c.conv1.bias.detach()) for c in cas],
                                  [_dev(c.conv2.weight.detach()) for c in cas],
                                  [None if c.conv2.bias is None else _dev(c.conv2.bias.detach()) for c in cas])
    return w.to(home), b.to(home), r.to(home)


def strip_factors(weight: torch.Tensor, rank: int):
    """Merged kernel (C,1,K,K) -> [(conv1 (C,1,1,K), conv2 (C,1,K,1))]*rank and the principal-component energy
    (reference core/msca_rep.py:154-171: conv1 = Vh[j] unscaled, conv2 = U[:,j]*s_j)."""
    home = weight.device
    u, s, vh = B.svd_small(_dev(weight.detach()))
    out = []
    for j in range(rank):
        c1 = vh[..., j, :][..., None, :].contiguous().to(home)
        c2 = (u[..., j] * s[..., j][..., None])[..., None].contiguous().to(home)
        out.append((c1, c2))
    e = s ** 2
    pce = torch.mean(torch.sum(e[..., :rank], dim=-1) / torch.sum(e, dim=-1)).item()
    return out, pce


def scheme2_svd(W: torch.Tensor, K: int):
    """Conv weight (N,C,d,d) -> v_conv.weight (K,C,d,1), h_conv.weight (N,K,1,d): the closed-form scheme-2 fit
    (Jaderberg et al. 2014, section 3.2): truncated SVD of A[(c,i),(j,n)] = W[n,c,i,j]; v = U*sqrt(S), h = sqrt(S)*Vh.
    The reference has no scheme-2 app; this is the analogue of its scheme-1 SVD fit (core/low_rank_exp.py:110-120)."""
    home = W.device
    N, C, dh, dw = W.shape
    A = _dev(W.detach()).to(torch.float32).permute(1, 2, 3, 0).reshape(C * dh, dw * N).contiguous()
    r = min(K, C * dh, dw * N)
    U, S, Vh = B.svd_topk(A, r)
    sq = torch.sqrt(S)
    v = torch.zeros(C * dh, K, device=A.device)
    h = torch.zeros(K, dw * N, device=A.device)
    v[:, :r] = U * sq[None, :]
    h[:r, :] = Vh * sq[:, None]
    v_w = v.reshape(C, dh, K).permute(2, 0, 1).unsqueeze(-1).contiguous()
    h_w = h.reshape(K, dw, N).permute(2, 0, 1).unsqueeze(2).contiguous()
    return v_w.to(home), h_w.to(home)
