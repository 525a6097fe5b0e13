"""Generate golden input/output vectors by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not travel
to the GPU box):

    python tests/golden/make_golden.py

Writes small .npz fixtures next to this file.  The reference's `approx.core`
imports cvxpy and `approx.hooks.finetune` imports timm; neither is installed, so
permissive stand-ins are injected: with the shipped "svd" configs
(max_iter=0, lmda_length=1; e.g. configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py)
`LowRankExpV1.optimize` never calls a solver — the stand-in only has to carry
`.value` arrays through `_get_bi_problem` (core/low_rank_exp.py:110-120).
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CNA_REFERENCE_SRC", "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))


class _Anything:
    """Absorbs every cvxpy expression-building operation."""
    def __init__(self, *a, **k):
        self.value = None
    def _same(self, *a, **k):
        return _Anything()
    __add__ = __radd__ = __sub__ = __rsub__ = __mul__ = __rmul__ = _same
    __matmul__ = __rmatmul__ = __getitem__ = __call__ = _same
    def is_dcp(self):
        return True


def _install_stubs():
    cp = types.ModuleType("cvxpy")
    for name in ("Variable", "Parameter", "normNuc", "reshape", "sum", "norm2", "Minimize", "Problem"):
        setattr(cp, name, _Anything)
    sys.modules["cvxpy"] = cp

    class _Stub(types.ModuleType):
        def __getattr__(self, item):
            if item.startswith("__"):
                raise AttributeError(item)
            return _Anything
    for name in ("timm", "timm.data", "timm.models", "timm.utils", "timm.loss", "timm.optim",
                 "timm.scheduler", "timm.layers", "timm.models.layers", "timm.utils.model_ema", "ptflops",
                 "timm.bits", "timm.models.helpers"):
        m = _Stub(name)
        m.__path__ = []
        sys.modules[name] = m


def _np(sd):
    return {k: v.detach().cpu().float().numpy() for k, v in sd.items()}


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    from approx.layers import LowRankExpConvV1, LowRankExpConvV2, MSCA, Substitution
    from approx.core import MscaRep, LowRankExpV1
    from torch import nn

    torch.manual_seed(42)

    # ---- 1. LowRankExpConvV2 exactly as shipped (depthwise h_conv) ----------
    out = {}
    for tag, (C, K, k, s, p, shp) in {
        "a": (16, 8, (3, 3), (1, 1), (1, 1), (2, 16, 12, 10)),
        "b": (6, 4, (5, 3), (2, 1), (2, 1), (3, 6, 11, 9)),
    }.items():
        m = LowRankExpConvV2(C, K, k, s, p).eval()
        x = torch.randn(*shp)
        with torch.no_grad():
            y = m(x)
        out.update({f"{tag}_x": x.numpy(), f"{tag}_y": y.numpy(), f"{tag}_cfg": np.array([C, K, *k, *s, *p])})
        out.update({f"{tag}_{kk}": vv for kk, vv in _np(m.state_dict()).items()})
    np.savez_compressed(os.path.join(HERE, "lowrank2_shipped.npz"), **out)

    # ---- 2. LowRankExpConvV1: forward, decomp(), forward again -------------
    out = {}
    for tag, (C, N, k, s, p, M, shp) in {
        "a": (8, 12, 3, 1, 1, 4, (2, 8, 9, 11)),
        "b": (6, 10, 5, 2, 2, 3, (2, 6, 13, 12)),
    }.items():
        m = LowRankExpConvV1(C, N, k, s, p, num_base=M).eval()
        x = torch.randn(*shp)
        with torch.no_grad():
            y = m(x)
        out.update({f"{tag}_x": x.numpy(), f"{tag}_y": y.numpy(), f"{tag}_cfg": np.array([C, N, k, s, p, M])})
        out.update({f"{tag}_{kk}": vv for kk, vv in _np(m.state_dict()).items()})
        m.decomp()
        with torch.no_grad():
            y2 = m(x)
        out[f"{tag}_y_decomp"] = y2.numpy()
        out.update({f"{tag}_dec_{kk}": vv for kk, vv in _np(m.state_dict()).items()})
    np.savez_compressed(os.path.join(HERE, "lowrank1.npz"), **out)

    # ---- 3. LowRankExpV1 app: initialize -> optimize (svd) -> postprocess ---
    out = {}
    src = nn.Conv2d(6, 10, 3, 1, 1)
    app = LowRankExpV1(num_bases=(4,), max_iter=0, lmda_length=1, min_lmda=0.001, max_lmda=0.001,
                       init_method="svd", do_decomp=True)
    sub = app.initialize(src)
    app.optimize(sub)
    x = torch.randn(2, 6, 7, 8)
    out.update({"src_weight": src.weight.detach().numpy(), "src_bias": src.bias.detach().numpy(), "x": x.numpy()})
    out.update({f"fit_{kk}": vv for kk, vv in _np(sub.new.state_dict()).items()})
    tgt = app.postprocess(sub)
    out.update({f"post_{kk}": vv for kk, vv in _np(tgt.state_dict()).items()})
    with torch.no_grad():
        out["y_post"] = tgt(x).numpy()
        out["y_src"] = src(x).numpy()
    np.savez_compressed(os.path.join(HERE, "lowrank1_app.npz"), **out)

    # ---- 4. MSCA original + MscaRep variants --------------------------------
    out = {}
    C = 8
    msca = MSCA(C, 5, (7, 11, 21)).eval()
    xs = {"big": torch.randn(2, C, 24, 26), "small": torch.randn(2, C, 7, 9)}
    out.update({f"src_{kk}": vv for kk, vv in _np(msca.state_dict()).items()})
    with torch.no_grad():
        for t, x in xs.items():
            out[f"x_{t}"] = x.numpy()
            out[f"y_src_{t}"] = msca(x).numpy()
    w, b, r = MscaRep.get_equivalent_kernel(msca.sd_convs)
    out.update({"eq_weight": w.detach().numpy(), "eq_bias": b.detach().numpy(), "eq_res": r.detach().numpy()})
    for decomp, fix in [(0, True), (0, False), (1, True), (2, True), (4, True), (3, False)]:
        import io, contextlib
        app = MscaRep(decomp=decomp, fix=fix)
        sub = app.initialize(msca)
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            app.optimize(sub)
        tgt = app.postprocess(sub)
        tag = f"d{decomp}{'f' if fix else 'n'}"
        if buf.getvalue().strip():
            out[f"{tag}_pce"] = np.array(float(buf.getvalue().strip().rstrip(',')))
        out.update({f"{tag}_{kk}": vv for kk, vv in _np(tgt.state_dict()).items()})
        with torch.no_grad():
            for t, x in xs.items():
                out[f"{tag}_y_{t}"] = tgt(x.clone()).numpy()
    np.savez_compressed(os.path.join(HERE, "msca.npz"), **out)

    # ---- 5. reconstruction loss helper (hooks/finetune.py:76-82) ------------
    from approx.hooks.finetune import get_l2_error, cache_module_output
    sub = Substitution(nn.Identity(), nn.Identity())
    y_old = torch.randn(4, 5, 6, 7)
    y_new = torch.randn(4, 5, 6, 7)
    cache_module_output(sub, None, y_old)
    get_l2_error(sub, None, y_new)
    np.savez_compressed(os.path.join(HERE, "l2_recon.npz"), y_old=y_old.numpy(), y_new=y_new.numpy(),
                        norm=sub.cache["norm"].numpy())

    # ---- 6. bf16 behaviour of the reference modules (CPU bf16) --------------
    out = {}
    torch.manual_seed(7)
    m = LowRankExpConvV2(16, 8, (3, 3), (1, 1), (1, 1)).eval().to(torch.bfloat16)
    x = torch.randn(2, 16, 12, 10).to(torch.bfloat16)
    with torch.no_grad():
        out["v2_x"] = x.float().numpy()
        out["v2_y"] = m(x).float().numpy()
    out.update({f"v2_{kk}": vv for kk, vv in _np(m.state_dict()).items()})
    np.savez_compressed(os.path.join(HERE, "bf16.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
