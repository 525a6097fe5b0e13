"""Low-rank expansion apps (Jaderberg et al., arXiv:1405.3866).

`LowRankExpV1` mirrors the reference's src/approx/core/low_rank_exp.py:12-189 (same constructor, same
per-layer `num_bases` cursor with `rewind()`); its SVD initialisation — which IS the result for every
shipped "svd" config (max_iter=0) — runs on the device.  `LowRankExpV2` is the scheme-2 app the
reference never shipped: closed-form SVD fit of the (C*d) x (d*N) reshaping of the filter bank.
"""
import numpy as np
from torch import nn

from approx.layers import LowRankExpConvV1, LowRankExpConvV2, Substitution
from . import fit
from .approximater import APP, Approximater


def _log():
    try:
        from approx.utils.logger import get_logger   # only present in overlay mode (reference's logger)
        return get_logger()
    except Exception:
        import logging
        return logging.getLogger("approx")


@APP.register_module()
class LowRankExpV1(Approximater):
    _src_type = nn.Conv2d
    _tgt_type = "LowRankExpConvV1"

    def __init__(self, num_bases, max_iter, lmda_length, min_lmda, max_lmda, init_method='svd', inc_rate=1.5,
                 do_decomp=False, init_decomp=False, verbose=False, epsilon=1e-3, deploy=False):
        super().__init__(deploy=deploy)
        assert max_lmda >= min_lmda >= 0.0
        self.num_bases = num_bases
        self.curr = 0
        self.max_iter = max_iter
        # log-spaced penalty schedule rescaled into [min_lmda, max_lmda] (reference :30-31)
        ramp = np.logspace(0, inc_rate, lmda_length + 1)[1:] - 1
        self.lmda_list = ramp / ramp[-1] * (max_lmda - min_lmda) + min_lmda
        self.do_decomp = do_decomp
        self.init_decomp = init_decomp
        self.init_method = init_method
        self.verbose = verbose
        self.epsilon = epsilon

    def rewind(self):
        self.curr = 0

    def _get_tgt_args(self, src: nn.Conv2d) -> dict:
        num_base = self.num_bases[self.curr]   # one entry per selected layer, in registration order
        self.curr += 1
        return dict(in_channels=src.in_channels, out_channels=src.out_channels, num_base=num_base,
                    kernel_size=src.kernel_size, stride=src.stride, padding=src.padding, decomp=self.init_decomp)

    def _fix_substitution(self, sub: Substitution):
        sub.new_module.bias.data = sub.old_module.bias.data

    def optimize(self, sub: Substitution):
        """Reference core/low_rank_exp.py:139-184.  The starting point (`init_method`) IS the result when max_iter = 0 (every shipped
        *_svd_* config); with max_iter > 0 the alternating nuclear-norm fit the reference delegates to cvxpy runs on the device
        (csrc/altfit.cu) over the same lambda schedule, iteration cap and stopping rule."""
        if self.init_decomp:
            return
        src, tgt = sub.old_module, sub.new_module
        W = src.weight.data
        N, C, d = W.shape[:3]
        assert W.shape[3] == d, "the reference assumes square kernels (core/low_rank_exp.py:149)"
        M = tgt.num_base
        W2, weights, bases = fit.scheme1_init(W, M, self.init_method)
        log = _log()
        if self.max_iter > 0 and len(self.lmda_list) > 0:
            log.info(f"lambda list: {self.lmda_list}")
            from approx import backend as B
            for lmda in self.lmda_list:          # one call per lambda so the principal-component energy can be logged like the reference
                trace = B.lowrank1_alt_fit(W2, d, weights, bases, [float(lmda)], self.max_iter, self.epsilon)
                for i, f in enumerate(trace):
                    log.info(f"[lamda: {lmda}]({i // 2 + 1}/{self.max_iter})[{i % 2 + 1}], total error: {f}")
                log.info(f"PC Energy = {fit.pc_energy(bases.reshape(M, d, d))}")
            self.last_trace = trace
        else:
            for _ in self.lmda_list:             # the reference logs the energy once per penalty value
                log.info(f"PC Energy = {fit.pc_energy(bases.reshape(M, d, d))}")
        s_w, d_w = fit.scheme1_writeback(weights, bases, N, C, d, W.device)
        tgt.s_conv.weight.data = s_w
        tgt.d_conv.weight.data = d_w

    def _postprocess(self, sub: Substitution):
        if self.do_decomp:
            sub.new_module.decomp()


@APP.register_module()
class LowRankExpV2(Approximater):
    """Scheme 2: W[n,c,i,j] ~ sum_k v[k,c,i] * h[n,k,j]; `num_bases[i]` = K of the i-th selected conv."""
    _src_type = nn.Conv2d
    _tgt_type = "LowRankExpConvV2"

    def __init__(self, num_bases, deploy=False):
        super().__init__(deploy=deploy)
        self.num_bases = num_bases
        self.curr = 0

    def rewind(self):
        self.curr = 0

    def _get_tgt_args(self, src: nn.Conv2d) -> dict:
        k = self.num_bases[self.curr]
        self.curr += 1
        return dict(in_channels=src.in_channels, num_base=k, kernel_size=src.kernel_size, stride=src.stride,
                    padding=src.padding, out_channels=src.out_channels)

    def _fix_substitution(self, sub: Substitution):
        sub.new_module.bias.data = sub.old_module.bias.data

    def optimize(self, sub: Substitution):
        src, tgt = sub.old_module, sub.new_module
        v_w, h_w = fit.scheme2_svd(src.weight.data, tgt.num_base)
        tgt.v_conv.weight.data = v_w
        tgt.h_conv.weight.data = h_w

    def _postprocess(self, sub: Substitution):
        pass
