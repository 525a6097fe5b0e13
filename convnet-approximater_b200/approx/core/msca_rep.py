"""MSCA re-parameterisation apps (reference src/approx/core/msca_rep.py).

`MscaRep(decomp, fix)` merges the three strip cascades + identity of an MSCA block into ONE KxK
depthwise kernel (decomp=0) or into its rank-`decomp` SVD split (1: a single 1xK . Kx1 cascade,
2-4: parallel cascades); `fix=True` appends a FixPaddingBias that makes the merge exact at the top /
bottom border.  The merge and the SVD run on the device (csrc/fit.cu).
"""
from typing import Dict

from torch import nn

from approx.layers import CascadeConv, FixPaddingBias, MSCA, ParallelConv, Substitution
from . import fit
from .approximater import APP, Approximater


def _msca_args(src) -> dict:
    return dict(num_channel=src.num_channel, k1_size=src.k1_size, k_sizes=src.k_sizes)


@APP.register_module()
class MscaRep(Approximater):
    _src_type = "MSCA"
    _tgt_type = "MSCA"

    def __init__(self, decomp: int, fix: bool, deploy: bool = False):
        super().__init__(deploy=deploy)
        assert 0 <= decomp <= 4
        self.decomp = decomp
        self.fix = fix

    def _get_tgt_args(self, src: MSCA) -> dict:
        return _msca_args(src)

    def _fix_substitution(self, sub: Substitution):
        src, tgt = sub.old_module, sub.new_module
        tgt.conv0.load_state_dict(src.conv0.state_dict())
        tgt.channel_mix.load_state_dict(src.channel_mix.state_dict())
        C, K = src.num_channel, max(src.k_sizes)
        if self.decomp == 0:
            core = nn.Conv2d(C, C, K, padding=K // 2, groups=C)
        elif self.decomp == 1:
            core = CascadeConv(C, K, K // 2, True, False)
        else:
            core = ParallelConv(C, K, K // 2, self.decomp, False, False)
        # state_dict keys: sd_convs.0.* / sd_convs.1.res with fix, sd_convs.* without (same as the reference)
        tgt.sd_convs = nn.Sequential(core, FixPaddingBias(C, K // 2)) if self.fix else core

    @staticmethod
    def get_equivalent_kernel(module: ParallelConv):
        """-> weight (C,1,K,K) = delta + sum_b pad(w2_b @ w1_b); bias (C,); res (2,C,K//2) border residues."""
        return fit.msca_equivalent_kernel(module)

    def optimize(self, sub: Substitution):
        src, tgt = sub.old_module, sub.new_module
        weight, bias, res = self.get_equivalent_kernel(src.sd_convs)
        core = tgt.sd_convs[0] if self.fix else tgt.sd_convs
        if self.decomp == 0:
            core.weight.data = weight
            core.bias.data = bias
        else:
            factors, pce = fit.strip_factors(weight, self.decomp)
            cascades = [core] if self.decomp == 1 else list(core.branches)[:self.decomp]
            for cas, (c1, c2) in zip(cascades, factors):
                cas.conv1.weight.data = c1
                cas.conv2.weight.data = c2
            cascades[-1].conv2.bias.data = bias
            print(f'{pce},')   # the reference prints the mean principal-component energy (msca_rep.py:171)
        if self.fix:
            tgt.sd_convs[1].res.data = res

    def _postprocess(self, sub: Substitution):
        pass


@APP.register_module()
class MscaProfile(Approximater):
    """Swaps MSCA for MSCAProfile (same weights, profiler ranges) without changing the math."""
    _src_type = "MSCA"
    _tgt_type = "MSCAProfile"

    def __init__(self, deploy):
        super().__init__(deploy)

    def _get_tgt_args(self, src: nn.Module) -> Dict:
        return _msca_args(src)

    def _fix_substitution(self, sub: Substitution):
        sub.new_module.load_state_dict(sub.old_module.state_dict())

    def optimize(self, sub: Substitution):
        pass

    def _postprocess(self, sub: Substitution):
        pass


@APP.register_module()
class MscaRepProfile(MscaRep):
    _src_type = "MSCA"
    _tgt_type = "MSCAProfile"
