from typing import Dict

from torch import nn

from approx.layers import Substitution
from .approximater import APP, Approximater


@APP.register_module()
class Dummy(Approximater):
    """No-op app (reference core/dummy.py:8-26): lets hooks run on an unmodified model."""
    _src_type = "DummyLayer"
    _tgt_type = "DummyLayer"

    def __init__(self, deploy):
        super().__init__(deploy)

    def _get_tgt_args(self, src: nn.Module) -> Dict:
        return {}

    def _fix_substitution(self, sub: Substitution):
        pass

    def optimize(self, sub: Substitution):
        pass

    def _postprocess(self, sub: Substitution):
        pass
