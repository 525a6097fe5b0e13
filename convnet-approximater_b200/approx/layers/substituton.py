"""`Substitution` (old/new layer switch) and the LAYER registry.

Module name and public names follow the reference's src/approx/layers/substituton.py:15-52
(the file name's spelling included) so `from approx.layers.substituton import LAYER` keeps working.
"""
from torch import nn

from approx.utils.registry import Registry, build_from_cfg


class Substitution(nn.Module):
    """Holds the source layer (`old`) and its approximation (`new`); `use_old` picks the active one.

    `cache` is the scratch dict the reconstruction hooks write teacher outputs / per-sample norms into
    (reference: hooks/finetune.py:76-82)."""

    def __init__(self, old_module: nn.Module, new_module: nn.Module, use_old: bool = True):
        super().__init__()
        self.old = old_module
        self.new = new_module
        self.use_old = use_old
        self.cache = {}

    @property
    def old_module(self) -> nn.Module:
        return self.old

    @property
    def new_module(self) -> nn.Module:
        return self.new

    def switch_new(self, remove_old=True):
        self.use_old = False
        if remove_old:
            del self.old

    def switch_old(self, remove_new=False):
        self.use_old = True
        if remove_new:
            del self.new

    def forward(self, x):
        return (self.old if self.use_old else self.new)(x)


LAYER = Registry()


def build_layer(cfg: dict, **kwargs) -> nn.Module:
    return build_from_cfg(cfg, LAYER, **kwargs)
