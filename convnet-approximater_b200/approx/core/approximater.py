"""Approximater base class and the APP registry.

Interface of the reference's src/approx/core/approximater.py:9-68, unchanged: an app names a source
layer type and a target layer type, `initialize(src)` builds the target and wraps both in a
`Substitution`, `optimize(sub)` fits the target's weights from the source's, `postprocess(sub)`
returns the finished target layer.  `Runner.run` (reference runner/runner.py:56-76) drives those three.
"""
from abc import ABCMeta, abstractmethod
from typing import Dict, Union

from torch import nn

from approx.layers import LAYER, Substitution, build_layer
from approx.utils.registry import Registry, build_from_cfg


def _resolve(t):
    return t if isinstance(t, type) else LAYER.get(t)


class Approximater(metaclass=ABCMeta):
    _src_type = ""   # class, or a LAYER registry name
    _tgt_type = ""

    def __init__(self, deploy=False):
        self.deploy = deploy

    @property
    def src_type(self):
        return _resolve(self._src_type)

    @property
    def tgt_type(self):
        return _resolve(self._tgt_type)

    @abstractmethod
    def _get_tgt_args(self, src: nn.Module) -> Dict:
        """Constructor arguments of the target layer that replaces `src`."""

    @abstractmethod
    def _fix_substitution(self, sub: Substitution):
        """Copy what carries over unchanged from `sub.old` into `sub.new` (biases, untouched sub-layers)."""

    @abstractmethod
    def optimize(self, sub: Substitution):
        """Fit `sub.new`'s weights to `sub.old`'s."""

    @abstractmethod
    def _postprocess(self, sub: Substitution):
        pass

    def initialize(self, src: nn.Module) -> Union[Substitution, nn.Module]:
        assert isinstance(src, self.src_type)
        tgt = build_layer(dict(type=self.tgt_type, **self._get_tgt_args(src)))
        sub = Substitution(src, tgt)
        self._fix_substitution(sub)
        return sub.new_module if self.deploy else sub   # deploy: bare target layers, weights come from a checkpoint

    def postprocess(self, sub: Substitution) -> nn.Module:
        self._postprocess(sub)
        return sub.new_module

    def rewind(self):
        pass


APP = Registry()


def build_app(cfg: dict, **kwargs) -> Approximater:
    return build_from_cfg(cfg, APP, **kwargs)
