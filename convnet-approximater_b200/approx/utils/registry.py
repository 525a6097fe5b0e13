"""Name -> class registries and config-driven construction.

Mirrors the behaviour of the reference's src/approx/utils/registry.py:1-50 (same class and
function names, same assertion / TypeError behaviour) so the shipped configs
(`dict(type='MscaRep', decomp=1, fix=True)`) build this backend's classes.
"""


class Registry:
    def __init__(self):
        self._modules = {}

    def __contains__(self, name):
        return name in self._modules

    def _add(self, cls, name):
        key = cls.__name__ if name is None else name
        assert key not in self._modules, f"{key} is already registered."
        self._modules[key] = cls
        return cls

    def register_module(self, name=None, module=None):
        """Usable as `@REG.register_module()` or `REG.register_module(module=cls)`."""
        if module is not None:
            return self._add(module, name)
        return lambda cls: self._add(cls, name)

    def get(self, name):
        assert name in self._modules, f"{name} is not registered."
        return self._modules[name]


def build_from_cfg(cfg, registry, **kwargs):
    """str -> registry[str](**kwargs); dict -> pop 'type' and construct; list -> nn.Sequential; None -> None."""
    if cfg is None:
        return None
    if isinstance(cfg, str):
        return registry.get(cfg)(**kwargs)
    if isinstance(cfg, list):
        from torch import nn
        return nn.Sequential([build_from_cfg(c, registry, **kwargs) for c in cfg])
    if not isinstance(cfg, dict):
        raise TypeError(f"type {type(cfg)} not support")
    args = dict(cfg)
    args.update(kwargs)
    kind = args.pop("type")
    cls = kind if isinstance(kind, type) else registry.get(kind)
    try:
        return cls(**args)
    except TypeError as e:
        msg = str(e)
        raise TypeError(msg if "<class" in msg else f"{cls}.{msg}")
