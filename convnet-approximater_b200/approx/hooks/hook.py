"""Hook base class and the HOOK registry (interface of the reference's src/approx/hooks/hook.py:5-44)."""
from approx.utils.registry import Registry, build_from_cfg
from .priority import get_priority

STAGES = ('before_run', 'after_register', 'after_initialize', 'after_optimize', 'after_run')


class Hook:
    """A hook is called by the Runner at each of `stages`; lower priority values run first."""
    stages = STAGES

    def __init__(self, runner, priority):
        self.runner = runner
        self._priority = get_priority(priority)

    @property
    def name(self):
        return type(self).__name__

    @property
    def priority(self):
        return self._priority


def _noop(self):
    return None


for _stage in STAGES:          # every stage defaults to a no-op; Runner.hook_info lists the overridden ones
    setattr(Hook, _stage, _noop)

HOOK = Registry()


def build_hook(cfg: dict, **kwargs) -> Hook:
    return build_from_cfg(cfg, HOOK, **kwargs)
