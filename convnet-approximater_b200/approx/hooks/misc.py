"""Small hooks: post-run rank-1 split of the scheme-1 bases, checkpoint save/load at any stage."""
import os

from .hook import HOOK, Hook, STAGES


@HOOK.register_module()
class LowRankExpV1Decomp(Hook):
    """after_run: `decomp()` on every replaced layer (reference hooks/low_rank_exp_v1_decomp.py:5-13); the batched rank-1 SVD
    runs on the device (cna_svd_small_batched)."""

    def __init__(self, runner, priority):
        super().__init__(runner, priority)

    def after_run(self):
        from approx.layers import LowRankExpConvV1
        for m in self.runner.model.switchable_modules():
            assert isinstance(m, LowRankExpConvV1), type(m).__name__
            m.decomp()


@HOOK.register_module()
class CkptHook(Hook):
    """ckpt_cfg = {stage: dict(action='save'|'load', path=...)} (reference hooks/checkpoint.py:8-47).  Checkpoints use the
    reference's format — torch.save({'state_dict': ...}), non-strict load (utils/serialize.py:11-34) — so files are
    interchangeable between the two backends."""

    def __init__(self, runner, priority, ckpt_cfg):
        super().__init__(runner, priority)
        self.plan = {}
        for stage in STAGES:
            cur = ckpt_cfg.get(stage) if hasattr(ckpt_cfg, "get") else None
            if cur is not None:
                assert cur['action'] in ('save', 'load'), cur['action']
                if cur['action'] == 'load':
                    assert os.path.isfile(cur['path']), cur['path']
                else:
                    assert os.path.isdir(os.path.dirname(cur['path']) or "."), cur['path']
            self.plan[stage] = cur

    def _apply(self, stage):
        cur = self.plan.get(stage)
        if cur is None:
            return
        try:
            from approx.utils.serialize import load_model, save_model       # the reference's own (de)serialiser in overlay mode
        except Exception:  # noqa: BLE001
            import torch

            def save_model(model, path):
                torch.save({'state_dict': model.state_dict()}, path)

            def load_model(model, path):
                sd = torch.load(path, map_location='cpu')
                model.load_state_dict(sd.get('state_dict', sd), strict=False)
        (save_model if cur['action'] == 'save' else load_model)(self.runner.model, cur['path'])


for _stage in STAGES:
    setattr(CkptHook, _stage, (lambda st: lambda self: self._apply(st))(_stage))
