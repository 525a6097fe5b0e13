"""Runner hooks on the B200 backend (reference package: src/approx/hooks).

Same registry (`HOOK`), base class, stage names and constructor signatures as the reference, so the shipped configs'
`hooks = [dict(type='InferenceTimeHook', ...), dict(type='L2Reconstruct', ...), ...]` entries build unchanged and the
reference's own `Runner` (runner/runner.py:48-52, 100-104) drives them.  What differs is what they measure with and train
through: CUDA events + this library's launch counter instead of host timers only, this library's loss / optimizer kernels
instead of torch ops, a synthetic loader when no dataset library is installed (timm is not a dependency of this backend).
"""
from .hook import build_hook, Hook, HOOK
from .priority import Priority, get_priority
from .measure import InferenceTimeHook, Fps, ModelAnalysis, ClassEvalHook
from .finetune import L2Reconstruct
from .misc import LowRankExpV1Decomp, CkptHook

__all__ = ["build_hook", "Hook", "HOOK", "Priority", "get_priority", "InferenceTimeHook", "Fps", "ModelAnalysis", "ClassEvalHook",
           "L2Reconstruct", "LowRankExpV1Decomp", "CkptHook"]
