"""Dispatch to the fused MSCA kernels (csrc/msca.cu) when a configuration is supported.

Each function returns None when the fused kernel does not apply; callers then run the
generic depthwise kernels (still CUDA, just more launches).
"""


def parallel_forward(par, x):
    return None


def conv0_sd_forward(msca, x):
    return None
