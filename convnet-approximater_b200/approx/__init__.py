"""B200 (sm_100a) backend for ConvNet-Approximater's hot path.

Same import path as the reference's `src/approx` package for the part it replaces:
`approx.layers` (LAYER registry, LowRankExpConvV1/V2, MSCA, CascadeConv, ParallelConv,
FixPaddingBias, Substitution) and `approx.core` (APP registry, LowRankExpV1, LowRankExpV2,
MscaRep, MscaProfile, MscaRepProfile, Dummy).  `forward()` / `optimize()` bodies call
hand-written CUDA through the C ABI in include/cna_b200.h (approx.backend); there is no CPU
fallback.

Overlay mode: set CNA_REFERENCE_SRC=<reference>/src to let every sub-package this backend does
not provide (approx.models, approx.runner, approx.hooks, approx.filters, approx.classification and
the rest of approx.utils) resolve to the reference's unchanged files, so `scripts/main.py` and the
shipped configs drive this backend as a drop-in.
"""
import os as _os

_ref = _os.environ.get("CNA_REFERENCE_SRC")
if _ref:
    _p = _os.path.join(_ref, "approx")
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)

__version__ = "0.1.0"
