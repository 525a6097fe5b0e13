import os as _os

_ref = _os.environ.get("CNA_REFERENCE_SRC")
if _ref:
    _p = _os.path.join(_ref, "approx", "utils")
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)

from .registry import Registry, build_from_cfg  # noqa: E402,F401
from .general import check_file, to_2tuple, is_method_overridden, parse_path  # noqa: E402,F401

if _ref:
    # overlay mode: the reference's approx.utils re-exports these at package level
    # (src/approx/utils/__init__.py:1-5 of the reference); they resolve to its own files.
    from .config import get_cfg, init_cfg, update_cfg, save_cfg, print_cfg  # noqa: F401
    from .serialize import load_model, save_model  # noqa: F401
    from .random import random_seed  # noqa: F401
    from .distributed import distribute_bn, unwrap_model, reduce_tensor  # noqa: F401
