"""Models this backend adds next to the reference's (AlexNet, MSCAN stay the reference's own files).

In overlay mode (CNA_REFERENCE_SRC set) the reference's approx.models package is merged in, so
`build_model(dict(type='AlexNet', ...))` keeps working and `VGG16` is registered in the same MODEL
registry."""
import os as _os

_ref = _os.environ.get("CNA_REFERENCE_SRC")
if _ref:
    _p = _os.path.join(_ref, "approx", "models")
    if _os.path.isdir(_p) and _p not in __path__:
        __path__.append(_p)
    from .switchable import SwitchableModel, MODEL, build_model  # noqa: F401  (reference files)
    from .alexnet import AlexNet  # noqa: F401
    from .mscan import MSCAN_Classifier  # noqa: F401

from .vgg import VGG16_CFG, VGG16Scheme2Trunk  # noqa: E402,F401
