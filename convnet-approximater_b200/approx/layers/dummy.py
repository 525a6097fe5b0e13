from torch import nn

from .substituton import LAYER


@LAYER.register_module()
class DummyLayer(nn.Module):
    """Never instantiated by a model: the `Dummy` app targets it so hooks can run on an unmodified model."""
