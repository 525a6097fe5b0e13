from torch import nn

from .substituton import LAYER


@LAYER.register_module()
class SimpleConv(nn.Module):
    """A conv wrapped in a module that exposes its hyper-parameters as read-only properties.
    Registered under the reference's name (layers/simple_conv.py:6-49); it is a test source layer,
    not part of the accelerated path, so it simply owns a stock `nn.Conv2d`."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self._hp = dict(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                        stride=stride, padding=padding)

    def forward(self, x):
        return self.conv(x)

    in_channels = property(lambda self: self._hp["in_channels"])
    out_channels = property(lambda self: self._hp["out_channels"])
    kernel_size = property(lambda self: self._hp["kernel_size"])
    stride = property(lambda self: self._hp["stride"])
    padding = property(lambda self: self._hp["padding"])
    weight = property(lambda self: self.conv.weight)
    bias = property(lambda self: self.conv.bias)
