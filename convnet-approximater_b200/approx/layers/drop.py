from torch import nn


class DropPath(nn.Module):
    """Stochastic depth: zero whole samples with probability `drop_prob` while training (identity in eval).
    Needed only because the reference's MSCAN model imports it from `approx.layers` (models/mscan.py)."""

    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if not self.training or self.drop_prob == 0.:
            return x
        keep = 1. - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if self.scale_by_keep and keep > 0.:
            mask.div_(keep)
        return x * mask
