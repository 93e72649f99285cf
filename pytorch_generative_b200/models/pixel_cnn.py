"""PixelCNN on the B200 path — API of reference models/autoregressive/pixel_cnn.py:23-110.

Same module tree and state-dict keys (`_input`, `_causal_layers.{i}._net.{1,3,5}`, `_head.{1,3}`).  The ReLUs of
the reference's `nn.Sequential`s are not separate ops here: each one is fused into the convolution that consumes it
(`pre_act`), so a block is three kernels-backed convs: 1x1, masked 3x3 (type B), 1x1.
"""

from torch import nn

from .. import _lib as L
from .. import nn as pg_nn
from . import base

RELU = L.ACT_RELU


class CausalResidualBlock(nn.Module):
    """x + net(x), net = ReLU-1x1-ReLU-causal3x3-ReLU-1x1 (reference pixel_cnn.py:23-53)."""

    def __init__(self, n_channels):
        super().__init__()
        self._net = nn.Sequential(  # container kept for the reference's parameter names; forward fuses the ReLUs
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=n_channels, out_channels=n_channels // 2, kernel_size=1),
            nn.ReLU(),
            pg_nn.CausalConv2d(mask_center=False, in_channels=n_channels // 2, out_channels=n_channels // 2,
                               kernel_size=3, padding=1),
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=n_channels // 2, out_channels=n_channels, kernel_size=1),
        )

    def forward(self, x):
        t = self._net[1](x, pre_act=RELU)
        t = self._net[3](t, pre_act=RELU)
        t = self._net[5](t, pre_act=RELU)
        return x + t


class PixelCNN(base.AutoregressiveModel):
    """The PixelCNN model — constructor of reference pixel_cnn.py:59-104."""

    def __init__(self, in_channels=1, out_channels=1, n_residual=15, residual_channels=128, head_channels=32,
                 sample_fn=None):
        super().__init__(sample_fn)
        self._input = pg_nn.CausalConv2d(mask_center=True, in_channels=in_channels,
                                         out_channels=2 * residual_channels, kernel_size=7, padding=3)
        self._causal_layers = nn.ModuleList(
            [CausalResidualBlock(n_channels=2 * residual_channels) for _ in range(n_residual)]
        )
        self._head = nn.Sequential(
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=2 * residual_channels, out_channels=head_channels, kernel_size=1),
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=head_channels, out_channels=out_channels, kernel_size=1),
        )

    def forward(self, x):
        x = self._input(x)
        for layer in self._causal_layers:
            x = x + layer(x)  # the reference's second residual (pixel_cnn.py:109 on top of :53)
        x = self._head[1](x, pre_act=RELU)
        return self._head[3](x, pre_act=RELU)


def reproduce(*args, **kwargs):
    """The recipe of this model (reference pixel_cnn.py `reproduce`); see `pytorch_generative_b200.recipes`."""
    from .. import recipes

    return recipes.reproduce_pixel_cnn(*args, **kwargs)
