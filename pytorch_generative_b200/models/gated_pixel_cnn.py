"""Gated PixelCNN on the B200 path — API of reference models/autoregressive/gated_pixel_cnn.py:31-190.

Same module tree and state-dict keys (`_input` / `_gated_layers.{i}` with `_vstack_1xN, _vstack_Nx1, _vstack_1x1,
_link, _hstack_1xN, _hstack_residual, _hstack_skip`, `_head.{1,3}`).  The reference gets causality from plain
convolutions with extra padding followed by a front crop; here each such conv is a tap list whose offsets already
encode pad + crop (no padded rows are ever computed), contracted on the tensor cores.
"""

from torch import nn

from .. import _lib as L
from .. import nn as pg_nn
from . import base

RELU = L.ACT_RELU


class GatedPixelCNNLayer(nn.Module):
    """One two-stream layer (reference gated_pixel_cnn.py:31-130)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, mask_center=False):
        super().__init__()
        assert kernel_size % 2 == 1, "kernel_size cannot be even"
        self._in_channels, self._out_channels = in_channels, out_channels
        self._activation = pg_nn.GatedActivation()
        self._kernel_size = kernel_size
        self._padding = (kernel_size - 1) // 2
        self._mask_center = mask_center
        k, p = kernel_size, self._padding
        # vertical stack: 1xN over the row, then (N//2+1)x1 shifted down by the extra padding row
        self._vstack_1xN = pg_nn.TapConv2d(in_channels, out_channels, kernel_size=(1, k), padding=(0, p))
        self._vstack_Nx1 = pg_nn.TapConv2d(out_channels, 2 * out_channels, kernel_size=(k // 2 + 1, 1),
                                           padding=(p + 1, 0))
        self._vstack_1x1 = pg_nn.TapConv2d(in_channels, 2 * out_channels, kernel_size=1)
        self._link = pg_nn.TapConv2d(2 * out_channels, 2 * out_channels, kernel_size=1)
        # horizontal stack: 1x(N//2+1) looking left (excluding the centre when causal)
        self._hstack_1xN = pg_nn.TapConv2d(in_channels, 2 * out_channels, kernel_size=(1, k // 2 + 1),
                                           padding=(0, p + int(mask_center)))
        self._hstack_residual = pg_nn.TapConv2d(out_channels, out_channels, kernel_size=1)
        self._hstack_skip = pg_nn.TapConv2d(out_channels, out_channels, kernel_size=1)

    def forward(self, vstack_input, hstack_input):
        vstack = self._vstack_Nx1(self._vstack_1xN(vstack_input))  # TapConv2d output == the reference's [:h] crop
        link = self._link(vstack)
        vstack = self._activation(vstack + self._vstack_1x1(vstack_input))
        hstack = self._activation(link + self._hstack_1xN(hstack_input))  # == the reference's [:w] crop
        skip = self._hstack_skip(hstack)
        hstack = self._hstack_residual(hstack)
        if not self._mask_center:  # no residual on the causal layer: it would leak the current pixel
            hstack = hstack + hstack_input
        return vstack, hstack, skip


class GatedPixelCNN(base.AutoregressiveModel):
    """The Gated PixelCNN model — constructor of reference gated_pixel_cnn.py:136-183."""

    def __init__(self, in_channels=1, out_channels=1, n_gated=10, gated_channels=128, head_channels=32, sample_fn=None):
        super().__init__(sample_fn)
        self._input = GatedPixelCNNLayer(in_channels=in_channels, out_channels=gated_channels, kernel_size=7,
                                         mask_center=True)
        self._gated_layers = nn.ModuleList(
            [GatedPixelCNNLayer(in_channels=gated_channels, out_channels=gated_channels, kernel_size=3,
                                mask_center=False) for _ in range(n_gated)]
        )
        self._head = nn.Sequential(
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=gated_channels, out_channels=head_channels, kernel_size=1),
            nn.ReLU(),
            pg_nn.TapConv2d(in_channels=head_channels, out_channels=out_channels, kernel_size=1),
        )

    def forward(self, x):
        vstack, hstack, skip_connections = self._input(x, x)
        for gated_layer in self._gated_layers:
            vstack, hstack, skip = gated_layer(vstack, hstack)
            skip_connections = skip_connections + skip
        t = self._head[1](skip_connections, pre_act=RELU)
        return self._head[3](t, pre_act=RELU)
