"""PixelSNAIL on the B200 path — API of reference models/autoregressive/pixel_snail.py:27-187.

Same module tree and state-dict keys (`_input`, `_pixel_snail_blocks.{i}` with `_residual.{j}.{_input_conv,
_output_conv}`, `_attention.{_q,_kv,_proj}`, `_residual_out`, `_attention_out`, `_out`; `_output.{0,1}`).  The 2x2
convolutions with pad 1 + crop are tap lists {(-1,-1),(-1,0),(0,-1),(0,0)} on the tensor-core GEMM; every ELU is
fused into the conv that consumes or produces it (`pre_act` / `post_act`); the gate uses the identity activation.
"""

import torch
from torch import nn

from .. import _lib as L
from .. import nn as pg_nn
from . import base

ELU = L.ACT_ELU


class ResidualBlock(nn.Module):
    """x + gate(conv2x2(elu(conv2x2(elu(x))))) (reference pixel_snail.py:31-56)."""

    def __init__(self, n_channels):
        super().__init__()
        self._input_conv = pg_nn.TapConv2d(n_channels, n_channels, kernel_size=2, padding=1)
        self._output_conv = pg_nn.TapConv2d(n_channels, 2 * n_channels, kernel_size=2, padding=1)
        self._activation = pg_nn.GatedActivation(activation_fn=nn.Identity())

    def forward(self, x):
        out = self._input_conv(x, pre_act=ELU)       # conv(elu(x)), cropped to h x w
        out = self._output_conv(out, pre_act=ELU)    # conv(elu(.)), cropped
        return x + self._activation(out)


class PixelSNAILBlock(nn.Module):
    """Residual blocks + causal attention over (position, features | image) (reference pixel_snail.py:59-119)."""

    def __init__(self, n_channels, input_img_channels=1, n_residual_blocks=2, attention_key_channels=4,
                 attention_value_channels=32):
        super().__init__()

        def conv(in_channels):
            return pg_nn.TapConv2d(in_channels, out_channels=n_channels, kernel_size=1)

        self._residual = nn.Sequential(*[ResidualBlock(n_channels) for _ in range(n_residual_blocks)])
        self._attention = pg_nn.CausalAttention(in_channels=n_channels + 2, embed_channels=attention_key_channels,
                                                out_channels=attention_value_channels, mask_center=True,
                                                extra_input_channels=input_img_channels)
        self._residual_out = conv(n_channels)
        self._attention_out = conv(attention_value_channels)
        self._out = conv(n_channels)
        self._pos_cache = {}

    def _positions(self, shape, device):
        key = (tuple(shape), str(device))
        if key not in self._pos_cache:  # same values as the reference's image_positional_encoding, kept on device
            self._pos_cache[key] = pg_nn.image_positional_encoding(tuple(shape)).to(device)
        return self._pos_cache[key]

    def forward(self, x, input_img):
        res = self._residual(x)
        pos = self._positions(input_img.shape, res.device)
        attn = self._attention(torch.cat((pos, res), dim=1), input_img)
        res = self._residual_out(res, pre_act=ELU, post_act=ELU)
        attn = self._attention_out(attn, pre_act=ELU, post_act=ELU)
        return self._out(res + attn, pre_act=ELU, post_act=ELU)


class PixelSNAIL(base.AutoregressiveModel):
    """The PixelSNAIL model — constructor of reference pixel_snail.py:130-180."""

    # the positional encoding is a function of the image height (arange(-.5, .5, 1/h)), so a pixel cannot be evaluated
    # on a truncated canvas: sample() runs the full forward per pixel, exactly like the reference
    _row_truncated_sampling = False

    def __init__(self, in_channels=1, out_channels=1, n_channels=64, n_pixel_snail_blocks=8, n_residual_blocks=2,
                 attention_key_channels=4, attention_value_channels=32, sample_fn=None):
        super().__init__(sample_fn)
        self._input = pg_nn.CausalConv2d(mask_center=True, in_channels=in_channels, out_channels=n_channels,
                                         kernel_size=3, padding=1)
        self._pixel_snail_blocks = nn.ModuleList(
            [PixelSNAILBlock(n_channels=n_channels, input_img_channels=in_channels,
                             n_residual_blocks=n_residual_blocks, attention_key_channels=attention_key_channels,
                             attention_value_channels=attention_value_channels)
             for _ in range(n_pixel_snail_blocks)]
        )
        self._output = nn.Sequential(
            pg_nn.TapConv2d(in_channels=n_channels, out_channels=n_channels // 2, kernel_size=1),
            pg_nn.TapConv2d(in_channels=n_channels // 2, out_channels=out_channels, kernel_size=1),
        )

    def forward(self, x):
        input_img = x
        x = self._input(x)
        for block in self._pixel_snail_blocks:
            x = x + block(x, input_img)
        return self._output[1](self._output[0](x))
