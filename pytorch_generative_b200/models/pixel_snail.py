"""PixelSNAIL on the B200 path — API of reference models/autoregressive/pixel_snail.py:27-187.

Same module tree and state-dict keys (`_input`, `_pixel_snail_blocks.{i}` with `_residual.{j}.{_input_conv,
_output_conv}`, `_attention.{_q,_kv,_proj}`, `_residual_out`, `_attention_out`, `_out`; `_output.{0,1}`).  The 2x2
convolutions with pad 1 + crop are tap lists {(-1,-1),(-1,0),(0,-1),(0,0)} on the tensor-core GEMM; every ELU is
fused into the conv that consumes or produces it (`pre_act` / `post_act`); the gate uses the identity activation.
"""

import os

import torch
from torch import nn

from .. import _lib as L
from .. import nn as pg_nn
from .. import ops
from ..nn import pm
from . import base

ELU, NONE = L.ACT_ELU, L.ACT_NONE


class ResidualBlock(nn.Module):
    """x + gate(conv2x2(elu(conv2x2(elu(x))))) (reference pixel_snail.py:31-56)."""

    def __init__(self, n_channels):
        super().__init__()
        self._input_conv = pg_nn.TapConv2d(n_channels, n_channels, kernel_size=2, padding=1)
        self._output_conv = pg_nn.TapConv2d(n_channels, 2 * n_channels, kernel_size=2, padding=1)
        self._activation = pg_nn.GatedActivation(activation_fn=nn.Identity())

    def forward_pm(self, geom, x_f, x_e=None):
        """Pixel-major: x_f fp32 stream [P, C] (x_e = bf16 elu(x_f) when a producer already emitted it) -> fp32 stream.
        conv -> elu -> conv runs as two tap-loop GEMMs: the first one's epilogue writes elu(.) only, the second one's
        dgrad epilogue applies elu'."""
        _, t = pm.conv(x_f, self._input_conv.weight, self._input_conv.bias, geom, (1, 1), in_act=ELU, xa=x_e, emit=ELU,
                       emit_mode=pm.PRE_GRAD, want_main=False)
        u, _ = pm.conv(t, self._output_conv.weight, self._output_conv.bias, geom, (1, 1), in_act=ELU, xa=t)
        return pm.gated_res(u, x_f, NONE)

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

    def forward_pm(self, geom, x_f, img_b, pos_b, pad_b, c_img):
        """Pixel-major block: x_f fp32 stream [P, C]; img_b / pos_b bf16 [P, c_img] / [P, 2]; returns the block output
        (to be added to the stream)."""
        res = x_f
        for rb in self._residual:
            res = rb.forward_pm(geom, res)
        c = res.shape[1]
        a_kv = torch.cat((pos_b, pm.act_cast(res), img_b, pad_b), dim=1)  # [pos | features | image | 0-pad], bf16
        attn = self._attention.forward_pm(a_kv, geom, c + 2, c_img)       # fp32 [P, value channels]
        _, r = pm.conv(res, self._residual_out.weight, self._residual_out.bias, geom, in_act=ELU, emit=ELU,
                       emit_mode=pm.POST, want_main=False)
        _, a = pm.conv(attn, self._attention_out.weight, self._attention_out.bias, geom, in_act=ELU, emit=ELU,
                       emit_mode=pm.POST, want_main=False)
        _, out = pm.conv(r + a, self._out.weight, self._out.bias, geom, in_act=ELU, emit=ELU, emit_mode=pm.POST,
                         want_main=False)
        return out

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

    def _forward_pm(self, x):
        """The whole network on pixel-major tensors: NCHW only at the image and at the logits."""
        n, c_img, h, w = x.shape
        geom = pm.Geom(n, h, w)
        self._input.weight.data *= self._input.mask  # CausalConv2d's in-place masking (reference nn/convolution.py:42)
        kh, kw = self._input.weight.shape[2:]
        s = pm.small_conv(x, self._input.weight, self._input.bias, (kh // 2, kw // 2))  # fp32 stream [P, C]
        blk0 = self._pixel_snail_blocks[0]
        pos_b = ops.nchw_to_pm(blk0._positions(x.shape, x.device), torch.bfloat16)
        img_b = ops.nchw_to_pm(x, torch.bfloat16)
        width = ops.round_up(2 + s.shape[1] + c_img, 8)
        pad_b = torch.zeros(n * h * w, width - (2 + s.shape[1] + c_img), dtype=torch.bfloat16, device=x.device)
        for block in self._pixel_snail_blocks:
            s = s + block.forward_pm(geom, s, img_b, pos_b, pad_b, c_img)
        t, _ = pm.conv(s, self._output[0].weight, self._output[0].bias, geom)
        logits, _ = pm.conv(t, self._output[1].weight, self._output[1].bias, geom, out_f32=True)
        return pm.from_pm(logits, geom, self._output[1].weight.shape[0])

    def _pm_ok(self, x):
        c = self._input.weight.shape[0]
        kh, kw = self._input.weight.shape[2:]
        return (x.is_cuda and os.environ.get("PG_NO_PM_STACK") != "1" and x.shape[1] * kh * kw <= 160
                and pm.supported(x.shape[2], x.shape[3], (c,)))

    def forward(self, x):
        if self._pm_ok(x):
            return self._forward_pm(x)
        input_img = x
        x = self._input(x)
        for block in self._pixel_snail_blocks:
            x = x + block(x, input_img)
        return self._output[1](self._output[0](x))


def reproduce(*args, **kwargs):
    """The recipe of this model (reference pixel_snail.py `reproduce`); see `pytorch_generative_b200.recipes`."""
    from .. import recipes

    return recipes.reproduce_pixel_snail(*args, **kwargs)
