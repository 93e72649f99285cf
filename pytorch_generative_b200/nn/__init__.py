"""Drop-in equivalents of `pytorch_generative.nn` for the classes on the hot path (reference nn/__init__.py:3-13)."""

from .modules import (
    CausalAttention,
    CausalConv2d,
    GatedActivation,
    NCHWLayerNorm,
    image_positional_encoding,
)

__all__ = ["CausalAttention", "CausalConv2d", "GatedActivation", "NCHWLayerNorm", "image_positional_encoding"]
