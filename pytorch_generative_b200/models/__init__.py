"""Drop-in equivalents of `pytorch_generative.models` for the autoregressive-image path
(reference models/__init__.py:4,5,8,9)."""

from .base import AutoregressiveModel, GenerativeModel
from .gated_pixel_cnn import GatedPixelCNN
from .image_gpt import ImageGPT
from .pixel_cnn import PixelCNN
from .pixel_snail import PixelSNAIL

__all__ = ["AutoregressiveModel", "GenerativeModel", "GatedPixelCNN", "ImageGPT", "PixelCNN", "PixelSNAIL"]
