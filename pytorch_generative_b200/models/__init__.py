"""Drop-in equivalents of `pytorch_generative.models` for the autoregressive-image path
(reference models/__init__.py:4,5,8,9)."""

from .base import AutoregressiveModel, GenerativeModel
from .image_gpt import ImageGPT

__all__ = ["AutoregressiveModel", "GenerativeModel", "ImageGPT"]
