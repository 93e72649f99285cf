"""Overlay onto an importable reference package: `install()` rebinds the hot-path names of
`pytorch_generative` to the B200 classes so that the reference's own `train.py` / `reproduce()` / `Trainer` drive them
unmodified; every other model of the reference keeps running on its own code.

The reference's `train.py:10-24` dereferences 13 model modules at import, so shadowing the whole package is not an
option; rebinding is.  `reproduce()` of each recipe constructs its model as `models.<Name>(...)`
(e.g. image_gpt.py:140-148), i.e. through the attribute this function replaces.
"""

import importlib

_NN_NAMES = ("CausalConv2d", "GatedActivation", "NCHWLayerNorm", "CausalAttention", "LinearCausalAttention",
             "image_positional_encoding")
_MODEL_NAMES = {"PixelCNN": "pixel_cnn", "GatedPixelCNN": "gated_pixel_cnn", "PixelSNAIL": "pixel_snail",
                "ImageGPT": "image_gpt"}
_saved = {}


def install():
    """Rebinds pytorch_generative.nn.* / pytorch_generative.models.* (hot-path names only).  Returns the names bound."""
    import pytorch_generative as ref  # the reference must be importable (pip-installed or on sys.path)

    from . import models as our_models
    from . import nn as our_nn

    bound = []

    def bind(obj, name, value):
        _saved.setdefault((obj, name), getattr(obj, name))
        setattr(obj, name, value)
        bound.append(f"{obj.__name__}.{name}")

    for name in _NN_NAMES:
        bind(ref.nn, name, getattr(our_nn, name))
    for cls, mod in _MODEL_NAMES.items():
        bind(ref.models, cls, getattr(our_models, cls))
        bind(importlib.import_module(f"pytorch_generative.models.autoregressive.{mod}"), cls, getattr(our_models, cls))
    return bound


def uninstall():
    """Restores every name `install()` replaced."""
    for (obj, name), value in _saved.items():
        setattr(obj, name, value)
    _saved.clear()
