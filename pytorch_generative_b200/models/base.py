"""Base classes of the drop-in models (API of reference pytorch_generative/models/base.py:28-120).

Behavioural contract kept from the reference:
  * `__call__` records the image shape of the first 4-D input in buffers `_c, _h, _w` (int64 scalars that are
    part of the state dict) — base.py:41-46, 55-61;
  * `load_state_dict` registers those buffers first when the checkpoint has them — base.py:48-53;
  * `sample(n_samples=None, conditioned_on=None)` walks the image in raster order, draws all channels of one
    pixel from `sample_fn(logits[:, :, r, c])` and only overwrites entries < 0 — base.py:97-120;
  * `device` property — base.py:63-65.
"""

import abc
import warnings

import torch
from torch import nn


def _bernoulli_from_logits(logits):
    """Default `sample_fn` (reference base.py:9-10): one Bernoulli draw per logit."""
    return torch.bernoulli(torch.sigmoid(logits))


class GenerativeModel(abc.ABC, nn.Module):
    """Shape-tracking nn.Module base."""

    def __call__(self, x, *args, **kwargs):
        if getattr(self, "_c", None) is None and x.dim() == 4:
            self._register_shape(*x.shape[1:])
        return super().__call__(x, *args, **kwargs)

    def load_state_dict(self, state_dict, strict=True):
        if "_c" in state_dict and not getattr(self, "_c", None):
            self._register_shape(state_dict["_c"], state_dict["_h"], state_dict["_w"])
        return super().load_state_dict(state_dict, strict)

    def _register_shape(self, c, h, w):
        as_t = lambda v: v if torch.is_tensor(v) else torch.tensor(v)
        self.register_buffer("_c", as_t(c))
        self.register_buffer("_h", as_t(h))
        self.register_buffer("_w", as_t(w))

    @property
    def device(self):
        return next(self.parameters()).device

    # Per-instance runtime caches (captured CUDA graphs of the samplers, line buffers, bf16 weight arenas keyed on the
    # parameters' version counters, the data-parallel bucket hook) are rebuilt on demand and must not travel with a
    # pickled or deep-copied model: a CUDA graph cannot be copied, and a copy must not replay the original's buffers.
    _RUNTIME_CACHES = ("_sample_graphs", "_samplers", "_pixel_states", "_wcache", "_grad_bucket_hook")

    def __getstate__(self):
        state = self.__dict__.copy()
        for key in self._RUNTIME_CACHES:
            state.pop(key, None)
        return state

    @abc.abstractmethod
    def sample(self, n_samples):
        ...


class AutoregressiveModel(GenerativeModel):
    """Adds raster-scan ancestral sampling on top of `forward`."""

    def __init__(self, sample_fn=None):
        super().__init__()
        self._sample_fn = sample_fn or _bernoulli_from_logits

    def _start_canvas(self, n_samples, conditioned_on):
        assert (
            n_samples is not None or conditioned_on is not None
        ), 'Must provided one, and only one, of "n_samples" or "conditioned_on"'
        if conditioned_on is not None:
            return conditioned_on.clone()
        shape = (n_samples, int(self._c), int(self._h), int(self._w))
        return torch.full(shape, -1.0, device=self.device)

    # Models whose forward is exactly row-causal (the logits of image row r depend on rows <= r only, and every
    # kernel computes an output row from the same operands in the same order whatever the image height) can evaluate
    # a pixel on the top (r + 1) rows of the canvas: bit-identical logits for roughly half the work on average.
    _row_truncated_sampling = True

    # sample() replays one CUDA graph per distinct forward shape (one per image row when row-truncated, one in total
    # otherwise): a per-pixel forward of a small batch is ~700 tiny launches, i.e. host-bound when launched eagerly.
    _sample_with_graphs = False  # opt-in: capture costs ~1 s per shape, worth it only for repeated sampling

    def _pixel_logits_fn(self, canvas, rows):
        """Returns a callable computing forward(canvas[:, :, :rows]) (eager, or a captured-graph replay)."""
        n, c, _, w = canvas.shape
        if not (self._sample_with_graphs and canvas.is_cuda):
            return lambda: self.forward(canvas[:, :, :rows])
        cache = self.__dict__.setdefault("_sample_graphs", {})
        key = (n, c, rows, w, str(canvas.device))
        if key not in cache:
            static_in = torch.empty(n, c, rows, w, dtype=canvas.dtype, device=canvas.device)
            static_in.copy_(canvas[:, :, :rows])
            try:
                self.forward(static_in)  # warm-up outside capture (lazy one-time initialisation in the kernels' host code)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    static_out = self.forward(static_in)
                cache[key] = (graph, static_in, static_out)
            except RuntimeError as exc:  # capture not possible in this context: launch eagerly (same kernels)
                torch.cuda.synchronize()
                cache[key] = None
                warnings.warn(f"{type(self).__name__}.sample(): CUDA-graph capture failed, launching the forward eagerly: "
                              f"{exc!r}", RuntimeWarning)
        entry = cache[key]
        if entry is None:
            return lambda: self.forward(canvas[:, :, :rows])
        graph, static_in, static_out = entry

        def run():
            static_in.copy_(canvas[:, :, :rows])
            graph.replay()
            return static_out

        return run

    @torch.no_grad()
    def sample(self, n_samples=None, conditioned_on=None):
        """Generates samples pixel by pixel; entries of `conditioned_on` that are >= 0 are kept."""
        canvas = self._start_canvas(n_samples, conditioned_on)
        n, c, h, w = canvas.shape
        for row in range(h):
            logits_fn = self._pixel_logits_fn(canvas, row + 1 if self._row_truncated_sampling else h)
            for col in range(w):
                logits = logits_fn()[:, :, row, col]
                drawn = self._sample_fn(logits).view(n, c)
                current = canvas[:, :, row, col]
                canvas[:, :, row, col] = torch.where(current < 0, drawn, current)
        return canvas
