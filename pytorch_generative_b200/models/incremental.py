"""Line-buffer (incremental) sampling for the convolutional models.

`AutoregressiveModel.sample` (reference models/base.py:97-120) runs one full forward per pixel.  Every model on the path is
exactly causal, so the logits of pixel p only need, per layer, the activations of the pixels its taps reach — which are
pixels generated earlier.  `PixelStepper` keeps one activation cache `[n, H*W + 1, C]` (bf16, the extra row stays zero =
the convolution's zero padding) per convolution input and evaluates the stack at ONE position per image:

    gather   the taps of position p from the layer's cache        (index tables built once, position read on the device)
    linear   [n, taps * C] x W^T on the skinny GEMM (`pg_gemm_bf16` impl 2: weights streamed once, all SMs), with the
             bias / activation / residual epilogues of training
    write    the layer's output row into the next layer's cache

A model describes its per-pixel program in `_pixel_program(stepper, state)`; the whole program is captured in one CUDA
graph whose position lives in device memory and is replayed H*W times, like ImageGPT's KV-cached sampler.  The raster
order, the `sample_fn` hook and the "only entries < 0 are overwritten" rule are the base class's.
"""

import warnings

import torch

from .. import _lib as L
from .. import ops

BF16, F32 = torch.bfloat16, torch.float32
MAX_ROWS = 32  # the skinny GEMM handles up to 32 rows (= images sampled at once)


def live_taps(mask2d, pad_h, pad_w):
    """[(i, j, dy, dx)] of the unmasked kernel positions of a CausalConv2d (row-major, like the weight)."""
    kh, kw = mask2d.shape
    return [(i, j, i - pad_h, j - pad_w) for i in range(kh) for j in range(kw) if float(mask2d[i, j]) != 0.0]


def pack_taps(weight, taps, cin_p):
    """[Cout, Cin, kh, kw] fp32 -> [Cout, T * cin_p] bf16 with the columns of tap t at [t * cin_p, ...)."""
    cout, cin = weight.shape[:2]
    out = torch.zeros(cout, len(taps), cin_p, dtype=F32, device=weight.device)
    for t, (i, j, _, _) in enumerate(taps):
        out[:, t, :cin] = weight.detach()[:, :, i, j]
    return ops.to_bf16(out.reshape(cout, len(taps) * cin_p))


class PixelStepper:
    """Caches, tap tables and the per-pixel primitives for a batch of `n` images of `h x w` pixels."""

    def __init__(self, n, h, w, device):
        self.n, self.h, self.w, self.S, self.device = n, h, w, h * w, device
        self.pos = torch.zeros(1, dtype=torch.int64, device=device)     # current position (device side: graph-replayable)
        self.pos32 = torch.zeros(1, dtype=torch.int32, device=device)   # the same for pg_attn_decode
        self._tables = {}

    def cache(self, channels):
        return torch.zeros(self.n, self.S + 1, channels, dtype=BF16, device=self.device)

    def table(self, offsets):
        """[S, T] flat indices of position p's taps (index S = the zero row for taps outside the image)."""
        key = tuple(offsets)
        if key not in self._tables:
            rows = torch.arange(self.h).view(self.h, 1, 1)
            cols = torch.arange(self.w).view(1, self.w, 1)
            dy = torch.tensor([o[0] for o in offsets]).view(1, 1, -1)
            dx = torch.tensor([o[1] for o in offsets]).view(1, 1, -1)
            r, c = rows + dy, cols + dx
            ok = (r >= 0) & (r < self.h) & (c >= 0) & (c < self.w)
            idx = torch.where(ok, r * self.w + c, torch.full_like(r * self.w + c, self.S))
            self._tables[key] = idx.reshape(self.S, len(offsets)).to(self.device)
        return self._tables[key]

    def gather(self, cache, offsets):
        """[n, T * C] bf16: the cache rows under position p's taps, tap-major."""
        idx = self.table(offsets).index_select(0, self.pos)[0]
        return cache.index_select(1, idx).reshape(self.n, -1)

    def write(self, cache, value):
        """cache[:, p] = value ([n, C]; cast to bf16)."""
        cache.index_copy_(1, self.pos, value.to(BF16).unsqueeze(1))

    @staticmethod
    def act(x, act):
        """bf16(act(x)) of an [n, C] row block (C % 8 == 0)."""
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
        L.act_cast(x.contiguous(), act, out)
        return out

    @staticmethod
    def linear(a, w, bias, *, act=L.ACT_NONE, res0=None, res1=None, f32=False):
        """a [n, K] bf16 x w [Cout, K]^T (+ bias, residuals) -> bf16(act(.)) or fp32, on the skinny GEMM."""
        ob, _, of = ops.linear_fwd(a, w, bias, act=act, res0=res0, res1=res1, want_bf16=not f32, want_f32=f32, skinny=True)
        return of if f32 else ob


class IncrementalSamplingMixin:
    """`sample()` through a model-specific per-pixel program (`_build_pixel_state`, `_pixel_program`)."""

    _incremental_sampling = True

    def _incremental_ok(self, canvas):
        return self._incremental_sampling and canvas.is_cuda and canvas.shape[0] <= MAX_ROWS

    @torch.no_grad()
    def sample(self, n_samples=None, conditioned_on=None):
        canvas = self._start_canvas(n_samples, conditioned_on)
        if not self._incremental_ok(canvas):
            return super().sample(conditioned_on=canvas)
        n, c, h, w = canvas.shape
        cache = self.__dict__.setdefault("_pixel_states", {})
        key = (n, c, h, w, str(canvas.device))
        st = cache.get(key)
        if st is None:
            st = cache[key] = dict(stepper=PixelStepper(n, h, w, canvas.device), graph=None)
            st.update(self._build_pixel_state(st["stepper"], c))
        self._refresh_pixel_weights(st)          # the weights may have been trained since the last call
        sp = st["stepper"]
        for buf in st["caches"]:
            buf.zero_()
        if st["graph"] is None:
            sp.pos.zero_()
            sp.pos32.zero_()
            try:
                self._pixel_program(sp, st)      # warm-up outside capture
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    st["logits"] = self._pixel_program(sp, st)
                st["graph"] = graph
            except RuntimeError as exc:
                torch.cuda.synchronize()
                st["graph"] = False
                warnings.warn(f"{type(self).__name__}.sample(): CUDA-graph capture of the per-pixel program failed, "
                              f"launching it eagerly: {exc!r}", RuntimeWarning)
            for buf in st["caches"]:
                buf.zero_()
        image = st["image"]                       # [n, S + 1, c_p] bf16: the pixels generated so far
        for row in range(h):
            for col in range(w):
                p = row * w + col
                sp.pos.fill_(p)
                sp.pos32.fill_(p)
                self._before_pixel(sp, st, canvas, row, col)
                if st["graph"]:
                    st["graph"].replay()
                    logits = st["logits"]
                else:
                    logits = self._pixel_program(sp, st)
                drawn = self._sample_fn(logits[:, :c]).view(n, c)
                current = canvas[:, :, row, col]
                new = torch.where(current < 0, drawn, current)
                canvas[:, :, row, col] = new
                image[:, p, :c] = new.to(BF16)
        return canvas

    def _before_pixel(self, sp, st, canvas, row, col):
        """Hook: work that must see the previous pixel's final value (PixelSNAIL's key / value fix-up)."""

    def _refresh_pixel_weights(self, st):
        new = self._pack_pixel_weights()
        for k, v in new.items():
            if k in st["weights"]:
                st["weights"][k].copy_(v)   # in place: a captured graph keeps reading the same buffers
            else:
                st["weights"][k] = v
