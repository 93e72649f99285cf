"""Pixel-major functional ops with autograd: the building blocks of the fused conv-model stacks.

The drop-in Modules (`CausalConv2d`, `TapConv2d`, `GatedActivation`, ...) take and return NCHW fp32 like the reference,
which costs a layout conversion on both sides of every module.  The model stacks (`models/gated_pixel_cnn.py`,
`models/pixel_snail.py`) instead keep every activation pixel-major between the image-channel input layer and the
logits: `[P = N*H*W, C]` matrices, bf16 where the tensor is only ever a tensor-core operand, fp32 for residual
streams.  Each function here is one `torch.autograd.Function` over such matrices whose forward / backward are the
C-ABI kernels:

  * `conv`      any stride-1 convolution of the path as a tap loop on the tcgen05 GEMM (`pg_gemm_bf16_conv`: the
                shifted input is read in place through 4-D TMA boxes, no im2col / gather buffer), with the
                bias, an fp32 residual, and the NEXT layer's input activation fused into the epilogue;
                backward = one wgrad and one dgrad launch of the same kernel, the dgrad epilogue applying the
                derivative of THIS layer's input activation;
  * `small_conv` image-channel input layers (direct fp32 kernel) straight to pixel-major;
  * `gated`     GatedActivation; `act_cast` materialises act(x) in bf16 where no producer epilogue could.

Reference call sites: gated_pixel_cnn.py:112-130, pixel_snail.py:27-28,52-56,112-119, nn/convolution.py:41-43,62-66.
"""

import collections

import torch

from .. import _lib as L
from .. import ops
from .tapconv import conv_taps, packed_tap_weight

F32, BF16 = torch.float32, torch.bfloat16
Geom = collections.namedtuple("Geom", "n h w")


def supported(h, w, channels):
    """True when every wide convolution of a stack with these channel counts can run as a TMA tap loop."""
    return all(L.conv_gemm_supported(h, w, c) for c in channels)


# --------------------------------------------------------------------------------------------------
# layout boundary
# --------------------------------------------------------------------------------------------------
class _FromPM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_pm, geom, c):
        ctx.geom, ctx.width = geom, x_pm.shape[1]
        return ops.pm_to_nchw(x_pm, geom.n, c, geom.h, geom.w)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_pm(dy, F32, width=ctx.width), None, None


def from_pm(x_pm, geom, c):
    """[P, >=c] fp32 pixel-major -> [N, c, H, W] fp32 (the logits at the Module boundary)."""
    return _FromPM.apply(x_pm, geom, c)


class _ToPM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, width):
        n, c, h, w = x.shape
        ctx.shape = (n, c, h, w)
        return ops.nchw_to_pm(x, BF16, width=width)

    @staticmethod
    def backward(ctx, dy):
        n, c, h, w = ctx.shape
        return ops.pm_to_nchw(dy.float().contiguous(), n, c, h, w), None


def to_pm_bf16(x, width=None):
    """[N, C, H, W] fp32 -> [P, width >= C] bf16 (extra columns zero)."""
    return _ToPM.apply(x, width or ops.round_up(x.shape[1], 8))


# --------------------------------------------------------------------------------------------------
# elementwise
# --------------------------------------------------------------------------------------------------
class _ActCast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
        L.act_cast(x, act, out)
        ctx.act = act
        ctx.save_for_backward(out if act != L.ACT_NONE else None)
        ctx.in_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, dy):
        (out,) = ctx.saved_tensors
        if ctx.act == L.ACT_NONE:
            return dy.to(ctx.in_dtype), None
        # act' from the activated value (relu / elu): torch elementwise on a [P, C] matrix, off the hot path (the
        # stacks take the fused route: the consumer's dgrad epilogue applies the derivative)
        a = out.float()
        d = (a > 0).float() if ctx.act == L.ACT_RELU else torch.where(a > 0, torch.ones_like(a), a + 1)
        return (dy.float() * d).to(ctx.in_dtype), None


def act_cast(x, act=L.ACT_NONE):
    """bf16(act(x)) of a pixel-major matrix (fp32 or bf16)."""
    if act == L.ACT_NONE and x.dtype == BF16:
        return x
    return _ActCast.apply(x, act)


class _Gated(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        P, c2 = x.shape
        y = torch.empty(P, c2 // 2, dtype=BF16, device=x.device)
        L.gated_act_fwd(x, y, act)
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dyc = dy.contiguous()
        if dyc.dtype != x.dtype:
            dyc = dyc.to(x.dtype)
        dx = torch.empty_like(x)
        L.gated_act_bwd(x, dyc, dx, ctx.act)
        return dx, None


def gated(x, act):
    """act(x[:, :C]) * sigmoid(x[:, C:]) -> bf16 [P, C] (reference nn/convolution.py:62-66)."""
    return _Gated.apply(x.contiguous(), act)


class _GatedRes(torch.autograd.Function):
    """res + gate(x) in one pass (fp32 stream in / out); backward: d res = dy, d x = gate'(x) dy."""

    @staticmethod
    def forward(ctx, x, res, act):
        y = torch.empty_like(res)
        L.gated_res_fwd(x, res, y, act)
        ctx.save_for_backward(x)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        L.gated_act_bwd(x, dy, dx, ctx.act)  # fp32 dy over a bf16 x is a supported combination
        return dx, dy, None


def gated_res(x, res, act):
    """res + act(x[:, :C]) * sigmoid(x[:, C:]) -> fp32 [P, C]: a gated residual block's output stream."""
    return _GatedRes.apply(x.contiguous(), res.contiguous(), act)


# --------------------------------------------------------------------------------------------------
# convolutions
# --------------------------------------------------------------------------------------------------
class _SmallConv(torch.autograd.Function):
    """Image-channel input convolution (Cin*kh*kw <= 160): NCHW fp32 image in, pixel-major fp32 out."""

    @staticmethod
    def forward(ctx, x, weight, bias, padding):
        x = x.contiguous().float()
        n, _, h, w = x.shape
        cout = weight.shape[0]
        out = torch.empty(n * h * w, cout, dtype=F32, device=x.device)
        L.conv_small_fwd(x, weight.detach().contiguous(), None if bias is None else bias.detach(), padding, out_f32=out)
        ctx.save_for_backward(x, weight)
        ctx.padding, ctx.has_bias = padding, bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous().float()
        dw = torch.zeros_like(weight)
        db = torch.zeros(weight.shape[0], dtype=F32, device=dy.device) if ctx.has_bias else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        L.conv_small_bwd(x, weight.detach().contiguous(), dy, ctx.padding, dw=dw, dbias=db, dx=dx)
        return dx, dw, db, None


def small_conv(x_nchw, weight, bias, padding):
    return _SmallConv.apply(x_nchw, weight, bias, tuple(padding))


COMPANION, PRE_GRAD, POST = 0, 1, 2  # what the activated output `ya` of a conv is to autograd (see `conv`)


class _Conv(torch.autograd.Function):
    """y = conv(xa) + bias (+ res), xa = bf16(in_act(x)) given by the caller.  Returns (y, ya): y fp32 / bf16 / None,
    ya = bf16(emit(y)) or None."""

    @staticmethod
    def forward(ctx, x, xa, weight, bias, res, geom, padding, in_act, emit, emit_mode, out_f32, want_main):
        cout, cin, kh, kw = weight.shape
        cin_p = xa.shape[1]
        taps = conv_taps(kh, kw, padding[0], padding[1])
        pointwise = len(taps) == 1 and taps[0] == (0, 0)
        wcat = packed_tap_weight(weight, cin_p)
        b = None if bias is None else bias.detach()
        want_act = emit is not None
        kw_out = dict(act=emit if want_act else L.ACT_NONE, res0=res, want_bf16=want_act,
                      want_pre=want_main and not out_f32, want_f32=want_main and out_f32)
        if pointwise:
            ya, yb, yf = ops.linear_fwd(xa, wcat, b, **kw_out)
        else:
            ya, yb, yf = ops.conv_fwd(xa, wcat, b, geom.n, geom.h, geom.w, taps, **kw_out)
        # backward needs the operand itself (wgrad); the activated input also yields in_act' (dgrad epilogue), the
        # activated output yields emit' when it is a true post-activation output
        # undefined output gradients arrive as None, not as zero tensors: the companion output is never differentiated, and
        # a materialised zero gradient for it costs a fill, a dtype conversion and an add per layer
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(xa, wcat, ya if (want_act and emit_mode == POST and emit != L.ACT_NONE) else None)
        ctx.meta = (geom, taps, pointwise, in_act, weight.shape, bias is not None, None if res is None else res.dtype,
                    x.dtype, emit, emit_mode)
        y = (yf if out_f32 else yb) if want_main else None
        if ya is not None and emit_mode == COMPANION:
            ctx.mark_non_differentiable(ya)
        return y, ya

    @staticmethod
    def backward(ctx, dy, dya):
        xa, wcat, ya = ctx.saved_tensors
        geom, taps, pointwise, in_act, wshape, has_bias, res_dtype, x_dtype, emit, emit_mode = ctx.meta
        cout, cin, kh, kw = wshape
        cin_p = xa.shape[1]
        cout_p = ops.round_up(cout, 8)
        T = len(taps)
        if dy is None and dya is None:  # nothing downstream used this convolution
            return (None,) * 12
        if emit_mode == COMPANION:
            dya = None
        if dya is not None and emit_mode == POST and emit != L.ACT_NONE:
            # gradient w.r.t. the activated output: back through emit (relu / elu) from the activated value itself
            d = torch.empty(ya.shape, dtype=BF16, device=ya.device)
            L.dact_from_out(dya.contiguous(), ya, emit, d)
            dya = d
        if dy is None:
            dy = dya
        elif dya is not None:
            dy = dy + dya.to(dy.dtype)
        dy = dy.contiguous()
        if dy.dtype == BF16 and cout_p == cout:
            dyb = dy
        elif cout_p == cout:
            dyb = torch.empty(dy.shape, dtype=BF16, device=dy.device)
            L.act_cast(dy, L.ACT_NONE, dyb)
        else:  # a handful of output channels (the logits): pad the operand to the 16-byte TMA pitch
            dyb = torch.zeros(dy.shape[0], cout_p, dtype=BF16, device=dy.device)
            dyb[:, :cout] = dy
        db = dw = None
        if ctx.needs_input_grad[2]:
            # weight and bias gradient in one launch: the wgrad GEMM reduces the dy tiles it stages (ops.linear_wgrad)
            dwcat = torch.zeros(cout_p, T * cin_p, dtype=F32, device=dy.device)
            dbp = torch.zeros(cout_p, dtype=F32, device=dy.device) if has_bias else None
            if pointwise:
                ops.linear_wgrad(dyb, xa, dwcat, db_out=dbp)
            else:
                ops.conv_wgrad(dyb, xa, dwcat, geom.n, geom.h, geom.w, taps, db_out=dbp)
            dw = dwcat[:cout].view(cout, kh, kw, cin_p)[..., :cin].permute(0, 3, 1, 2).contiguous()
            db = dbp[:cout] if has_bias else None
        elif has_bias:
            db = ops.bias_grad(dyb[:, :cout])
        dx = None
        if ctx.needs_input_grad[0]:
            want_f32 = x_dtype == F32
            dact = L.DACT_FROM_OUT.get(in_act, L.ACT_NONE)
            aux = xa if dact != L.ACT_NONE else None
            if pointwise:
                r = ops.linear_dgrad(dyb[:, :cout], wcat, aux=aux, dact=dact, want_f32=want_f32)
                dx = r[1] if want_f32 else r
            else:
                dxb, dxf = ops.conv_dgrad(dyb, wcat, cin_p, geom.n, geom.h, geom.w, taps, aux=aux, dact=dact,
                                          want_f32=want_f32, want_bf16=not want_f32)
                dx = dxf if want_f32 else dxb
        dres = None
        if res_dtype is not None:  # d(res) = dy: hand over the copy that already has the residual's dtype
            dres = dy if dy.dtype == res_dtype else (dyb if (res_dtype == BF16 and cout_p == cout) else dy.to(res_dtype))
        return dx, None, dw, db, dres, None, None, None, None, None, None, None


def conv(x, weight, bias, geom, padding=(0, 0), *, in_act=L.ACT_NONE, xa=None, res=None, emit=None, emit_mode=COMPANION,
         out_f32=False, want_main=True):
    """Convolution of a pixel-major activation.

    x        [P, Cin_p] differentiable input (bf16, or an fp32 residual stream), BEFORE its input activation;
    in_act   activation the reference applies in front of this conv (ReLU / ELU / none); its derivative is applied by
             this conv's dgrad epilogue, so the gradient this op returns for x is w.r.t. the PRE-activation value;
    xa       bf16(in_act(x)) if a producer epilogue already emitted it (else built here with one elementwise pass);
    res      optional [P, Cout] added to the output in the epilogue: fp32 (a residual / skip stream) or bf16 (a short-lived
             sum such as GatedPixelCNN's vertical-to-horizontal link; its gradient then stays bf16 as well);
    emit     activation id (or ACT_NONE for a plain bf16 copy) of a second, bf16 output produced by the same epilogue;
    emit_mode COMPANION: `ya` is a non-differentiable operand copy of y (pass it as `xa` to the consumers of y);
             PRE_GRAD: `ya` stands for y in the graph and may only feed `conv(ya, in_act=emit, xa=ya)`, whose fused
             derivative makes the gradient it receives the gradient w.r.t. y (use with want_main=False: the
             pre-activation tensor is then never written);  POST: `ya` is an ordinary activated output;
    out_f32  the main output is fp32 (a stream) instead of bf16.
    Returns (y, ya)."""
    if xa is None:
        xa = act_cast(x, in_act) if (in_act != L.ACT_NONE or x.dtype != BF16) else x
    return _Conv.apply(x, xa.detach() if xa is not x else xa, weight, bias, res, geom, tuple(padding), in_act, emit, emit_mode,
                       out_f32, want_main)
