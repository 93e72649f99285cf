"""Host-side primitives on pixel-major tensors (thin, allocation + one C-ABI call each).

Everything the model stacks and nn modules do on the device goes through these helpers, which only
allocate outputs and forward to `_lib` (libpg_b200.so).  There is no autograd here: forward and backward
are explicit functions, composed by the `torch.autograd.Function`s in `nn/` and `models/`.

Conventions: `P` = N*H*W pixels; activations are [P, C] row-major ("pixel-major", i.e. NHWC); GEMM
operands are bf16, the residual stream and all gradients that are summed are fp32.
"""

import os

import torch

from . import _lib as L

BF16 = torch.bfloat16
F32 = torch.float32

# 0 = tcgen05 kernels (product); 1 = SIMT cross-check kernels (tests/debug only, set by tests).
ATTN_IMPL = int(os.environ.get("PG_ATTN_IMPL", "0"))
GEMM_IMPL = int(os.environ.get("PG_GEMM_IMPL", "0"))

HEAD_SLOT = 64  # attention kernels work on 64-wide head slots (dk padded with zero columns)


def empty(shape, dtype, like):
    return torch.empty(shape, dtype=dtype, device=like.device)


def zeros(shape, dtype, like):
    return torch.zeros(shape, dtype=dtype, device=like.device)


def round_up(v, m):
    return (v + m - 1) // m * m


def to_bf16(t):
    """fp32 -> bf16 copy through pg_cast_f32_to_bf16 (weights packing)."""
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=BF16, device=t.device)
    L.cast_bf16(t.view(-1), out.view(-1))
    return out


def pack_weight(w, pad_in=None):
    """[Cout, Cin, 1, 1] fp32 Parameter -> [Cout, Cin_p] bf16 matrix (Cin padded to a multiple of 8 with zeros)."""
    w2 = w.detach().reshape(w.shape[0], -1)
    cin = w2.shape[1]
    cin_p = pad_in or round_up(cin, 8)
    if cin_p == cin:
        return to_bf16(w2)
    out = torch.zeros(w2.shape[0], cin_p, dtype=BF16, device=w.device)
    out[:, :cin] = to_bf16(w2)
    return out


def nchw_to_pm(x, dtype, width=None):
    """[N, C, H, W] fp32 -> [N*H*W, width>=C] pixel-major (extra columns zero)."""
    x = x.contiguous()
    if x.dtype != F32:
        x = x.float()
    n, c, h, w = x.shape
    width = width or c
    if width == c:
        out = torch.empty(n * h * w, c, dtype=dtype, device=x.device)
        L.nchw_to_pm(x, out)
    else:
        out = torch.zeros(n * h * w, width, dtype=dtype, device=x.device)
        L.nchw_to_pm(x, out[:, :c])
    return out


def pm_to_nchw(x_pm, n, c, h, w, act=L.ACT_NONE):
    out = torch.empty(n, c, h, w, dtype=F32, device=x_pm.device)
    L.pm_to_nchw(x_pm[:, :c] if x_pm.shape[1] != c else x_pm, out, act=act)
    return out


# --------------------------------------------------------------------------------------------------
# Linear (1x1 conv) forward / dgrad / wgrad on pixel-major activations
# --------------------------------------------------------------------------------------------------
def linear_fwd(a, w, bias=None, *, act=L.ACT_NONE, res0=None, res1=None, want_bf16=True, want_pre=False,
               want_f32=False, n_out=None, skinny=False, pre_deriv=False):
    """y = a @ w.T (+bias) (+res0 +res1).  a: [P, K] bf16, w: [Cout, K] bf16.
    Returns (out_bf16 = act(pre), out_pre = bf16(pre), out_f32 = pre), each None unless requested.
    pre_deriv: out_pre holds act'(pre) instead (consumed by linear_dgrad(dact=L.ACT_GIVEN))."""
    P, K = a.shape
    n = n_out or w.shape[0]
    ob = empty((P, n), BF16, a) if want_bf16 else None
    op = empty((P, n), BF16, a) if want_pre else None
    of = empty((P, n), F32, a) if want_f32 else None
    L.gemm(a, w, P, n, K, bias=bias, res0=res0, res1=res1, out_bf16=ob, out_pre=op, out_f32=of,
           act=(act | L.ACT_STORE_DERIV) if (pre_deriv and want_pre) else act,
           impl=2 if (skinny and P <= 32) else GEMM_IMPL)
    return ob, op, of


def linear_dgrad(dy, w, *, aux=None, dact=L.ACT_NONE, want_f32=False, k_in=None):
    """dx = dy @ w (optionally * act'(aux)).  dy: [P, Cout] bf16, w: [Cout, Cin] bf16 (read MN-major)."""
    P, cout = dy.shape
    cin = k_in or w.shape[1]
    ob = empty((P, cin), BF16, dy)
    of = empty((P, cin), F32, dy) if want_f32 else None
    L.gemm(dy, w[:, :cin], P, cin, min(cout, w.shape[0]), b_mn=True, aux=aux, dact=dact, out_bf16=ob, out_f32=of,
           impl=GEMM_IMPL)
    return (ob, of) if want_f32 else ob


def _split_k_for(m_out, n_out, k):
    """Split-K factor of a wgrad GEMM: the largest one whose work items (output tiles x splits) still fit ONE wave of
    the persistent grid.  Rounding up instead (e.g. 32 tiles x 5 = 160 items on 148 SMs) makes a few CTAs run two
    items back to back, so the launch lasts two split-lengths: 2 x 205 k-iterations instead of 1 x 256 at the C5 MLP
    shapes (measured: split 4 and split 8 of that GEMM take the same 114-119 us, profiles/r01_gemm_microbench_final.txt)."""
    tiles = ((m_out + 127) // 128) * ((n_out + 255) // 256)
    sms = L.sm_count()
    if os.environ.get("PG_SPLITK_ROUND_UP") == "1":  # previous heuristic, kept for A/B runs
        want = max(1, (sms + tiles - 1) // tiles)
    else:
        want = max(1, sms // tiles)
    k_iters = (k + 63) // 64
    return max(1, min(want, k_iters // 8 if k_iters >= 16 else 1))


def linear_wgrad(dy, a, dw_out, db_out=None):
    """dw_out[Cout, Cin] += dy.T @ a over the pixel dimension (fp32 atomics, split along pixels).
    db_out (fp32 [Cout], pre-zeroed or holding a running sum): += column sums of dy, reduced by the same launch from the
    dy tiles it stages anyway (the bias gradient without a second pass over dy)."""
    P, cout = dy.shape
    cin = a.shape[1]
    assert dw_out.dtype == F32 and dw_out.shape[0] >= cout
    if db_out is not None and GEMM_IMPL != 0:   # the SIMT cross-check build of the stacks: separate column sums
        L.colsum(dy, db_out, accumulate=True)
        db_out = None
    L.gemm(dy, a, cout, cin, P, a_mn=True, b_mn=True, out_f32=dw_out, accumulate=True,
           split_k=_split_k_for(cout, cin, P), impl=GEMM_IMPL, bias_grad=db_out)


# --------------------------------------------------------------------------------------------------
# Tap-loop convolutions (im2col-free): the shifted operand is read in place through 4-D TMA boxes
# --------------------------------------------------------------------------------------------------
def conv_fwd(x, wcat, bias, n_img, h, w, taps, *, act=L.ACT_NONE, res0=None, res1=None, want_bf16=True, want_pre=False,
             want_f32=False, pre_deriv=False):
    """y[p] = bias + sum_t W_t . x[p + taps[t]] (zero outside the image).  x: [P, C] bf16 (C % 64 == 0),
    wcat: [Cout, T*C] bf16 with tap-major columns.  Outputs as linear_fwd."""
    P, C = x.shape
    n = wcat.shape[0]
    ob = empty((P, n), BF16, x) if want_bf16 else None
    op = empty((P, n), BF16, x) if want_pre else None
    of = empty((P, n), F32, x) if want_f32 else None
    L.gemm_conv(x, wcat, P, n, len(taps) * C, L.CONV_FWD, n_img, h, w, C, taps, bias=bias, res0=res0, res1=res1,
                out_bf16=ob, out_pre=op, out_f32=of, act=(act | L.ACT_STORE_DERIV) if (pre_deriv and want_pre) else act)
    return ob, op, of


def conv_dgrad(dy, wcat, cin, n_img, h, w, taps, *, aux=None, dact=L.ACT_NONE, want_f32=False, want_bf16=True, res0=None):
    """dx[p] = sum_t W_t^T . dy[p - taps[t]] (optionally * act'(aux), + res0).  dy: [P, Cout] bf16 (Cout % 64 == 0),
    wcat: [Cout, T*cin] bf16."""
    P, cout = dy.shape
    ob = empty((P, cin), BF16, dy) if want_bf16 else None
    of = empty((P, cin), F32, dy) if want_f32 else None
    L.gemm_conv(dy, wcat, P, cin, len(taps) * cout, L.CONV_DGRAD, n_img, h, w, cout, [(-a, -b) for a, b in taps], aux=aux,
                dact=dact, res0=res0, out_bf16=ob, out_f32=of)
    return ob, of


def conv_wgrad(dy, x, dw_out, n_img, h, w, taps, db_out=None):
    """dw_out[Cout, T*C] += sum_p dy[p]^T . x[p + taps[t]] (fp32 accumulation, split along pixels); db_out as in
    linear_wgrad."""
    P, cout = dy.shape
    C = x.shape[1]
    T = len(taps)
    assert dw_out.dtype == F32 and dw_out.shape[0] >= cout and dw_out.shape[1] == T * C
    bn = 256 if C % 256 == 0 else (128 if C % 128 == 0 else 64)
    tiles = ((cout + 127) // 128) * (T * C // bn)
    k_iters = (P + 63) // 64
    split = max(1, min(max(1, L.sm_count() // tiles), k_iters // 8 if k_iters >= 16 else 1))
    L.gemm_conv(dy, x, cout, T * C, P, L.CONV_WGRAD, n_img, h, w, C, taps, out_f32=dw_out, accumulate=True, split_k=split,
                bias_grad=db_out)


def bias_grad(dy, out=None):
    """Column sums of dy [P, C] into a fp32 [C] vector."""
    C = dy.shape[1]
    if out is None:
        out = torch.zeros(C, dtype=F32, device=dy.device)
    L.colsum(dy, out, accumulate=True)
    return out


# --------------------------------------------------------------------------------------------------
# LayerNorm
# --------------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, want_bf16=True, want_f32=False):
    P, C = x.shape
    yb = empty((P, C), BF16, x) if want_bf16 else None
    yf = empty((P, C), F32, x) if want_f32 else None
    mean = empty((P,), F32, x)
    rstd = empty((P,), F32, x)
    L.layernorm_fwd(x, gamma, beta, eps, y_bf16=yb, y_f32=yf, mean=mean, rstd=rstd)
    return yb, yf, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dres0=None, dres1=None, want_bf16=True, want_f32=True, want_colsum=False,
                  stats=None):
    """Returns (dx_f32 (+dres0+dres1), dx_bf16, dgamma, dbeta[, colsum(dx)]).  `stats`: optional zero-filled fp32 [3, C]
    (a slice of the caller's gradient arena) that receives dgamma, dbeta and the column sums."""
    P, C = x.shape
    dxf = empty((P, C), F32, x) if want_f32 else None
    dxb = empty((P, C), BF16, x) if want_bf16 else None
    if stats is None:
        stats = zeros((3, C), F32, x)  # dgamma, dbeta, column sums of dx: one memset
    L.layernorm_bwd(dy, x, gamma, mean, rstd, dres0=dres0, dres1=dres1, dx_f32=dxf, dx_bf16=dxb, dgamma=stats[0],
                    dbeta=stats[1], dx_colsum=stats[2] if want_colsum else None)
    if want_colsum:
        return dxf, dxb, stats[0], stats[1], stats[2]
    return dxf, dxb, stats[0], stats[1]


# --------------------------------------------------------------------------------------------------
# Attention on padded 64-wide head slots
# --------------------------------------------------------------------------------------------------
def attn_fwd(q, k, v, n_img, seq, heads, dk_true, dv_slot, strict):
    """q, k: [P, heads*64] bf16 (dk_true valid columns per slot, rest zero); v: [P, heads*dv_slot]."""
    P = q.shape[0]
    o = empty((P, heads * dv_slot), BF16, q)
    lse = empty((n_img, heads, seq), F32, q)
    L.causal_attn_fwd(q, k, v, o, lse, n_img, seq, heads, HEAD_SLOT, dv_slot, strict, impl=ATTN_IMPL,
                      dk_true=dk_true)
    return o, lse


def attn_bwd(q, k, v, o, do, lse, dq, dk, dv, n_img, seq, heads, dk_true, dv_slot, strict):
    P = q.shape[0]
    delta = empty((n_img, heads, seq), F32, q)
    dq_acc = empty((P, heads * HEAD_SLOT), F32, q) if ATTN_IMPL != 1 else None  # cleared by the library's delta pass
    L.causal_attn_bwd(q, k, v, o, do, lse, delta, dq_acc, dq, dk, dv, n_img, seq, heads, HEAD_SLOT, dv_slot, strict,
                      impl=ATTN_IMPL, dk_true=dk_true)
