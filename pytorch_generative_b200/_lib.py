"""ctypes binding of libpg_b200.so (the C ABI declared in include/pg_b200.h).

This is the only place Python touches the native library.  Tensors cross the boundary as raw device
pointers plus sizes; the current torch CUDA stream is passed explicitly to every call, so the kernels
are ordered with the rest of the autograd graph (forward on the main thread, backward on autograd's
device thread).  There is deliberately NO CPU fallback: if the shared library is missing or the inputs
are not CUDA tensors, the call raises.
"""

import ctypes
import functools
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PG_B200_LIB: developer knob to load an alternative build of the same ABI (kernel A/B experiments)
LIB_PATH = os.environ.get("PG_B200_LIB") or os.path.join(_HERE, "libpg_b200.so")

ACT_NONE, ACT_RELU, ACT_GELU, ACT_ELU, ACT_TANH = 0, 1, 2, 3, 4
ACT_GIVEN = 5  # dact only: aux already holds the derivative
ACT_RELU_OUT, ACT_ELU_OUT = 6, 7  # dact only: aux holds the activated value
DACT_FROM_OUT = {ACT_RELU: ACT_RELU_OUT, ACT_ELU: ACT_ELU_OUT}
ACT_STORE_DERIV = 0x100  # OR-ed into act: out_pre receives act'(pre)
ACT_RES_BF16 = 0x200  # OR-ed into act: res0 / res1 are bf16 matrices (set by _epilogue from the tensors' dtype)
ACT_BY_NAME = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU, "elu": ACT_ELU, "tanh": ACT_TANH}

_vp, _i32, _i64, _f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float


class GemmEpilogue(ctypes.Structure):
    """Mirror of `pg_gemm_epilogue` (include/pg_b200.h)."""

    _fields_ = [
        ("bias", _vp), ("aux", _vp), ("res0", _vp), ("res1", _vp),
        ("out_bf16", _vp), ("out_pre", _vp), ("out_f32", _vp),
        ("ld_aux", _i64), ("ld_res", _i64), ("ld_out_bf16", _i64), ("ld_out_pre", _i64), ("ld_out_f32", _i64),
        ("act", ctypes.c_int32), ("dact", ctypes.c_int32), ("accumulate", ctypes.c_int32), ("alpha", _f32),
        ("bias_grad", _vp),
    ]


class ConvGeom(ctypes.Structure):
    """Mirror of `pg_conv_geom` (include/pg_b200.h)."""

    _fields_ = [("mode", ctypes.c_int32), ("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
                ("C", ctypes.c_int32), ("n_taps", ctypes.c_int32), ("dy", ctypes.c_int32 * 32), ("dx", ctypes.c_int32 * 32)]


CONV_FWD, CONV_DGRAD, CONV_WGRAD = 1, 2, 3

# name -> argtypes (restype is always int unless listed in _SPECIAL)
_SIGNATURES = {
    "pg_gemm_bf16": [_vp, _i32, _i64, _vp, _i32, _i64, _i32, _i32, _i32, _i32, ctypes.POINTER(GemmEpilogue), _i32, _vp],
    "pg_gemm_bf16_conv": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, ctypes.POINTER(GemmEpilogue), ctypes.POINTER(ConvGeom), _vp],
    "pg_colsum_bf16": [_vp, _i64, _i32, _i32, _vp, _i32, _vp],
    "pg_colsum_f32": [_vp, _i64, _i32, _i32, _vp, _i32, _vp],
    "pg_layernorm_fwd": [_vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp],
    "pg_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "pg_gated_act_fwd": [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "pg_gated_act_bwd": [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp],
    "pg_gated_res_fwd": [_vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp],
    "pg_dact_from_out": [_vp, _i32, _vp, _i64, _i32, _vp, _vp],
    "pg_bce_logits_fwd_bwd": [_vp, _vp, _i64, _f32, _vp, _vp, _vp],
    "pg_nchw_to_pm": [_vp, _i32, _i32, _i32, _vp, _i32, _i64, _vp],
    "pg_pm_to_nchw": [_vp, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _vp],
    "pg_dact_mul": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _vp],
    "pg_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "pg_act_cast_bf16": [_vp, _i32, _i64, _i32, _i32, _i32, _vp, _i64, _vp],
    "pg_causal_attn_fwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _vp],
    "pg_causal_attn_bwd": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _i64,
                           _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _i32, _i32, _vp],
    "pg_conv_small_fwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp],
    "pg_conv_small_bwd": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "pg_attn_decode": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32,
                       _f32, _i32, _vp],
    "pg_linear_attn_fwd": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "pg_linear_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "pg_cast_multi_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _vp],
    "pg_grad_sqnorm": [_vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "pg_adam_step": [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _f32, _f32, ctypes.c_double, ctypes.c_double,
                     ctypes.c_double, ctypes.c_double, _i32, _vp, _vp],
    "pg_tap_gather": [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp],
    "pg_tap_scatter": [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _i64, _vp, _vp, _i64, _vp],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["pg_abi_version", "pg_last_error", "pg_sm_count", "pg_launch_count",
                                                 "pg_reserve_sms"])

_lib = None


def load():
    """Loads libpg_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m pytorch_generative_b200._build` "
            "(or __graft_entry__.build()); the CUDA path has no CPU fallback"
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.pg_last_error.restype = ctypes.c_char_p
    lib.pg_last_error.argtypes = []
    lib.pg_abi_version.restype = ctypes.c_int
    lib.pg_abi_version.argtypes = []
    lib.pg_sm_count.restype = ctypes.c_int
    lib.pg_sm_count.argtypes = []
    lib.pg_launch_count.restype = ctypes.c_ulonglong
    lib.pg_launch_count.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    if lib.pg_abi_version() != 1:
        raise RuntimeError(f"libpg_b200.so ABI version {lib.pg_abi_version()} != 1")
    _lib = lib
    return lib


def _check(rc, name):
    if rc != 0:
        raise RuntimeError(f"{name} failed: {_lib.pg_last_error().decode(errors='replace')}")


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("pytorch_generative_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _device_guarded(fn):
    """Runs a binding on the device its tensors live on: every operand must share one CUDA device; when that is not
    the current device the call (stream lookup, TMA descriptor encode, launch) happens under `torch.cuda.device(idx)`,
    like torch's own ops do."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        idx = None
        for a in (*args, *kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if idx is None:
                    idx = a.device.index
                elif a.device.index != idx:
                    raise RuntimeError(f"{fn.__name__}: operands live on different CUDA devices ({idx} and {a.device.index})")
        if idx is None or idx == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(idx):
            return fn(*args, **kwargs)

    return wrapper


def _pm(t):
    """Checks a pixel-major 2-D view (unit inner stride) and returns (ptr, pitch)."""
    assert t.dim() == 2 and t.stride(1) == 1, f"expected a [P, C] matrix with unit inner stride, got {t.shape} {t.stride()}"
    return _ptr(t), t.stride(0)


def launch_count():
    """Kernels launched by libpg_b200.so so far in this process."""
    return int(load().pg_launch_count())


# Optional per-call timing hook used by bench.py: when set, gemm() brackets its launch with CUDA events on
# the launching stream and reports (flops, start_event, end_event, algorithmic_bytes).
gemm_timing_hook = None

_sm_count = None


def sm_count():
    global _sm_count
    if _sm_count is None:
        _sm_count = load().pg_sm_count()
    return _sm_count


def reserve_sms(n):
    """SMs left out of the persistent grids (see pg_reserve_sms); returns the previous setting."""
    global _sm_count
    lib = load()
    lib.pg_reserve_sms.restype = ctypes.c_int
    lib.pg_reserve_sms.argtypes = [ctypes.c_int]
    old = lib.pg_reserve_sms(int(n))
    _sm_count = None
    return old


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
@_device_guarded
def gemm(A, B, M, N, K, *, a_mn=False, b_mn=False, bias=None, aux=None, dact=ACT_NONE, res0=None, res1=None,
         out_bf16=None, out_pre=None, out_f32=None, act=ACT_NONE, accumulate=False, alpha=1.0, split_k=1, impl=0,
         bias_grad=None):
    """acc = A·Bᵀ (see pg_gemm_bf16 in include/pg_b200.h); all tensors are 2-D bf16/fp32 CUDA views.
    bias_grad (weight-gradient GEMMs, a_mn=True): fp32 [M] += sum_k A(m, k), reduced from the staged A tiles."""
    lib = load()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    a_ptr, lda = _pm(A)
    b_ptr, ldb = _pm(B)
    e = _epilogue(M, N, bias, aux, dact, res0, res1, out_bf16, out_pre, out_f32, act, accumulate, alpha, bias_grad)
    hook = gemm_timing_hook
    if hook is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _check(lib.pg_gemm_bf16(a_ptr, int(a_mn), lda, b_ptr, int(b_mn), ldb, M, N, K, split_k, ctypes.byref(e), impl,
                            _stream()), "pg_gemm_bf16")
    if hook is not None:
        ev1.record()
        hook(2.0 * M * N * K, ev0, ev1, _gemm_io(M, N, K, out_bf16, out_pre, out_f32, aux, res0, res1))


def _gemm_io(M, N, K, out_bf16, out_pre, out_f32, aux, res0, res1):
    """Algorithmic HBM bytes of one contraction: both operands once, every epilogue tensor once."""
    return 2 * (M * K + N * K) + M * N * (2 * (out_bf16 is not None) + 2 * (out_pre is not None) + 4 * (out_f32 is not None)
                                          + 2 * (aux is not None) + sum(r.element_size() for r in (res0, res1) if r is not None))


@_device_guarded
def gemm_conv(A, B, M, N, K, mode, n_img, H, W, C, taps, *, bias=None, aux=None, dact=ACT_NONE, res0=None, res1=None,
              out_bf16=None, out_pre=None, out_f32=None, act=ACT_NONE, accumulate=False, alpha=1.0, split_k=1,
              bias_grad=None):
    """Tap-loop convolution on the GEMM kernel (pg_gemm_bf16_conv); `taps` = [(dy, dx), ...] as the kernel applies them
    (the caller negates them for dgrad)."""
    lib = load()
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    a_ptr, lda = _pm(A)
    b_ptr, ldb = _pm(B)
    e = _epilogue(M, N, bias, aux, dact, res0, res1, out_bf16, out_pre, out_f32, act, accumulate, alpha, bias_grad)
    g = ConvGeom()
    g.mode, g.N, g.H, g.W, g.C, g.n_taps = mode, n_img, H, W, C, len(taps)
    for t, (dy, dx) in enumerate(taps):
        g.dy[t], g.dx[t] = int(dy), int(dx)
    hook = gemm_timing_hook
    if hook is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    _check(lib.pg_gemm_bf16_conv(a_ptr, lda, b_ptr, ldb, M, N, K, split_k, ctypes.byref(e), ctypes.byref(g), _stream()),
           "pg_gemm_bf16_conv")
    if hook is not None:
        ev1.record()
        # the shifted operand is read once per tap from L2 but only once from HBM
        io = _gemm_io(M, N, K, out_bf16, out_pre, out_f32, aux, res0, res1)
        if mode == CONV_WGRAD:
            io -= 2 * K * (N - C)
        else:
            io -= 2 * M * (K - C)
        hook(2.0 * M * N * K, ev0, ev1, io)


def conv_gemm_supported(H, W, C):
    """Geometry the TMA tap loop handles (pg_gemm_bf16_conv); other shapes go through pg_tap_gather."""
    return C % 64 == 0 and 1 <= W <= 64 and 64 % W == 0 and (H * W) % 128 == 0


def _epilogue(M, N, bias, aux, dact, res0, res1, out_bf16, out_pre, out_f32, act, accumulate, alpha, bias_grad=None):
    e = GemmEpilogue()
    if bias_grad is not None:
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() >= M and bias_grad.is_contiguous()
    e.bias_grad = _ptr(bias_grad)
    e.bias = _ptr(bias)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N and bias.is_contiguous()
    ld_res, res_dtype = 0, None
    for r in (res0, res1):
        if r is not None:
            assert r.dtype in (torch.float32, torch.bfloat16) and r.shape[0] >= M
            assert res_dtype in (None, r.dtype), "res0/res1 must share a dtype"
            p, ld = _pm(r)
            assert ld_res in (0, ld), "res0/res1 must share a pitch"
            ld_res, res_dtype = ld, r.dtype
    e.res0, e.res1, e.ld_res = _ptr(res0), _ptr(res1), ld_res
    if res_dtype == torch.bfloat16:
        act = act | ACT_RES_BF16
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        e.aux, e.ld_aux = _pm(aux)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.shape[0] >= M
        e.out_bf16, e.ld_out_bf16 = _pm(out_bf16)
    if out_pre is not None:
        assert out_pre.dtype == torch.bfloat16 and out_pre.shape[0] >= M
        e.out_pre, e.ld_out_pre = _pm(out_pre)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.shape[0] >= M
        e.out_f32, e.ld_out_f32 = _pm(out_f32)
    e.act, e.dact, e.accumulate, e.alpha = act, dact, int(accumulate), alpha
    return e


@_device_guarded
def colsum(x, out, accumulate=False):
    lib = load()
    p, ld = _pm(x)
    P, C = x.shape
    assert out.dtype == torch.float32 and out.numel() >= C
    fn = lib.pg_colsum_bf16 if x.dtype == torch.bfloat16 else lib.pg_colsum_f32
    _check(fn(p, ld, P, C, _ptr(out), int(accumulate), _stream()), "pg_colsum")


# ------------------------------------------------------------------------------------------------
# LayerNorm / gated activation / loss / converters
# ------------------------------------------------------------------------------------------------
@_device_guarded
def layernorm_fwd(x, gamma, beta, eps, y_bf16=None, y_f32=None, mean=None, rstd=None):
    lib = load()
    P, C = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    _check(lib.pg_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), P, C, eps, _ptr(y_bf16), _ptr(y_f32), _ptr(mean),
                                _ptr(rstd), _stream()), "pg_layernorm_fwd")


@_device_guarded
def layernorm_bwd(dy, x, gamma, mean, rstd, dres0=None, dres1=None, dx_f32=None, dx_bf16=None, dgamma=None,
                  dbeta=None, dx_colsum=None):
    lib = load()
    P, C = x.shape
    assert dy.is_contiguous() and x.is_contiguous()
    dy_b = _ptr(dy) if dy.dtype == torch.bfloat16 else None
    dy_f = _ptr(dy) if dy.dtype == torch.float32 else None
    _check(lib.pg_layernorm_bwd(dy_b, dy_f, _ptr(x), _ptr(gamma), _ptr(mean), _ptr(rstd), P, C, _ptr(dres0),
                                _ptr(dres1), _ptr(dx_f32), _ptr(dx_bf16), _ptr(dgamma), _ptr(dbeta), _ptr(dx_colsum),
                                _stream()),
           "pg_layernorm_bwd")


@_device_guarded
def gated_act_fwd(x, y, act):
    lib = load()
    P, C2 = x.shape
    assert x.is_contiguous() and y.is_contiguous() and y.shape == (P, C2 // 2)
    _check(lib.pg_gated_act_fwd(_ptr(x), int(x.dtype == torch.float32), P, C2 // 2, act, _ptr(y),
                                int(y.dtype == torch.float32), _stream()), "pg_gated_act_fwd")


@_device_guarded
def gated_res_fwd(x, res, y, act):
    """y = res + act(x[:, :C]) * sigmoid(x[:, C:]); res, y fp32 [P, C]."""
    P, C2 = x.shape
    assert x.is_contiguous() and res.is_contiguous() and y.is_contiguous() and res.dtype == y.dtype == torch.float32
    _check(load().pg_gated_res_fwd(_ptr(x), int(x.dtype == torch.float32), _ptr(res), P, C2 // 2, act, _ptr(y), _stream()),
           "pg_gated_res_fwd")


@_device_guarded
def dact_from_out(dy, ya, act, out):
    """out = bf16(dy * act'(pre)) with ya = act(pre) (relu / elu)."""
    assert dy.is_contiguous() and ya.is_contiguous() and out.is_contiguous() and ya.dtype == out.dtype == torch.bfloat16
    _check(load().pg_dact_from_out(_ptr(dy), int(dy.dtype == torch.float32), _ptr(ya), dy.numel(), act, _ptr(out), _stream()),
           "pg_dact_from_out")


@_device_guarded
def gated_act_bwd(x, dy, dx, act):
    lib = load()
    P, C2 = x.shape
    assert x.is_contiguous() and dy.is_contiguous() and dx.is_contiguous()
    _check(lib.pg_gated_act_bwd(_ptr(x), int(x.dtype == torch.float32), _ptr(dy), int(dy.dtype == torch.float32), P,
                                C2 // 2, act, _ptr(dx), int(dx.dtype == torch.float32), _stream()), "pg_gated_act_bwd")


@_device_guarded
def bce_logits(logits, target, grad_scale, loss_sum, dlogits=None):
    lib = load()
    assert logits.dtype == torch.float32 and target.dtype == torch.float32
    assert logits.is_contiguous() and target.is_contiguous() and logits.numel() == target.numel()
    _check(lib.pg_bce_logits_fwd_bwd(_ptr(logits), _ptr(target), logits.numel(), grad_scale, _ptr(loss_sum),
                                     _ptr(dlogits), _stream()), "pg_bce_logits_fwd_bwd")


@_device_guarded
def nchw_to_pm(x, out):
    """x: [N, C, H, W] fp32 contiguous -> out: [N*H*W, >=C] bf16/fp32 (pixel-major)."""
    lib = load()
    N, C, H, W = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    p, ld = _pm(out)
    _check(lib.pg_nchw_to_pm(_ptr(x), N, C, H * W, p, int(out.dtype == torch.float32), ld, _stream()), "pg_nchw_to_pm")


@_device_guarded
def pm_to_nchw(x_pm, out, act=ACT_NONE):
    lib = load()
    N, C, H, W = out.shape
    assert out.dtype == torch.float32 and out.is_contiguous()
    p, ld = _pm(x_pm)
    _check(lib.pg_pm_to_nchw(p, int(x_pm.dtype == torch.float32), ld, N, C, H * W, act, _ptr(out), _stream()),
           "pg_pm_to_nchw")


@_device_guarded
def dact_mul(dy, pre, act, out):
    """out = bf16(dy * act'(pre)); dy/out bf16 [P, C] views, pre fp32."""
    lib = load()
    (dp, ldd), (pp, ldp), (op, ldo) = _pm(dy), _pm(pre), _pm(out)
    P, C = dy.shape
    assert dy.dtype == torch.bfloat16 and pre.dtype == torch.float32 and out.dtype == torch.bfloat16
    _check(lib.pg_dact_mul(dp, ldd, pp, ldp, P, C, act, op, ldo, _stream()), "pg_dact_mul")


@_device_guarded
def act_cast(x, act, out):
    """out = bf16(act(x)); x: [P, C] fp32 or bf16 view, out: [P, C] bf16 view."""
    lib = load()
    (xp, ldx), (op, ldo) = _pm(x), _pm(out)
    P, C = x.shape
    assert out.dtype == torch.bfloat16 and out.shape == x.shape and x.dtype in (torch.float32, torch.bfloat16)
    _check(lib.pg_act_cast_bf16(xp, int(x.dtype == torch.float32), ldx, P, C, act, op, ldo, _stream()), "pg_act_cast_bf16")


@_device_guarded
def cast_bf16(x, y):
    lib = load()
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.is_contiguous() and y.is_contiguous()
    _check(lib.pg_cast_f32_to_bf16(_ptr(x), _ptr(y), x.numel(), _stream()), "pg_cast_f32_to_bf16")


# ------------------------------------------------------------------------------------------------
# Attention / small conv
# ------------------------------------------------------------------------------------------------
@_device_guarded
def causal_attn_fwd(q, k, v, o, lse, N, S, H, dk, dv, strict, impl=0, dk_true=None):
    """dk is the column width of a head slot; dk_true (default dk) sets the 1/sqrt(dk) scale."""
    lib = load()
    (qp, ldq), (kp, ldk), (vp, ldv), (op, ldo) = _pm(q), _pm(k), _pm(v), _pm(o)
    scale = 1.0 / math.sqrt(dk_true or dk)
    _check(lib.pg_causal_attn_fwd(qp, ldq, kp, ldk, vp, ldv, op, ldo, _ptr(lse), N, S, H, dk, dv, scale, int(strict),
                                  impl, _stream()), "pg_causal_attn_fwd")


@_device_guarded
def causal_attn_bwd(q, k, v, o, do, lse, delta, dq_accum, dq, dk_, dv_, N, S, H, dk, dv, strict, impl=0, dk_true=None):
    lib = load()
    (qp, ldq), (kp, ldk), (vp, ldv), (op, ldo), (dop, lddo) = _pm(q), _pm(k), _pm(v), _pm(o), _pm(do)
    (dqp, lddq), (dkp, lddk), (dvp, lddv) = _pm(dq), _pm(dk_), _pm(dv_)
    scale = 1.0 / math.sqrt(dk_true or dk)
    _check(lib.pg_causal_attn_bwd(qp, ldq, kp, ldk, vp, ldv, op, ldo, dop, lddo, _ptr(lse), _ptr(delta),
                                  _ptr(dq_accum), dqp, lddq, dkp, lddk, dvp, lddv, N, S, H, dk, dv, scale, int(strict),
                                  impl, _stream()), "pg_causal_attn_bwd")


@_device_guarded
def conv_small_fwd(x, w, bias, pad, out_f32=None, out_bf16=None, act_bf16=ACT_NONE, pre_act=ACT_NONE):
    lib = load()
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.float32 and w.dtype == torch.float32
    _check(lib.pg_conv_small_fwd(_ptr(x), _ptr(w), _ptr(bias), N, Cin, H, W, Cout, kh, kw, pad[0], pad[1], pre_act,
                                 _ptr(out_f32), _ptr(out_bf16), act_bf16, _stream()), "pg_conv_small_fwd")


@_device_guarded
def conv_small_bwd(x, w, dy_pm, pad, dw=None, dbias=None, dx=None, pre_act=ACT_NONE):
    lib = load()
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    assert dy_pm.dtype == torch.float32 and dy_pm.is_contiguous()
    _check(lib.pg_conv_small_bwd(_ptr(x), _ptr(w), _ptr(dy_pm), N, Cin, H, W, Cout, kh, kw, pad[0], pad[1], pre_act,
                                 _ptr(dw), _ptr(dbias), _ptr(dx), _stream()), "pg_conv_small_bwd")


@_device_guarded
def linear_attn_fwd(q, k, v, out):
    """q, k: [B, L, d]; v, out: [B, L, dv]; fp32 contiguous."""
    B, Lq, d = q.shape
    for t in (q, k, v, out):
        assert t.dtype == torch.float32 and t.is_contiguous()
    _check(load().pg_linear_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(out), B, Lq, d, v.shape[2], _stream()), "pg_linear_attn_fwd")


@_device_guarded
def linear_attn_bwd(q, k, v, g, dq, dk, dv):
    B, Lq, d = q.shape
    for t in (q, k, v, g, dq, dk, dv):
        assert t.dtype == torch.float32 and t.is_contiguous()
    _check(load().pg_linear_attn_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(g), _ptr(dq), _ptr(dk), _ptr(dv), B, Lq, d, v.shape[2],
                                     _stream()), "pg_linear_attn_bwd")


@_device_guarded
def cast_multi(src_ptrs, dst_ptrs, numel, chunks, n_chunks, chunk_elems):
    _check(load().pg_cast_multi_bf16(_ptr(src_ptrs), _ptr(dst_ptrs), _ptr(numel), _ptr(chunks), n_chunks, chunk_elems,
                                     _stream()), "pg_cast_multi_bf16")


@_device_guarded
def grad_sqnorm(grad_ptrs, numel, chunks, n_chunks, chunk_elems, partials):
    _check(load().pg_grad_sqnorm(_ptr(grad_ptrs), _ptr(numel), _ptr(chunks), n_chunks, chunk_elems, _ptr(partials),
                                 _stream()), "pg_grad_sqnorm")


@_device_guarded
def adam_step(param_ptrs, grad_ptrs, m_ptrs, v_ptrs, numel, chunks, n_chunks, chunk_elems, partials, max_norm, skip_above,
              lr, beta1, beta2, eps, step, norm_out):
    _check(load().pg_adam_step(_ptr(param_ptrs), _ptr(grad_ptrs), _ptr(m_ptrs), _ptr(v_ptrs), _ptr(numel), _ptr(chunks),
                               n_chunks, chunk_elems, _ptr(partials), max_norm, skip_above, lr, beta1, beta2, eps, step,
                               _ptr(norm_out), _stream()), "pg_adam_step")


def _int_array(vals):
    arr = (ctypes.c_int * len(vals))(*[int(v) for v in vals])
    return arr


@_device_guarded
def tap_gather(x_pm, N, H, W, C, taps, act, out):
    """x_pm: [P, >=C] bf16; taps: list of (dy, dx); out: [P, T*C] bf16 contiguous."""
    lib = load()
    p, ld = _pm(x_pm)
    dy, dx = _int_array([t[0] for t in taps]), _int_array([t[1] for t in taps])
    assert out.is_contiguous() and out.dtype == torch.bfloat16 and x_pm.dtype == torch.bfloat16
    _check(lib.pg_tap_gather(p, ld, N, H, W, C, len(taps), ctypes.cast(dy, ctypes.c_void_p), ctypes.cast(dx, ctypes.c_void_p),
                             act, _ptr(out), _stream()), "pg_tap_gather")


@_device_guarded
def tap_scatter(dxcat, N, H, W, C, taps, act, x_pre, dx_f32=None, dx_bf16=None):
    lib = load()
    assert dxcat.is_contiguous() and dxcat.dtype == torch.bfloat16
    dy, dx = _int_array([t[0] for t in taps]), _int_array([t[1] for t in taps])
    pre_p, pre_ld = (None, 0) if x_pre is None else _pm(x_pre)
    tgt = dx_f32 if dx_f32 is not None else dx_bf16
    _, ld_dx = _pm(tgt)
    if dx_f32 is not None and dx_bf16 is not None:
        assert dx_bf16.stride(0) == ld_dx
    _check(lib.pg_tap_scatter(_ptr(dxcat), N, H, W, C, len(taps), ctypes.cast(dy, ctypes.c_void_p),
                              ctypes.cast(dx, ctypes.c_void_p), act, pre_p, pre_ld, _ptr(dx_f32), _ptr(dx_bf16), ld_dx,
                              _stream()), "pg_tap_scatter")


@_device_guarded
def attn_decode(q, k_new, v_new, k_cache, v_cache, o, pos_dev, N, S, H, dk, dv, strict, dk_true=None):
    """One new position per image against the KV caches (see pg_attn_decode); pos_dev: int32 device scalar."""
    lib = load()
    (qp, ldq), (knp, ldkn), (vnp, ldvn) = _pm(q), _pm(k_new), _pm(v_new)
    (kcp, ldkc), (vcp, ldvc), (op, ldo) = _pm(k_cache), _pm(v_cache), _pm(o)
    assert pos_dev.dtype == torch.int32 and pos_dev.is_cuda
    scale = 1.0 / math.sqrt(dk_true or dk)
    _check(lib.pg_attn_decode(qp, ldq, knp, ldkn, vnp, ldvn, kcp, ldkc, vcp, ldvc, op, ldo, _ptr(pos_dev), N, S, H, dk, dv,
                              scale, int(strict), _stream()), "pg_attn_decode")
