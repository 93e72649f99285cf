"""Parity of the CUDA path with the reference — the first gate (task §③).

Checker = the oracle (oracle/reference_path.py, pinned bit-exact to the live reference in tests/test_oracle.py)
and the committed golden fixtures (outputs of the reference itself).  The product modules are driven through
their public Module API (the drop-in boundary); everything underneath is the C ABI.

Tolerance (BASELINE.json north_star): 1e-2 for the bf16 tensor-core path, 1e-3 where a module computes in
fp32 end to end (LayerNorm, GatedActivation, small-Cin CausalConv2d), both relative to max(1, max|ref|).
"""

import os

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_BF16, TOL_F32 = 1e-2, 1e-3
# Gradients of the BCE recipe loss inherit the (in-tolerance) logit error through sigmoid' and, for bias terms, sum it
# over every pixel; they are checked at 2.5e-2 end to end, and at 1e-2 as fixed-cotangent VJPs (the *_match_oracle tests).
TOL_GRAD_E2E = 2.5e-2


def dev():
    return torch.device("cuda:0")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def check(name, got, ref, tol):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    bound = tol * max(1.0, ref.abs().max().item())
    err = (got - ref).abs().max().item()
    assert err <= bound and not torch.isnan(got).any(), f"{name}: max err {err:.3e} > {bound:.3e} (|ref|max {ref.abs().max():.3e})"
    return err


@pytest.fixture(scope="module")
def pg():
    import pytorch_generative_b200 as pkg
    from pytorch_generative_b200 import models, nn  # noqa: F401

    return pkg


# --------------------------------------------------------------------------------------------------
# nn blocks against the reference fixtures
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["conv3x3A", "conv3x3B", "conv7x7A", "conv3x5B"])
def test_causal_conv2d_matches_reference(pg, tag):
    f = load("nn_blocks.pt")[tag]
    cout, cin, kh, kw = f["weight_before"].shape
    m = pg.nn.CausalConv2d(f["mask_center"], in_channels=cin, out_channels=cout, kernel_size=(kh, kw),
                           padding=f["padding"]).to(dev())
    assert torch.equal(m.mask.cpu(), f["mask"])
    with torch.no_grad():
        m.weight.copy_(f["weight_before"])
        m.bias.copy_(f["bias"])
    x = f["x"].to(dev()).requires_grad_(True)
    y = m(x)
    assert y.is_contiguous()
    assert torch.equal(m.weight.detach().cpu(), f["weight_after"]), "masked taps must be zeroed in place"
    y.backward(f["dy"].to(dev()))
    check(tag + " y", y, f["y"], TOL_F32)
    check(tag + " dx", x.grad, f["dx"], TOL_F32)
    check(tag + " dw", m.weight.grad, f["dw"], TOL_F32)  # dense: masked taps get gradient too
    check(tag + " db", m.bias.grad, f["db"], TOL_F32)


@pytest.mark.parametrize("tag,act", [("gated_tanh", torch.tanh), ("gated_identity", torch.nn.Identity())])
def test_gated_activation_matches_reference(pg, tag, act):
    f = load("nn_blocks.pt")[tag]
    m = pg.nn.GatedActivation(activation_fn=act)
    x = f["x"].to(dev()).requires_grad_(True)
    y = m(x)
    y.backward(f["dy"].to(dev()))
    check(tag + " y", y, f["y"], TOL_F32)
    check(tag + " dx", x.grad, f["dx"], TOL_F32)
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 3, 2, 2, device=dev()))


def test_layernorm_matches_reference(pg):
    f = load("nn_blocks.pt")["layernorm"]
    m = pg.nn.NCHWLayerNorm(f["gamma"].numel()).to(dev())
    with torch.no_grad():
        m.weight.copy_(f["gamma"])
        m.bias.copy_(f["beta"])
    x = f["x"].to(dev()).requires_grad_(True)
    y = m(x)
    y.backward(f["dy"].to(dev()))
    check("ln y", y, f["y"], TOL_F32)
    check("ln dx", x.grad, f["dx"], TOL_F32)
    check("ln dgamma", m.weight.grad, f["dgamma"], TOL_F32)
    check("ln dbeta", m.bias.grad, f["dbeta"], TOL_F32)


@pytest.mark.parametrize("tag", ["attn_causal_mh", "attn_strict_extra", "attn_defaults"])
def test_causal_attention_matches_reference(pg, tag):
    f = load("nn_blocks.pt")[tag]
    m = pg.nn.CausalAttention(**f["kwargs"]).to(dev())
    m.load_state_dict(f["state"])
    x = f["x"].to(dev()).requires_grad_(True)
    extra = None if f["extra"] is None else f["extra"].to(dev()).requires_grad_(True)
    y = m(x, extra) if extra is not None else m(x)
    y.backward(f["dy"].to(dev()))
    check(tag + " y", y, f["y"], TOL_BF16)
    check(tag + " dx", x.grad, f["grads"]["x"], TOL_BF16)
    if extra is not None:
        check(tag + " dextra", extra.grad, f["grads"]["extra"], TOL_BF16)
    for name, p in m.named_parameters():
        check(f"{tag} d{name}", p.grad, f["grads"][name], TOL_BF16)
    if f["kwargs"].get("mask_center"):
        # strict mask: position 0 attends to nothing -> exactly the projection bias
        check(tag + " first pixel", y[:, :, 0, 0], m._proj.bias.detach().expand(y.shape[0], -1), 2e-3)


@pytest.mark.parametrize("tag", ["one_head", "two_heads"])
def test_linear_causal_attention_matches_reference(pg, tag):
    """LinearCausalAttention (reference nn/attention.py:209-275) against the reference's own outputs and gradients: the
    1x1 projections run on the fp32 direct kernel at these channel counts, the sequential numerator on pg_linear_attn_*."""
    f = load("nn_linear_attention.pt")[tag]
    m = pg.nn.LinearCausalAttention(**f["kwargs"]).to(dev())
    m.load_state_dict(f["state"])
    x = f["x"].to(dev()).requires_grad_(True)
    y = m(x)
    y.backward(f["dy"].to(dev()))
    check(tag + " y", y, f["y"], TOL_F32)
    check(tag + " dx", x.grad, f["grads"]["x"], TOL_F32)
    for name, p in m.named_parameters():
        check(f"{tag} d{name}", p.grad, f["grads"][name], TOL_F32)


def test_linear_causal_attention_long_sequence_matches_oracle(pg):
    """32 x 32 pixels, 4 heads of 16 -> 32 channels: the scan kernels across many staging blocks, against the oracle."""
    from oracle import reference_path as O

    torch.manual_seed(2)
    m = pg.nn.LinearCausalAttention(in_channels=16, n_heads=4, embed_channels=64, out_channels=128)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 32, 32, generator=g) * 0.5
    dy = torch.randn(2, 128, 32, 32, generator=g)
    pt = O.trainable({k: v.detach().clone() for k, v in m.state_dict().items()})
    xr = x.clone().requires_grad_(True)
    yr = O.linear_causal_attention(xr, pt, "", 4, 64, 128)
    yr.backward(dy)
    m = m.to(dev())
    xd = x.to(dev()).requires_grad_(True)
    y = m(xd)
    y.backward(dy.to(dev()))
    check("linear attn y", y, yr, TOL_F32)
    check("linear attn dx", xd.grad, xr.grad, TOL_F32)
    for name, p in m.named_parameters():
        check("linear attn d" + name, p.grad, pt[name].grad, TOL_F32)


def test_positional_encoding_bit_identical(pg):
    f = load("nn_blocks.pt")["posenc"]
    assert torch.equal(pg.nn.image_positional_encoding(f["shape"]), f["value"])


# --------------------------------------------------------------------------------------------------
# ImageGPT against the reference fixture and the oracle
# --------------------------------------------------------------------------------------------------
def _loss(x, logits):
    """The recipes' loss through the fused B200 kernel (checked against torch's BCE in test_recipe_loss_kernel)."""
    from pytorch_generative_b200 import losses

    return losses.bce_with_logits_sum_mean(logits, x)


def test_recipe_loss_kernel(pg):
    from pytorch_generative_b200 import losses

    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(6, 3, 16, 16, generator=g) * 3).to(dev()).requires_grad_(True)
    x = torch.rand(6, 3, 16, 16, generator=g).to(dev())
    loss = losses.bce_with_logits_sum_mean(logits, x)
    (loss * 1.7).backward()
    lr = logits.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.binary_cross_entropy_with_logits(lr.reshape(6, -1), x.reshape(6, -1), reduction="none").sum(1).mean()
    (ref * 1.7).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    check("dlogits", logits.grad, lr.grad, 1e-5)


def _build(pg, cls, cfg, state, sample_fn=None):
    m = getattr(pg.models, cls)(sample_fn=sample_fn, **cfg)
    m.load_state_dict(state)
    return m.to(dev())


MODEL_FIXTURES = ["image_gpt", "pixel_cnn", "gated_pixel_cnn", "pixel_snail"]


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_model_matches_reference_fixture(pg, name):
    """Logits, recipe loss, every parameter gradient and the in-place weight masking against outputs of the
    unmodified reference (tests/golden/make_golden.py)."""
    fx = load(f"model_{name}.pt")
    m = _build(pg, fx["cls"], fx["cfg"], fx["state_before"])
    x = fx["x"].to(dev())
    logits = m(x)
    assert logits.is_contiguous() and logits.shape == fx["logits"].shape
    loss = _loss(x, logits)
    loss.backward()
    check("logits", logits, fx["logits"], TOL_BF16)
    assert abs(loss.item() - fx["loss"].item()) <= TOL_BF16 * abs(fx["loss"].item())
    for pname, p in m.named_parameters():
        if pname not in fx["grads"]:  # parameters the reference's graph never reaches (last layer's unused streams)
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, pname
            continue
        assert p.grad is not None, pname
        check("d" + pname, p.grad, fx["grads"][pname], TOL_GRAD_E2E)
    # state-dict round trip keeps the reference's keys, incl. the dynamic shape buffers and the masked weights
    sd = m.state_dict()
    assert {"_c", "_h", "_w"} <= set(sd) and int(sd["_h"]) == x.shape[2]
    assert set(sd) == set(fx["state_after"])
    for k, v in fx["state_after"].items():
        if k.endswith("weight") and (k[: -len("weight")] + "mask") in fx["state_after"]:
            assert torch.equal(sd[k].cpu(), v), f"{k}: masked taps must be zeroed in place like the reference"


@pytest.mark.parametrize("cfg,shape", [
    (dict(in_channels=1, out_channels=1, in_size=28, n_transformer_blocks=8, n_attention_heads=4,
          n_embedding_channels=64), (2, 1, 28, 28)),                                     # BASELINE config C2
    (dict(in_channels=3, out_channels=3, in_size=32, n_transformer_blocks=2, n_attention_heads=8,
          n_embedding_channels=512), (2, 3, 32, 32)),                                    # C5 block geometry
])
def test_image_gpt_matches_oracle(pg, cfg, shape):
    from oracle import reference_path as O

    torch.manual_seed(0)
    m = pg.models.ImageGPT(**cfg)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = (torch.bernoulli(torch.full(shape, 0.5), generator=g) if shape[1] == 1
         else torch.randint(0, 256, shape, generator=g).float() / 255)
    # Forward + recipe loss against the oracle; gradients are compared as a VJP with a FIXED cotangent G
    # (loss' = <logits, G>), which isolates the backward arithmetic from the sigmoid's amplification of the
    # (in-tolerance) forward error — at 8+ blocks the reference's doubling residual makes BCE-driven bias
    # sums cancellation-dominated (measured: up to 35 % on `_out.bias` from a 1 % logit error).
    pt = O.trainable(state)
    ref_logits = O.forward("image_gpt", pt, x, cfg)
    ref_loss = O.recipe_loss(x, ref_logits).detach()
    G = torch.randn(ref_logits.shape, generator=g) / ref_logits[0].numel()
    (ref_logits * G).sum().backward()
    ref_grads = {k: v.grad for k, v in pt.items() if v.requires_grad and v.grad is not None}
    ref_logits = ref_logits.detach()
    m = m.to(dev())
    xd = x.to(dev())
    logits = m(xd)
    loss = _loss(xd, logits)
    (logits * G.to(dev())).sum().backward()
    check("logits", logits, ref_logits, TOL_BF16)
    assert abs(loss.item() - ref_loss.item()) <= TOL_BF16 * abs(ref_loss.item())
    report, worst = [], 0.0
    for name, p in m.named_parameters():
        g, r = p.grad.detach().float().cpu(), ref_grads[name]
        e_max = (g - r).abs().max().item() / max(1.0, r.abs().max().item())
        e_l2 = ((g - r).norm() / r.norm().clamp_min(1e-30)).item()
        report.append(f"{name:40s} max-rel {e_max:.3e}  l2-rel {e_l2:.3e}  |ref|max {r.abs().max().item():.3e}")
        worst = max(worst, e_max)
    print("\n".join(report))
    assert worst <= TOL_BF16, "gradient parity:\n" + "\n".join(report)


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_sampling_follows_reference_raster_order(pg, name):
    """Same pre-drawn uniforms, consumed in raster order: pixels must be identical to the reference's sample
    except, at most, from a knife-edge draw (|u - p| within the bf16 tolerance) onwards."""
    from oracle import reference_path as O

    fx = load(f"model_{name}.pt")
    u = list(fx["sample_uniforms"])
    m = _build(pg, fx["cls"], fx["cfg"], fx["state_before"], sample_fn=O.uniform_sample_fn(u))
    m(fx["x"].to(dev()))  # registers _c/_h/_w like the reference
    n, c, h, w = fx["x"].shape
    got = m.sample(n_samples=n).cpu()
    ref = fx["sample"]
    assert got.shape == ref.shape and set(got.unique().tolist()) <= {0.0, 1.0}
    if not torch.equal(got, ref):
        diff = (got != ref).any(dim=1).any(dim=0)  # [h, w]
        first = diff.flatten().nonzero()[0].item()
        r, col = divmod(first, w)
        canvas = ref.clone()
        canvas.view(n, c, -1)[:, :, first:] = -1
        p_ref = torch.sigmoid(O.forward(name, fx["state_before"], canvas, fx["cfg"])[:, :, r, col])
        margin = (u[first] - p_ref).abs().min().item()
        assert margin < 2e-2, f"samples diverge at pixel ({r},{col}) without a knife-edge draw (margin {margin:.3e})"
    # conditional sampling leaves given pixels untouched (reference models/tests.py:92-95)
    m._sample_fn = O.uniform_sample_fn(u)
    cs = m.sample(conditioned_on=fx["cond"].to(dev())).cpu()
    assert torch.equal(cs[:, :, : h // 2], fx["cond"][:, :, : h // 2])


@pytest.mark.parametrize("cls,cfg,shape", [
    ("PixelCNN", dict(in_channels=1, out_channels=1, n_residual=3, residual_channels=16, head_channels=32), (3, 1, 28, 28)),
    ("PixelCNN", dict(in_channels=3, out_channels=3, n_residual=2, residual_channels=32, head_channels=16), (2, 3, 8, 16)),
    ("PixelSNAIL", dict(in_channels=3, out_channels=3, n_channels=64, n_pixel_snail_blocks=2, n_residual_blocks=2,
                        attention_key_channels=16, attention_value_channels=32), (2, 3, 16, 16)),
    ("GatedPixelCNN", dict(in_channels=1, out_channels=1, n_gated=3, gated_channels=32, head_channels=16), (3, 1, 28, 28)),
    ("GatedPixelCNN", dict(in_channels=3, out_channels=3, n_gated=2, gated_channels=64, head_channels=32), (2, 3, 8, 16)),
    ("GatedPixelCNN", dict(in_channels=1, out_channels=1, n_gated=0, gated_channels=16, head_channels=8), (2, 1, 6, 5)),
    ("PixelSNAIL", dict(in_channels=1, out_channels=1, n_channels=32, n_pixel_snail_blocks=1, n_residual_blocks=1,
                        attention_key_channels=4, attention_value_channels=128), (4, 1, 28, 28)),
])
def test_incremental_sampler_logits_match_the_full_forward(pg, cls, cfg, shape):
    """Teacher-forced sampling: with every pixel given (conditioned_on >= 0) `sample()` still evaluates each pixel's
    logits on the line buffers / K/V caches; they must equal the full forward's logits of the same image.  Run twice:
    the second call replays the captured per-pixel graph on reset caches and re-packed weights."""
    torch.manual_seed(7)
    m = getattr(pg.models, cls)(**cfg).to(dev())
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.5)
    x = torch.bernoulli(torch.full(shape, 0.5)).to(dev())
    with torch.no_grad():
        ref = m(x)
    n, c, h, w = shape
    assert m._incremental_ok(x)
    for rep in range(2):
        seen = []
        m._sample_fn = lambda logits: (seen.append(logits.detach().clone()), torch.zeros_like(logits))[1]
        out = m.sample(conditioned_on=x)
        assert torch.equal(out, x)
        got = torch.stack(seen, dim=-1).view(n, c, h, w)
        check(f"incremental logits (call {rep})", got, ref, TOL_BF16)
    assert m._pixel_states and all(st["graph"] for st in m._pixel_states.values()), "per-pixel program was not graph-captured"


def test_incremental_sampler_falls_back_beyond_its_row_limit(pg):
    m = pg.models.PixelCNN(in_channels=1, out_channels=1, n_residual=1, residual_channels=8, head_channels=8).to(dev())
    x = torch.bernoulli(torch.full((33, 1, 4, 4), 0.5)).to(dev())
    m(x)
    assert not m._incremental_ok(x)
    assert torch.equal(m.sample(conditioned_on=x), x)


@pytest.mark.parametrize("name,cls,cfg,shape", [
    ("pixel_cnn", "PixelCNN", dict(in_channels=1, out_channels=1, n_residual=3, residual_channels=32, head_channels=16),
     (2, 1, 28, 28)),
    ("gated_pixel_cnn", "GatedPixelCNN", dict(in_channels=3, out_channels=3, n_gated=3, gated_channels=32,
                                              head_channels=16), (2, 3, 32, 32)),
    ("pixel_snail", "PixelSNAIL", dict(in_channels=3, out_channels=3, n_channels=64, n_pixel_snail_blocks=2,
                                       n_residual_blocks=2, attention_key_channels=16, attention_value_channels=32),
     (2, 3, 32, 32)),
    # 64-channel variants: wide enough for the fused pixel-major stacks (TMA tap-loop convolutions, nn/pm.py)
    ("gated_pixel_cnn", "GatedPixelCNN", dict(in_channels=3, out_channels=3, n_gated=3, gated_channels=64,
                                              head_channels=32), (2, 3, 32, 32)),
    ("gated_pixel_cnn", "GatedPixelCNN", dict(in_channels=1, out_channels=1, n_gated=2, gated_channels=64,
                                              head_channels=16), (3, 1, 16, 32)),
    # 28 x 28 images: no TMA tap loop (W does not divide 64) -> the module path with gathered taps
    ("pixel_snail", "PixelSNAIL", dict(in_channels=1, out_channels=1, n_channels=64, n_pixel_snail_blocks=1,
                                       n_residual_blocks=1, attention_key_channels=8, attention_value_channels=32),
     (2, 1, 28, 28)),
])
def test_conv_models_match_oracle(pg, name, cls, cfg, shape):
    """Mid-size PixelCNN / GatedPixelCNN / PixelSNAIL (tap-list convs on the tensor-core GEMM, wide channels) against
    the oracle: logits, recipe loss, and a fixed-cotangent VJP for every parameter."""
    from oracle import reference_path as O

    torch.manual_seed(0)
    m = getattr(pg.models, cls)(**cfg)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.02)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = (torch.bernoulli(torch.full(shape, 0.5), generator=g) if shape[1] == 1
         else torch.randint(0, 256, shape, generator=g).float() / 255)
    pt = O.trainable(state)
    ref_logits = O.forward(name, pt, x, cfg)
    ref_loss = O.recipe_loss(x, ref_logits).detach()
    G = torch.randn(ref_logits.shape, generator=g) / ref_logits[0].numel()
    (ref_logits * G).sum().backward()
    ref_grads = {k: v.grad for k, v in pt.items() if v.requires_grad and v.grad is not None}
    m = m.to(dev())
    xd = x.to(dev())
    logits = m(xd)
    loss = _loss(xd, logits)
    (logits * G.to(dev())).sum().backward()
    check("logits", logits, ref_logits.detach(), TOL_BF16)
    assert abs(loss.item() - ref_loss.item()) <= TOL_BF16 * abs(ref_loss.item())
    report, worst = [], 0.0
    for pname, p in m.named_parameters():
        if pname not in ref_grads:
            continue
        gq, r = p.grad.detach().float().cpu(), ref_grads[pname]
        e = (gq - r).abs().max().item() / max(1.0, r.abs().max().item())
        report.append(f"{pname:50s} max-rel {e:.3e} |ref|max {r.abs().max().item():.3e}")
        worst = max(worst, e)
    assert worst <= TOL_BF16, "gradient parity:\n" + "\n".join(report)


def test_wide_causal_conv2d_matches_oracle(pg):
    """CausalConv2d with wide channels runs as a tap list on the GEMM (dense weight gradient over all 9 taps)."""
    from oracle import reference_path as O

    g = torch.Generator().manual_seed(3)
    for mask_center in (False, True):
        m = pg.nn.CausalConv2d(mask_center, in_channels=32, out_channels=48, kernel_size=3, padding=1)
        x = torch.randn(2, 32, 12, 10, generator=g)
        dy = torch.randn(2, 48, 12, 10, generator=g)
        w0, b0 = m.weight.detach().clone(), m.bias.detach().clone()
        xr, wr, br = x.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        yr, w_masked = O.causal_conv2d(xr, wr, br, mask_center, 1)
        gx, gw, gb = torch.autograd.grad(yr, [xr, wr, br], dy)
        m = m.to(dev())
        xd = x.to(dev()).requires_grad_(True)
        y = m(xd)
        y.backward(dy.to(dev()))
        assert torch.equal(m.weight.detach().cpu(), w_masked.detach())
        check("wide conv y", y, yr, TOL_BF16)
        check("wide conv dx", xd.grad, gx, TOL_BF16)
        check("wide conv dw", m.weight.grad, gw, TOL_BF16)
        check("wide conv db", m.bias.grad, gb, TOL_BF16)
        assert (m.weight.grad.cpu() * (1 - m.mask.cpu())).abs().sum() > 0  # masked taps receive gradient (dense wgrad)


@pytest.mark.parametrize("name", ["image_gpt", "pixel_cnn", "gated_pixel_cnn"])
def test_row_truncated_forward_is_bit_identical(pg, name):
    """sample() evaluates pixel (r, c) on the top r+1 rows of the canvas: the logits of those rows must be bit-for-bit
    the ones of the full forward (row causality + row-independent kernels), so the sampling order/values are unchanged."""
    fx = load(f"model_{name}.pt")
    m = _build(pg, fx["cls"], fx["cfg"], fx["state_before"]).eval()
    assert m._row_truncated_sampling
    x = fx["x"].to(dev())
    with torch.no_grad():
        full = m(x)
        for r in (0, 3, 6):
            part = m(x[:, :, : r + 1].contiguous())
            assert torch.equal(part, full[:, :, : r + 1]), (name, r)


def test_transformer_block_standalone_forward_matches_oracle(pg):
    """`TransformerBlock(x)` on its own (reference image_gpt.py:50-52: h = x + attn(ln1(x)); h + mlp(ln2(h)))."""
    from oracle import reference_path as O
    from pytorch_generative_b200.models.image_gpt import TransformerBlock

    torch.manual_seed(4)
    blk = TransformerBlock(n_channels=64, n_attention_heads=4)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in blk.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    x = torch.randn(2, 64, 8, 16, generator=g)
    p = {k: v.detach() for k, v in blk.state_dict().items()}
    a = O.nchw_layer_norm(x, p["_ln1.weight"], p["_ln1.bias"])
    h = x + O.causal_attention(a, p, "_attn.", 4, 64, 64, False)
    m = O.nchw_layer_norm(h, p["_ln2.weight"], p["_ln2.bias"])
    ref = h + torch.nn.functional.conv2d(torch.nn.functional.gelu(torch.nn.functional.conv2d(m, p["_out.0.weight"], p["_out.0.bias"])),
                                         p["_out.2.weight"], p["_out.2.bias"])
    y = blk.to(dev())(x.to(dev()))
    check("transformer block", y, ref, TOL_BF16)


def test_image_gpt_eval_forward_keeps_no_activations(pg):
    """Under torch.no_grad() (eval, sampling) the fused stack must not retain the per-block activations, and a second
    backward through a consumed graph raises a clear error instead of a TypeError."""
    m = pg.models.ImageGPT(in_channels=1, out_channels=1, in_size=8, n_transformer_blocks=2, n_attention_heads=2,
                           n_embedding_channels=32).to(dev())
    x = torch.rand(2, 1, 8, 8, device=dev())
    with torch.no_grad():
        y = m(x)
    assert y.grad_fn is None and not y.requires_grad
    y = m(x)
    y.sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="already consumed"):
        y.sum().backward()
