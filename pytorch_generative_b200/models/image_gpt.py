"""ImageGPT on the B200 path — API of reference models/autoregressive/image_gpt.py:21-109.

Module tree, parameter names and shapes are the reference's (`_pos`, `_input`, `_transformer.{i}.{_ln1,_ln2,
_attn.{_q,_kv,_proj},_out.{0,2}}`, `_ln`, `_out`), so checkpoints are interchangeable.  `forward` does not walk
that tree: the whole stack runs as ONE autograd node over pixel-major tensors —

    stream x (fp32 [P, C])  --LN-->  bf16  --GEMM(q|k|v)-->  causal attention  --GEMM(proj)+x--> h (fp32)
    h --LN--> bf16 --GEMM(4C)+GELU--> bf16 --GEMM(C) + x + h--> next stream      (x <- x + h + mlp, the
                                                                                  reference's double residual)

with every bias / activation / residual folded into a GEMM epilogue, the residual stream and all summed
gradients kept in fp32, and bf16 used only for tensor-core operands.
"""

import os
import warnings

import torch
from torch import nn

from .. import _lib as L
from .. import nn as pg_nn
from .. import ops
from ..nn.modules import pack_qkv_weights
from . import base

F32, BF16 = torch.float32, torch.bfloat16
PARAMS_PER_BLOCK = 14


class TransformerBlock(nn.Module):
    """Holds the parameters of one block (reference image_gpt.py:21-52); standalone `forward` composes the
    drop-in nn modules, the fused model path reads the parameters directly."""

    def __init__(self, n_channels, n_attention_heads):
        super().__init__()
        self._ln1 = pg_nn.NCHWLayerNorm(n_channels)
        self._ln2 = pg_nn.NCHWLayerNorm(n_channels)
        self._attn = pg_nn.CausalAttention(in_channels=n_channels, n_heads=n_attention_heads,
                                           embed_channels=n_channels, out_channels=n_channels)
        self._out = nn.Sequential(
            nn.Conv2d(in_channels=n_channels, out_channels=4 * n_channels, kernel_size=1),
            nn.GELU(),
            nn.Conv2d(in_channels=4 * n_channels, out_channels=n_channels, kernel_size=1),
        )

    def forward(self, x):
        """Standalone use of one block on an NCHW tensor (reference image_gpt.py:50-52), composed of the drop-in
        modules; `ImageGPT.forward` does not come through here (it runs the fused stack)."""
        from ..nn.tapconv import tap_conv2d

        h = x + self._attn(self._ln1(x))
        t = tap_conv2d(self._ln2(h), self._out[0].weight, self._out[0].bias, (0, 0), post_act=L.ACT_GELU)
        return h + tap_conv2d(t, self._out[2].weight, self._out[2].bias, (0, 0))

    def flat_params(self):
        a = self._attn
        return [self._ln1.weight, self._ln1.bias, a._q.weight, a._q.bias, a._kv.weight, a._kv.bias, a._proj.weight,
                a._proj.bias, self._ln2.weight, self._ln2.bias, self._out[0].weight, self._out[0].bias,
                self._out[2].weight, self._out[2].bias]


class _ImageGPTStack(torch.autograd.Function):
    """forward(x_nchw, params...) -> logits_nchw; one node for the whole network."""

    @staticmethod
    def forward(ctx, x, n_heads, eps, packed, opts, *params):
        pos, in_w, in_b = params[0], params[1], params[2]
        n_blocks = (len(params) - 7) // PARAMS_PER_BLOCK
        ln_w, ln_b, out_w, out_b = params[-4:]
        n, cin, h, w = x.shape
        S, P, H = h * w, n * h * w, n_heads
        C = in_w.shape[0]
        # needs_input_grad ignores torch.no_grad(): eval / sampling must not retain every block's activations
        keep = any(ctx.needs_input_grad) and opts["grad"]

        x_in = (x + pos[:, :, : h, : w]).contiguous()  # sampling evaluates the top rows of the canvas only
        xs = torch.empty(P, C, dtype=F32, device=x.device)
        L.conv_small_fwd(x_in, in_w.detach().contiguous(), in_b.detach(), (in_w.shape[2] // 2, in_w.shape[3] // 2),
                         out_f32=xs)
        saved = []
        for b in range(n_blocks):
            (ln1_w, ln1_b, q_w, q_b, kv_w, kv_b, p_w, p_b, ln2_w, ln2_b, f1_w, f1_b, f2_w,
             f2_b) = params[3 + b * PARAMS_PER_BLOCK: 3 + (b + 1) * PARAMS_PER_BLOCK]
            pk = packed["blocks"][b]
            wqkv, bqkv, meta = pk["wqkv"], pk["bqkv"], pk["meta"]
            dv_slot, slot = meta["dv_slot"], ops.HEAD_SLOT
            a1, _, mean1, rstd1 = ops.layernorm_fwd(xs, ln1_w.detach(), ln1_b.detach(), eps)
            qkv, _, _ = ops.linear_fwd(a1, wqkv, bqkv)
            q, k, v = qkv[:, : H * slot], qkv[:, H * slot: 2 * H * slot], qkv[:, 2 * H * slot:]
            o, lse = ops.attn_fwd(q, k, v, n, S, H, meta["dk"], dv_slot, False)
            wp, cols_v = pk["wp"], pk["cols_v"]
            _, _, hres = ops.linear_fwd(o, wp, p_b.detach(), res0=xs, want_bf16=False, want_f32=True)
            a2, _, mean2, rstd2 = ops.layernorm_fwd(hres, ln2_w.detach(), ln2_b.detach(), eps)
            w1, w2 = pk["w1"], pk["w2"]
            g, u, _ = ops.linear_fwd(a2, w1, f1_b.detach(), act=L.ACT_GELU, want_pre=True, pre_deriv=True)  # u = GELU'(pre)
            _, _, xs_new = ops.linear_fwd(g, w2, f2_b.detach(), res0=xs, res1=hres, want_bf16=False, want_f32=True)
            if keep:
                saved.append(dict(xs=xs, a1=a1, qkv=qkv, o=o, lse=lse, h=hres, a2=a2, u=u, g=g, mean1=mean1, rstd1=rstd1,
                                  mean2=mean2, rstd2=rstd2, wqkv=wqkv, wp=wp, w1=w1, w2=w2, meta=meta, cols_v=cols_v))
            xs = xs_new
        af, _, mean_f, rstd_f = ops.layernorm_fwd(xs, ln_w.detach(), ln_b.detach(), eps)
        cout = out_w.shape[0]
        wo = packed["wo"]
        _, _, logits_pm = ops.linear_fwd(af, wo, out_b.detach(), want_bf16=False, want_f32=True)
        if keep:
            ctx.saved = dict(blocks=saved, x_in=x_in, xs_final=xs, af=af, mean_f=mean_f, rstd_f=rstd_f, wo=wo,
                             params=params, dims=(n, cin, h, w, C, H, cout), eps=eps, hook=opts.get("hook"))
        return ops.pm_to_nchw(logits_pm, n, cout, h, w)

    @staticmethod
    def backward(ctx, dlogits):
        sv = getattr(ctx, "saved", None)
        if sv is None:
            raise RuntimeError("ImageGPT: the activations of this forward were already consumed by a backward pass "
                               "(retain_graph is not supported by the fused stack)")
        params = sv["params"]
        n, cin, h, w, C, H, cout = sv["dims"]
        S, P = h * w, n * h * w
        dev = dlogits.device
        slot = ops.HEAD_SLOT
        n_blocks = len(sv["blocks"])
        grads = [None] * len(params)
        ln_w = params[-4]

        # head: logits = 1x1(LN(x))
        dl = ops.nchw_to_pm(dlogits, BF16, width=ops.round_up(cout, 8))
        grads[-1] = ops.bias_grad(dl[:, :cout])
        dwo = torch.zeros(ops.round_up(cout, 8), C, dtype=F32, device=dev)
        ops.linear_wgrad(dl, sv["af"], dwo)
        grads[-2] = dwo[:cout].reshape(cout, C, 1, 1)
        daf = ops.linear_dgrad(dl[:, :cout], sv["wo"])
        # every LayerNorm backward also emits the column sums of the gradient it writes = the bias gradient of the
        # linear layer that produced its input (fc2 of the block above / the attention projection)
        dx, dx_b, grads[-4], grads[-3], dx_sum = ops.layernorm_bwd(daf, sv["xs_final"], ln_w.detach(), sv["mean_f"],
                                                                    sv["rstd_f"], want_colsum=True)
        del daf

        # one zero-filled fp32 arena for every weight gradient of the stack (the wgrad GEMMs accumulate into it with
        # TMA reduce-adds): a single memset instead of four per block
        qkv_rows = sv["blocks"][0]["wqkv"].shape[0] if n_blocks else 0
        dvs = sv["blocks"][0]["meta"]["dv_slot"] if n_blocks else 0
        per_block = 4 * C * C + 4 * C * C + C * H * dvs + qkv_rows * C
        # ... followed by the small per-block gradients (LayerNorm dgamma / dbeta / column sums, bias gradients of the qkv
        # and fc1 layers), which their kernels accumulate with atomics: they share the one memset too
        per_small = 6 * C + qkv_rows + 4 * C
        arena = torch.zeros(n_blocks * (per_block + per_small), dtype=F32, device=dev)
        small_base = n_blocks * per_block

        def carve_small(b, off, n):
            start = small_base + b * per_small + off
            return arena[start: start + n]

        def carve(b, off, rows, cols):
            start = b * per_block + off
            return arena[start: start + rows * cols].view(rows, cols)

        # data parallelism: each block's slice of the arena is handed to the bucket hook (an asynchronous all-reduce)
        # as soon as its last wgrad GEMM is queued; see parallel.OverlappedGradAverager
        bucket_hook = sv["hook"] if _arena_views_are_grads(sv) else None
        # transformer blocks per all-reduce; 0 = the whole stack in one bucket, issued when block 0's last wgrad is queued
        # (overlaps the input convolution's backward and the small-gradient bucket only).  Finer buckets overlap more but
        # every NCCL kernel that runs next to the GEMMs slows them by more than it hides: 8 x B200, 24 / 6 / 1 blocks per
        # bucket = 8000 / 7855 / 7637 img/s (profiles/r02_bench_multigpu.txt).
        bucket_blocks = int(os.environ.get("PG_DP_BUCKET_BLOCKS", "0"))
        if bucket_blocks <= 0:
            bucket_blocks = n_blocks
        pending = []

        for b in reversed(range(n_blocks)):
            blk = sv["blocks"][b]
            base_i = 3 + b * PARAMS_PER_BLOCK
            (ln1_w, _, q_w, _, kv_w, _, p_w, _, ln2_w, _, f1_w, _, f2_w, _) = params[base_i: base_i + PARAMS_PER_BLOCK]
            meta, dv_slot = blk["meta"], blk["meta"]["dv_slot"]
            # x_new = x + h + fc2(gelu(fc1(ln2(h))))
            grads[base_i + 13] = dx_sum
            dw2 = carve(b, 0, C, 4 * C)
            ops.linear_wgrad(dx_b, blk["g"], dw2)
            grads[base_i + 12] = dw2.view(C, 4 * C, 1, 1)
            du = ops.linear_dgrad(dx_b, blk["w2"], aux=blk["u"], dact=L.ACT_GIVEN)
            # the bias gradient (column sums of du) is reduced by the wgrad launch from the du tiles it stages
            grads[base_i + 11] = carve_small(b, 6 * C + qkv_rows, 4 * C)
            dw1 = carve(b, 4 * C * C, 4 * C, C)
            ops.linear_wgrad(du, blk["a2"], dw1, db_out=grads[base_i + 11])
            grads[base_i + 10] = dw1.view(4 * C, C, 1, 1)
            da2 = ops.linear_dgrad(du, blk["w1"])
            del du
            # h receives: LN2 path + direct (x_new = ... + h)
            dh, dh_b, grads[base_i + 8], grads[base_i + 9], grads[base_i + 7] = ops.layernorm_bwd(
                da2, blk["h"], ln2_w.detach(), blk["mean2"], blk["rstd2"], dres0=dx, want_colsum=True,
                stats=carve_small(b, 3 * C, 3 * C).view(3, C))
            del da2
            # h = x + proj(attn)
            dwp = carve(b, 8 * C * C, C, H * dv_slot)
            ops.linear_wgrad(dh_b, blk["o"], dwp)
            grads[base_i + 6] = (dwp if meta["identity"] else dwp[:, blk["cols_v"]]).reshape(C, C, 1, 1)
            do = ops.linear_dgrad(dh_b, blk["wp"])
            qkv = blk["qkv"]
            q, k, v = qkv[:, : H * slot], qkv[:, H * slot: 2 * H * slot], qkv[:, 2 * H * slot:]
            dqkv = torch.empty_like(qkv)
            ops.attn_bwd(q, k, v, blk["o"], do, blk["lse"], dqkv[:, : H * slot], dqkv[:, H * slot: 2 * H * slot],
                         dqkv[:, 2 * H * slot:], n, S, H, meta["dk"], dv_slot, False)
            del do
            dbqkv = carve_small(b, 6 * C, qkv_rows)
            dwqkv = carve(b, 8 * C * C + C * H * dv_slot, qkv_rows, C)
            ops.linear_wgrad(dqkv, blk["a1"], dwqkv, db_out=dbqkv)
            if meta["identity"]:  # heads fill their slots: plain slices of the fused gradient buffers
                grads[base_i + 2] = dwqkv[:C].view(C, C, 1, 1)
                grads[base_i + 3] = dbqkv[:C]
                grads[base_i + 4] = dwqkv[C:].view(2 * C, C, 1, 1)
                grads[base_i + 5] = dbqkv[C:]
            else:
                rq, rv = meta["rows_q"], meta["rows_v"]
                grads[base_i + 2] = dwqkv[rq].reshape(C, C, 1, 1)
                grads[base_i + 3] = dbqkv[rq]
                grads[base_i + 4] = torch.cat((dwqkv[rq + H * slot], dwqkv[rv + H * slot])).reshape(2 * C, C, 1, 1)
                grads[base_i + 5] = torch.cat((dbqkv[rq + H * slot], dbqkv[rv + H * slot]))
            da1 = ops.linear_dgrad(dqkv, blk["wqkv"])
            del dqkv
            # x receives: LN1 path + direct from h (dh) + direct from x_new (dx)
            dx, dx_b, grads[base_i + 0], grads[base_i + 1], dx_sum = ops.layernorm_bwd(
                da1, blk["xs"], ln1_w.detach(), blk["mean1"], blk["rstd1"], dres0=dx, dres1=dh, want_colsum=True,
                stats=carve_small(b, 0, 3 * C).view(3, C))
            del da1, dh, dh_b
            if bucket_hook is not None and b % bucket_blocks == 0:
                # blocks b .. b + bucket_blocks - 1 are complete: one contiguous slice of the arena
                hi = min(b + bucket_blocks, n_blocks)
                pending.append(bucket_hook(arena[b * per_block: hi * per_block]))
            sv["blocks"][b] = None  # release this block's activations

        in_w = params[1]
        dw_in = torch.zeros_like(in_w)
        db_in = dx_sum  # bias gradient of the input conv = column sums of the stream gradient
        dx_in = torch.empty(n, cin, h, w, dtype=F32, device=dev)
        L.conv_small_bwd(sv["x_in"], in_w.detach().contiguous(), dx, (in_w.shape[2] // 2, in_w.shape[3] // 2), dw=dw_in,
                         dbias=None, dx=dx_in)
        grads[1], grads[2] = dw_in, db_in
        dpos = torch.zeros_like(params[0])
        dpos[:, :, : h, : w] = dx_in.sum(dim=0, keepdim=True)
        grads[0] = dpos
        ctx.saved = None
        for handle in pending:  # the gradients leave this node averaged
            handle.wait()
        return (dx_in if ctx.needs_input_grad[0] else None, None, None, None, None, *grads)


def _arena_views_are_grads(sv):
    """True when every block's weight gradients are plain views of the gradient arena (heads fill their 64-wide
    slots, e.g. 512 channels / 8 heads): only then can the arena slice be averaged in place."""
    return bool(sv["blocks"]) and all(blk["meta"]["identity"] for blk in sv["blocks"])




class ImageGPT(base.AutoregressiveModel):
    """The (convolutional) ImageGPT model — constructor of reference image_gpt.py:64-103."""

    def __init__(self, in_channels=1, out_channels=1, in_size=28, n_transformer_blocks=8, n_attention_heads=4,
                 n_embedding_channels=16, sample_fn=None):
        super().__init__(sample_fn)
        self._pos = nn.Parameter(torch.zeros(1, in_channels, in_size, in_size))
        self._input = pg_nn.CausalConv2d(mask_center=True, in_channels=in_channels,
                                         out_channels=n_embedding_channels, kernel_size=3, padding=1)
        self._transformer = nn.ModuleList(
            TransformerBlock(n_channels=n_embedding_channels, n_attention_heads=n_attention_heads)
            for _ in range(n_transformer_blocks)
        )
        self._ln = pg_nn.NCHWLayerNorm(n_embedding_channels)
        self._out = nn.Conv2d(in_channels=n_embedding_channels, out_channels=out_channels, kernel_size=1)
        self._n_heads = n_attention_heads
        if n_embedding_channels % 8 != 0:
            raise NotImplementedError("ImageGPT: n_embedding_channels must be a multiple of 8 on the B200 path")

    # ------------------------------------------------------------------------------------------------------------
    # bf16 tensor-core copies of the weight matrices.  The fp32 Parameters stay the master weights (reference
    # semantics: the optimizer updates them in place); the copies are rebuilt only when a parameter's version counter
    # has moved (once per optimizer step; never between the forwards of eval / sampling), by ONE multi-tensor cast
    # launch into a fresh bf16 arena (q | kv weights land adjacent, so the fused qkv projection needs no torch.cat) and
    # one concatenation of all the q / kv biases.  A fresh arena per refresh: a backward that is still pending keeps
    # reading the copies its forward used.
    # ------------------------------------------------------------------------------------------------------------
    def _packed_training_weights(self):
        C, H = self._input.weight.shape[0], self._n_heads
        blocks = list(self._transformer)
        mats = [w for blk in blocks for w in (blk._attn._q.weight, blk._attn._kv.weight, blk._attn._proj.weight,
                                              blk._out[0].weight, blk._out[2].weight)] + [self._out.weight]
        biases = [b for blk in blocks for b in (blk._attn._q.bias, blk._attn._kv.bias)]
        sig = (mats[0].data_ptr(), tuple(p._version for p in mats), tuple(p._version for p in biases))
        cache = self.__dict__.setdefault("_wcache", {})
        capturing = mats[0].is_cuda and torch.cuda.is_current_stream_capturing()
        if cache.get("sig") == sig and not capturing:  # inside a CUDA graph the casts must be captured kernels
            return cache["packed"]
        dev = mats[0].device
        cout = self._out.weight.shape[0]
        identity = C // H == ops.HEAD_SLOT and C % 8 == 0
        packed = {"blocks": []}
        if identity and blocks:
            per_block = 12 * C * C
            arena = torch.empty(len(blocks) * per_block + cout * C, dtype=BF16, device=dev)
            views, dsts = [], []
            for b in range(len(blocks)):
                base = b * per_block
                wqkv = arena[base: base + 3 * C * C].view(3 * C, C)
                wp = arena[base + 3 * C * C: base + 4 * C * C].view(C, C)
                w1 = arena[base + 4 * C * C: base + 8 * C * C].view(4 * C, C)
                w2 = arena[base + 8 * C * C: base + 12 * C * C].view(C, 4 * C)
                views.append((wqkv, wp, w1, w2))
                dsts += [wqkv[:C], wqkv[C:], wp, w1, w2]
            wo = arena[len(blocks) * per_block:].view(cout, C)
            dsts.append(wo)
            plan = cache.get("plan")
            if plan is None or plan["src_key"] != tuple(p.data_ptr() for p in mats):
                from .. import optim

                numel = [p.numel() for p in mats]
                chunks = [(t, c) for t, n in enumerate(numel) for c in range((n + optim.CHUNK - 1) // optim.CHUNK)]
                plan = dict(src_key=tuple(p.data_ptr() for p in mats), n_chunks=len(chunks), chunk=optim.CHUNK,
                            numel=torch.tensor(numel, dtype=torch.int64, device=dev),
                            chunks=torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous(),
                            src=torch.tensor([p.data_ptr() for p in mats], dtype=torch.int64, device=dev),
                            host_dst=torch.empty(len(mats), dtype=torch.int64).pin_memory(),
                            dst=torch.empty(len(mats), dtype=torch.int64, device=dev))
                cache["plan"] = plan
            plan["host_dst"].numpy()[:] = [d.data_ptr() for d in dsts]
            plan["dst"].copy_(plan["host_dst"], non_blocking=True)
            L.cast_multi(plan["src"], plan["dst"], plan["numel"], plan["chunks"], plan["n_chunks"], plan["chunk"])
            ball = torch.cat([b.detach() for b in biases])  # [blocks * 3C]: q | kv biases of every block
            meta = dict(dk=ops.HEAD_SLOT, dv=ops.HEAD_SLOT, dv_slot=ops.HEAD_SLOT, rows_q=None, rows_v=None, identity=True)
            for b, (wqkv, wp, w1, w2) in enumerate(views):
                packed["blocks"].append(dict(wqkv=wqkv, bqkv=ball[b * 3 * C: (b + 1) * 3 * C], wp=wp, w1=w1, w2=w2, meta=meta,
                                             cols_v=None))
            packed["wo"], packed["arena"] = wo, arena
        else:  # narrow heads live in zero-padded 64-wide slots: scatter-pack per block
            for blk in blocks:
                a = blk._attn
                wq, bq, wkv, bkv, meta = pack_qkv_weights(a._q.weight, a._q.bias, a._kv.weight, a._kv.bias, H, C, C, C, C)
                if meta["identity"]:
                    wp, cols_v = ops.pack_weight(a._proj.weight), None
                else:
                    cols_v = meta["rows_v"] - H * ops.HEAD_SLOT
                    wp32 = torch.zeros(C, H * meta["dv_slot"], dtype=F32, device=dev)
                    wp32[:, cols_v] = a._proj.weight.detach().reshape(C, -1)
                    wp = ops.to_bf16(wp32)
                packed["blocks"].append(dict(wqkv=torch.cat((wq, wkv)), bqkv=torch.cat((bq, bkv)), wp=wp,
                                             w1=ops.pack_weight(blk._out[0].weight), w2=ops.pack_weight(blk._out[2].weight),
                                             meta=meta, cols_v=cols_v))
            packed["wo"] = ops.pack_weight(self._out.weight)
        if not capturing:
            cache["sig"], cache["packed"] = sig, packed
        return packed

    # ------------------------------------------------------------------------------------------------------------
    # Incremental sampling.  The reference's sample() (models/base.py:97-120) runs a full forward per pixel; the model is
    # exactly causal, so the logits of pixel p only need p's own row through the stack plus the keys / values of the
    # pixels before it.  Per pixel: input conv on the 3x3 window around p, and per block LN -> q|k|v GEMM (M = batch
    # rows) -> pg_attn_decode over the K/V caches -> proj / MLP GEMMs with the same fused epilogues as training.  The
    # step is captured once in a CUDA graph (the position lives in device memory) and replayed for every pixel; the
    # raster order, the `sample_fn` hook and the "only entries < 0 are overwritten" rule are the base class's.
    # ------------------------------------------------------------------------------------------------------------
    _incremental_sampling = True

    def _packed_weights(self):
        C, H = self._input.weight.shape[0], self._n_heads
        blocks = []
        for blk in self._transformer:
            (ln1_w, ln1_b, q_w, q_b, kv_w, kv_b, p_w, p_b, ln2_w, ln2_b, f1_w, f1_b, f2_w, f2_b) = blk.flat_params()
            wq, bq, wkv, bkv, meta = pack_qkv_weights(q_w, q_b, kv_w, kv_b, H, C, C, C, C)
            if meta["identity"]:
                wp = ops.pack_weight(p_w)
            else:
                wp32 = torch.zeros(C, H * meta["dv_slot"], dtype=F32, device=p_w.device)
                wp32[:, meta["rows_v"] - H * ops.HEAD_SLOT] = p_w.detach().reshape(C, -1)
                wp = ops.to_bf16(wp32)
            blocks.append(dict(wqkv=torch.cat((wq, wkv)), bqkv=torch.cat((bq, bkv)).contiguous(), wp=wp,
                               w1=ops.pack_weight(f1_w), w2=ops.pack_weight(f2_w), dk=meta["dk"], dv_slot=meta["dv_slot"]))
        return blocks, ops.pack_weight(self._out.weight)

    def _sampler_step(self, st):
        """One position for every image of the batch: st["patch"] (window of x + pos around the pixel) -> logits."""
        n, C, H, S, eps = st["n"], st["C"], self._n_heads, st["S"], self._ln.eps
        kh, kw = self._input.weight.shape[2:]
        taps_out = torch.empty(n * kh * kw, C, dtype=F32, device=st["patch"].device)
        L.conv_small_fwd(st["patch"], self._input.weight.detach().contiguous(), self._input.bias.detach(),
                         (kh // 2, kw // 2), out_f32=taps_out)
        xs = taps_out.view(n, kh * kw, C)[:, (kh // 2) * kw + kw // 2].contiguous()  # the window's centre pixel
        slot = ops.HEAD_SLOT
        for b, blk in enumerate(self._transformer):
            wb = st["w"][b]
            a1, _, _, _ = ops.layernorm_fwd(xs, blk._ln1.weight.detach(), blk._ln1.bias.detach(), eps)
            qkv, _, _ = ops.linear_fwd(a1, wb["wqkv"], wb["bqkv"], skinny=True)
            q, k, v = qkv[:, : H * slot], qkv[:, H * slot: 2 * H * slot], qkv[:, 2 * H * slot:]
            o = torch.empty(n, H * wb["dv_slot"], dtype=BF16, device=xs.device)
            L.attn_decode(q, k, v, st["kc"][b], st["vc"][b], o, st["pos"], n, S, H, slot, wb["dv_slot"], False,
                          dk_true=wb["dk"])
            _, _, hres = ops.linear_fwd(o, wb["wp"], blk._attn._proj.bias.detach(), res0=xs, want_bf16=False, want_f32=True,
                                        skinny=True)
            a2, _, _, _ = ops.layernorm_fwd(hres, blk._ln2.weight.detach(), blk._ln2.bias.detach(), eps)
            g, _, _ = ops.linear_fwd(a2, wb["w1"], blk._out[0].bias.detach(), act=L.ACT_GELU, skinny=True)
            _, _, xs = ops.linear_fwd(g, wb["w2"], blk._out[2].bias.detach(), res0=xs, res1=hres, want_bf16=False,
                                      want_f32=True, skinny=True)
        af, _, _, _ = ops.layernorm_fwd(xs, self._ln.weight.detach(), self._ln.bias.detach(), eps)
        _, _, logits = ops.linear_fwd(af, st["wo"], self._out.bias.detach(), want_bf16=False, want_f32=True, skinny=True)
        return logits

    def _sampler_state(self, n, c, h, w, device):
        cache = self.__dict__.setdefault("_samplers", {})
        key = (n, c, h, w, str(device))
        blocks, wo = self._packed_weights()
        st = cache.get(key)
        if st is None:
            C, H, S = self._input.weight.shape[0], self._n_heads, h * w
            kh, kw = self._input.weight.shape[2:]
            st = dict(n=n, C=C, S=S, w=blocks, wo=wo, graph=None,
                      patch=torch.zeros(n, c, kh, kw, dtype=F32, device=device),
                      pos=torch.zeros(1, dtype=torch.int32, device=device),
                      kc=[torch.zeros(n * S, H * ops.HEAD_SLOT, dtype=BF16, device=device) for _ in blocks],
                      vc=[torch.zeros(n * S, H * bw["dv_slot"], dtype=BF16, device=device) for bw in blocks])
            cache[key] = st
        else:  # refresh the packed weights in place: a captured graph keeps reading the same buffers
            for old, new in zip(st["w"], blocks):
                for k2 in ("wqkv", "bqkv", "wp", "w1", "w2"):
                    old[k2].copy_(new[k2])
            st["wo"].copy_(wo)
        return st

    @torch.no_grad()
    def sample(self, n_samples=None, conditioned_on=None):
        canvas = self._start_canvas(n_samples, conditioned_on)
        n, c, h, w = canvas.shape
        if not (self._incremental_sampling and canvas.is_cuda and h * w <= 1024 and h <= self._pos.shape[2]
                and w <= self._pos.shape[3]):
            return super().sample(conditioned_on=canvas)
        self._input.weight.data *= self._input.mask
        st = self._sampler_state(n, c, h, w, canvas.device)
        kh, kw = self._input.weight.shape[2:]
        ph, pw = kh // 2, kw // 2
        xin = torch.zeros(n, c, h + 2 * ph, w + 2 * pw, dtype=F32, device=canvas.device)  # zero-padded (x + pos)
        pos_emb = self._pos[:, :, :h, :w]
        xin[:, :, ph: ph + h, pw: pw + w] = canvas + pos_emb
        if st["graph"] is None:
            st["patch"].copy_(xin[:, :, :kh, :kw])
            st["pos"].fill_(0)
            try:
                self._sampler_step(st)  # warm-up outside capture
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    st["logits"] = self._sampler_step(st)
                st["graph"] = graph
            except RuntimeError as exc:
                torch.cuda.synchronize()
                st["graph"] = False  # capture unavailable here: launch the same step eagerly
                st["graph_error"] = repr(exc)
                warnings.warn("ImageGPT.sample(): CUDA-graph capture of the per-pixel step failed, launching it eagerly "
                              f"(same kernels, ~3x slower): {exc!r}", RuntimeWarning)
        for row in range(h):
            for col in range(w):
                st["patch"].copy_(xin[:, :, row: row + kh, col: col + kw])
                st["pos"].fill_(row * w + col)
                if st["graph"]:
                    st["graph"].replay()
                    logits = st["logits"]
                else:
                    logits = self._sampler_step(st)
                drawn = self._sample_fn(logits).view(n, c)
                current = canvas[:, :, row, col]
                new = torch.where(current < 0, drawn, current)
                canvas[:, :, row, col] = new
                xin[:, :, row + ph, col + pw] = new + pos_emb[0, :, row, col]
        return canvas

    # ---- data-parallel bucket protocol (parallel.OverlappedGradAverager) ----
    def set_grad_bucket_hook(self, fn):
        """fn(flat_fp32_bucket) -> handle with wait(); called once per transformer block during backward."""
        self.__dict__["_grad_bucket_hook"] = fn  # per model instance; None removes it

    def bucketed_parameters(self):
        """Parameters whose gradients are averaged by the bucket hook (the block weight matrices), or [] when the
        head geometry needs slot padding (their gradients are then gathered copies, averaged by the flat bucket)."""
        c = self._ln.weight.numel()
        if c // self._n_heads != ops.HEAD_SLOT:
            return []
        out = []
        for blk in self._transformer:
            a = blk._attn
            out += [a._q.weight, a._kv.weight, a._proj.weight, blk._out[0].weight, blk._out[2].weight]
        return out

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("ImageGPT (B200 path) needs CUDA tensors; there is no CPU fallback")
        self._input.weight.data *= self._input.mask  # same in-place side effect as the reference's CausalConv2d
        flat = [self._pos, self._input.weight, self._input.bias]
        for blk in self._transformer:
            flat.extend(blk.flat_params())
        flat.extend([self._ln.weight, self._ln.bias, self._out.weight, self._out.bias])
        return _ImageGPTStack.apply(x.float(), self._n_heads, self._ln.eps, self._packed_training_weights(),
                                    dict(grad=torch.is_grad_enabled(), hook=self.__dict__.get("_grad_bucket_hook")), *flat)


def reproduce(*args, **kwargs):
    """The recipe of this model (reference image_gpt.py `reproduce`); see `pytorch_generative_b200.recipes`."""
    from .. import recipes

    return recipes.reproduce_image_gpt(*args, **kwargs)
