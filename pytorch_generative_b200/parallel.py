"""Data parallelism for the training step: the only multi-GPU mechanism on the path (reference trainer.py:78-82 wraps
the model in DistributedDataParallel).

One process per GPU, full replica each, rank r draws its own batch; once per step the gradients are averaged with ONE
flat NCCL all-reduce over NVLink/NVSwitch.  The fused model stacks produce all parameter gradients inside a single
autograd node, so there is nothing for DDP's per-bucket hooks to overlap with: a flat bucket (one collective launch,
no per-tensor launches) is both simpler and cheaper.  The helpers are backend-agnostic (`gloo` on CPU in the tests).
"""

import torch
import torch.distributed as dist


def broadcast_parameters(module, src=0):
    """Makes every rank start from rank `src`'s parameters and buffers (what DDP does at construction)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


class FlatGradAverager:
    """Averages the gradients of `params` across ranks with one all-reduce over a persistent flat fp32 bucket."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(numel, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            off += p.numel()

    @torch.no_grad()
    def average_(self):
        """In place: p.grad <- mean over ranks of p.grad (parameters without a gradient contribute zeros).

        Packing and unpacking are multi-tensor copies (a few launches for the whole bucket, not two per parameter);
        NCCL averages inside the collective, other backends sum and the bucket is scaled afterwards."""
        if self.world == 1:
            return
        have = [(p, v) for p, v in zip(self.params, self.views) if p.grad is not None]
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
        if have:
            torch._foreach_copy_([v for _, v in have], [p.grad for p, _ in have])
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(self.world)
        if have:
            torch._foreach_copy_([p.grad for p, _ in have], [v for _, v in have])
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()


def shard_seed(base_seed, rank):
    """Synthetic-data seed of a rank: every rank draws its own batch (weak scaling, like the reference's loaders)."""
    return int(base_seed) + int(rank)
