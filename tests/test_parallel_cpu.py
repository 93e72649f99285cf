"""World-size-2 gloo test of the N>1 host logic (parameter broadcast + flat gradient averaging) on CPU."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_generative_b200 import parallel

    torch.manual_seed(100 + rank)  # different init per rank -> broadcast must equalise
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    parallel.broadcast_parameters(model)
    avg = parallel.FlatGradAverager(model.parameters())
    g = torch.Generator().manual_seed(parallel.shard_seed(7, rank))
    x = torch.randn(4, 6, generator=g)
    model(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    avg.average_()
    torch.save({"params": [p.detach().clone() for p in model.parameters()], "local": local,
                "avg": [p.grad.clone() for p in model.parameters()], "x": x}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_gradient_average_matches_mean_over_ranks(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt") for i in range(world)]
    assert not torch.equal(r[0]["x"], r[1]["x"])  # each rank drew its own shard of synthetic data
    for a, b in zip(r[0]["params"], r[1]["params"]):
        assert torch.equal(a, b)  # broadcast from rank 0
    for i in range(len(r[0]["avg"])):
        want = (r[0]["local"][i] + r[1]["local"][i]) / 2
        assert torch.allclose(r[0]["avg"][i], want, atol=1e-6) and torch.equal(r[0]["avg"][i], r[1]["avg"][i])


class _ArenaFn(torch.autograd.Function):
    """Stand-in for the fused ImageGPT stack: both weight gradients are views of one arena, handed to the bucket hook
    per 'block' inside backward (same protocol as models/image_gpt.py)."""

    hook = None

    @staticmethod
    def forward(ctx, x, w0, w1, b):
        ctx.save_for_backward(x, w0, w1)
        return torch.tanh(x @ w0.t()) @ w1.t() + b

    @staticmethod
    def backward(ctx, dy):
        x, w0, w1 = ctx.saved_tensors
        h = torch.tanh(x @ w0.t())
        arena = torch.zeros(w0.numel() + w1.numel())
        dw1 = arena[w0.numel():].view_as(w1)
        dw1 += dy.t() @ h
        pending = []
        if _ArenaFn.hook is not None:
            pending.append(_ArenaFn.hook(arena[w0.numel():]))
        dh = (dy @ w1) * (1 - h * h)
        dw0 = arena[: w0.numel()].view_as(w0)
        dw0 += dh.t() @ x
        if _ArenaFn.hook is not None:
            pending.append(_ArenaFn.hook(arena[: w0.numel()]))
        for hd in pending:
            hd.wait()
        return None, dw0, dw1, dy.sum(0)


class _ArenaModel(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w0 = torch.nn.Parameter(torch.randn(5, 6))
        self.w1 = torch.nn.Parameter(torch.randn(3, 5))
        self.b = torch.nn.Parameter(torch.randn(3))

    def set_grad_bucket_hook(self, fn):
        _ArenaFn.hook = fn

    def bucketed_parameters(self):
        return [self.w0, self.w1]

    def forward(self, x):
        return _ArenaFn.apply(x, self.w0, self.w1, self.b)


def _overlap_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pytorch_generative_b200 import parallel

    torch.manual_seed(5)
    model = _ArenaModel()
    x = torch.randn(4, 6, generator=torch.Generator().manual_seed(parallel.shard_seed(11, rank)))
    model(x).pow(2).sum().backward()  # no hook yet: local gradients
    local = [p.grad.clone() for p in model.parameters()]
    model.zero_grad(set_to_none=True)
    avg = parallel.OverlappedGradAverager(model)
    assert avg.n_bucketed == 2
    model(x).pow(2).sum().backward()  # buckets averaged inside backward
    avg.average_()                    # the bias through the flat bucket
    torch.save({"local": local, "avg": [p.grad.clone() for p in model.parameters()]}, os.path.join(out_dir, f"o{rank}.pt"))
    dist.destroy_process_group()


def test_block_bucket_average_inside_backward(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_overlap_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"o{i}.pt") for i in range(world)]
    for i in range(3):
        want = (r[0]["local"][i] + r[1]["local"][i]) / 2
        assert torch.allclose(r[0]["avg"][i], want, atol=1e-6), i
        assert torch.equal(r[0]["avg"][i], r[1]["avg"][i])


def test_sample_sharding_is_balanced_and_complete():
    from pytorch_generative_b200 import parallel

    for n in (0, 1, 7, 16, 17):
        for world in (1, 2, 8):
            shares = [parallel.shard_samples(n, r, world) for r in range(world)]
            assert sum(shares) == n and max(shares) - min(shares) <= 1
