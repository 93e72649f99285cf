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
