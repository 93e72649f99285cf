"""Data-parallel gradient averaging on real GPUs over NCCL (reference trainer.py:75-82 wraps the model in DDP): the
block-bucketed all-reduce that runs inside the fused ImageGPT backward must leave every rank with the mean of the
per-rank gradients.  Needs two GPUs (skipped on a single-GPU box; the gloo variant runs on CPU in test_parallel_cpu.py)."""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(in_channels=3, out_channels=3, in_size=16, n_transformer_blocks=2, n_attention_heads=8, n_embedding_channels=512)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randint(0, 256, (2, 3, 16, 16), generator=g).float() / 255


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from pytorch_generative_b200 import losses, models, parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(rank)  # different initial weights per rank: broadcast_parameters must make them rank 0's
    m = models.ImageGPT(**CFG).to(dev)
    parallel.broadcast_parameters(m)
    avg = parallel.OverlappedGradAverager(m)
    assert avg.n_bucketed == 2 * 5
    x = _batch(rank).to(dev)
    losses.bce_with_logits_sum_mean(m(x), x).backward()
    avg.average_()
    torch.save({k: p.grad.detach().cpu() for k, p in m.named_parameters()}, os.path.join(out_dir, f"grads_{rank}.pt"))
    if rank == 0:
        torch.save({k: v.detach().cpu() for k, v in m.state_dict().items()}, os.path.join(out_dir, "state.pt"))
    avg.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_overlapped_grad_averaging_over_nccl(tmp_path):
    import torch.multiprocessing as mp

    from pytorch_generative_b200 import losses, models

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0, g1 = torch.load(tmp_path / "grads_0.pt"), torch.load(tmp_path / "grads_1.pt")
    for k in g0:
        assert torch.equal(g0[k], g1[k]), f"{k}: ranks disagree after averaging"
    # single-process reference: mean over the two batches of the gradients of rank 0's weights
    m = models.ImageGPT(**CFG)
    m.load_state_dict(torch.load(tmp_path / "state.pt"))
    m = m.to("cuda:0")
    ref = None
    for r in range(world):
        m.zero_grad()
        x = _batch(r).to("cuda:0")
        losses.bce_with_logits_sum_mean(m(x), x).backward()
        cur = {k: p.grad.detach().cpu() / world for k, p in m.named_parameters()}
        ref = cur if ref is None else {k: ref[k] + cur[k] for k in cur}
    for k in ref:
        scale = max(1.0, float(ref[k].abs().max()))
        assert float((g0[k] - ref[k]).abs().max()) <= 2e-3 * scale, k
