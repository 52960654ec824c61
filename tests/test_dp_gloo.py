"""world_size-2 gloo test of the data-parallel plumbing (CPU): sharding the batch after full-batch augmentation and
averaging gradients over ranks reproduces the single-process full-batch gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from src.training import distributed as D
    assert D.init_from_env("gloo") == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    unused = torch.nn.Parameter(torch.zeros(4))           # never receives a gradient (like MixerBlock.token_mix*)
    params = list(model.parameters()) + [unused]
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    xs, ys = D.shard(x, rank, world), D.shard(y, rank, world)
    ((model(xs) - ys) ** 2).mean().backward()
    n = D.allreduce_gradients(params, world, bucket_bytes=64)
    assert n >= 2 and unused.grad is None
    sums = D.allreduce_scalars([1.0, float(rank)], torch.device("cpu"))
    assert sums == [float(world), float(sum(range(world)))]
    if rank == 0:
        torch.save([p.grad.clone() for p in model.parameters()], out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradients_equal_full_batch(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    ((model(x) - y) ** 2).mean().backward()
    for g, p in zip(got, model.parameters()):
        assert torch.allclose(g, p.grad, atol=1e-6), (g - p.grad).abs().max()
