"""Two-GPU NCCL test of the data-parallel training path (skipped on a single-GPU box): one process per GPU, the global
batch sharded as src.training.distributed.shard does, FusedAdamW all-reducing the flat gradient bucket.
Invariants: (a) the averaged gradient equals the single-process full-batch gradient (bf16 tolerance), (b) after the
optimizer step every rank holds bit-identical parameters, (c) the loss of the sharded run averages to the full-batch loss."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build(device):
    import cases
    from oracle import model as om
    from src.models.vit import VisionTransformer
    vk, tcase, mkw, _ = cases.MODEL_CASES["vit_conv_hilbert_tiny"]
    kind, kw, shape = cases.TOKENIZER_CASES[tcase]
    torch.manual_seed(cases.INIT_SEED)
    m = om.zero_dropout(VisionTransformer(patch_embed=cases.build_src_tokenizer(kind, kw), **mkw)).to(device).train()
    return m, shape, mkw["num_classes"]


def _step(m, x, t, world_expected):
    from oracle.model import soft_target_cross_entropy
    from src.training.optim import FusedAdamW
    opt = FusedAdamW(m.parameters(), lr=0.0, weight_decay=0.0, max_grad_norm=1.0, comm_buckets=3, overlap=True)
    # step 1 (lr = 0) teaches the optimizer which parameters receive gradients: its ranges are reduced inside step();
    # step 2 is the overlapped path — per-parameter hooks all-reduce each range on the side stream during backward
    for it in range(2):
        opt.zero_grad()
        loss = soft_target_cross_entropy(m(x).float(), t)
        loss.backward()
        for g in opt.param_groups:
            g["lr"] = 0.0 if it == 0 else 1e-3
        opt.step()
    assert world_expected == 1 or opt.last_overlapped_buckets >= 2, (opt.last_overlapped_buckets, opt.last_num_buckets)
    flat = torch.cat([fb["g"].float() for fb in opt._flat]) / world_expected
    return float(loss.detach()), flat.cpu(), torch.cat([p.detach().float().reshape(-1) for p in m.parameters()]).cpu()


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "space-filling-curves-for-vision-transformers_b200"), os.path.join(root, "tests", "golden")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    import cases
    from src.training import distributed as D
    assert D.init_from_env("nccl") == (rank, world)
    dev = torch.device("cuda", rank)
    m, shape, classes = _build(dev)
    x = cases.make_input((8,) + tuple(shape[1:]))
    t = cases.make_soft_targets(8, classes)
    loss, grad, params = _step(m, D.shard(x, rank, world).to(dev), D.shard(t, rank, world).to(dev), world)
    torch.save(dict(loss=loss, grad=grad, params=params), f"{out}.{rank}")
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_gpu_step_matches_full_batch(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import cases
    out = str(tmp_path / "dp")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["params"], r1["params"])                       # (b) replicas stay identical
    assert torch.equal(r0["grad"], r1["grad"])
    dev = torch.device("cuda", 0)
    m, shape, classes = _build(dev)
    x = cases.make_input((8,) + tuple(shape[1:]))
    t = cases.make_soft_targets(8, classes)
    loss, grad, _ = _step(m, x.to(dev), t.to(dev), 1)
    assert float((r0["grad"] - grad).norm() / grad.norm()) < 3e-2        # (a) bf16 gradients, each rounded once
    assert abs(0.5 * (r0["loss"] + r1["loss"]) - loss) < 2e-2            # (c)
