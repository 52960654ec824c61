"""Data-parallel plumbing for the training loops (the reference is single-GPU; this is the one added strategy:
batch sharding + gradient all-reduce).

One process per GPU (torchrun). The loader is identical and identically seeded on every rank (main.py:151-154,
227-228 has no DistributedSampler), so each rank applies mixup/cutmix to the FULL batch with the shared host RNG and
then keeps `batch[rank::world]`. Gradients of the local mean loss are averaged over ranks with one all-reduce per
flat bucket (NCCL over NVLink on GPUs, gloo in the CPU tests) — equal shard sizes make this the full-batch gradient.
"""
import os
import weakref

import torch
import torch.distributed as dist

BUCKET_BYTES = 256 << 20      # one NCCL call for the 213 MB of ViT-B bf16 gradients (launch latency, not link count, is the cost)


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (no-op when WORLD_SIZE is unset or 1)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or (dist.is_available() and dist.is_initialized()):
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)
    return world()


def shard(t, rank, world_size):
    """This rank's slice of a full TRAINING batch: every world_size-th sample starting at `rank`, equal shard sizes
    (mean-of-means == global mean; up to world_size - 1 trailing samples of a ragged last batch are dropped, as a
    DistributedSampler with drop_last would)."""
    if world_size == 1:
        return t
    n = t.shape[0] - t.shape[0] % world_size
    return t[rank:n:world_size]


def shard_all(t, rank, world_size):
    """This rank's slice of an EVALUATION batch: every sample goes to exactly one rank (the remainder to the low
    ranks); callers weight their sums by the actual sample count, so metrics cover the full test set."""
    if world_size == 1:
        return t
    return t[rank::world_size]


def agree(t, src=0):
    """Rank `src`'s value of a small device tensor on every rank (the full-batch mixup / cutmix permutation is drawn from
    the per-rank CUDA generator: equal seeds make it equal, this makes it certain)."""
    _, ws = world()
    if ws > 1:
        dist.broadcast(t, src=src)
    return t


_synced = weakref.WeakSet()


def sync_module(module, src=0, check=True):
    """Broadcast parameters and buffers from rank `src` (once per module) so the replicas start bit-identical even if
    their RNG streams diverged during construction; with check=True, also assert afterwards that a checksum agrees."""
    rank, ws = world()
    if ws == 1 or module in _synced:
        return
    with torch.no_grad():
        tensors = [p.data for p in module.parameters()] + [b for b in module.buffers()]
        for t in tensors:
            if t.numel():
                dist.broadcast(t, src=src)
        if check and tensors:
            dev = tensors[0].device
            cs = torch.stack([t.double().sum().to(dev) for t in tensors if t.numel() and t.is_floating_point()]).sum().reshape(1)
            lo, hi = cs.clone(), cs.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if float(lo) != float(hi):
                raise RuntimeError("data-parallel replicas differ after the initial broadcast")
    _synced.add(module)


def build_buckets(tensors, bucket_bytes=BUCKET_BYTES):
    """Group tensors (in the given order) into lists of at most bucket_bytes, one dtype per bucket."""
    buckets, cur, cur_bytes, cur_dtype = [], [], 0, None
    for t in tensors:
        nbytes = t.numel() * t.element_size()
        if cur and (cur_dtype != t.dtype or cur_bytes + nbytes > bucket_bytes):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(t)
        cur_bytes += nbytes
        cur_dtype = t.dtype
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params, world_size=None, average=True, bucket_bytes=BUCKET_BYTES):
    """Sum (and average) `.grad` of the parameters that received one, over all ranks, in flat buckets walked in
    reverse registration order (the order backward produces them). Parameters without a gradient (e.g. the unused
    MixerBlock.token_mix*, vit.py:269-271) are skipped consistently because every rank runs the same graph."""
    if world_size is None:
        _, world_size = world()
    if world_size == 1:
        return 0
    grads = [p.grad for p in reversed(list(params)) if p.grad is not None]
    n_buckets = 0
    for bucket in build_buckets(grads, bucket_bytes):
        flat = torch.cat([g.reshape(-1) for g in bucket]) if len(bucket) > 1 else bucket[0].reshape(-1)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world_size)
        if len(bucket) > 1:
            off = 0
            for g in bucket:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        n_buckets += 1
    return n_buckets


def allreduce_scalars(values, device):
    """Sum a list of python floats over ranks (epoch metrics); one tiny collective."""
    _, ws = world()
    if ws == 1:
        return list(values)
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()
