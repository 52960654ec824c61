"""Training / evaluation loops — mirror of the reference's src/training/train.py (same names, signatures, return
values: (avg_loss, avg_acc) python floats). Differences, all invisible to a single-process caller such as main.py:
the model's forward/backward run on the libsfcvit kernels, and when torch.distributed is initialised (torchrun) each
step is batch-sharded across ranks with an NCCL gradient all-reduce between backward and the clip/optimizer step
(reference insertion point: train.py:163-165). Rank 0 alone shows progress bars."""
import os
import weakref

import numpy as np
import torch
import torch.nn.functional as F
from tqdm import tqdm

from . import distributed as D
from ._nvtx import span


def mixup_data(x, y, alpha=0.2):
    """Convex mix of the batch with a shuffled copy (reference :7-14). Host RNG: numpy for lambda, torch for the perm."""
    lam = np.random.beta(alpha, alpha) if alpha > 0 else 1.0
    idx = D.agree(torch.randperm(x.size(0), device=x.device))
    return lam * x + (1 - lam) * x[idx], y, y[idx], lam


def rand_bbox(H, W, lam):
    """Random box of relative area (1 - lam) (reference :17-30)."""
    cut_rat = np.sqrt(1. - lam)
    cut_w, cut_h = int(W * cut_rat), int(H * cut_rat)
    cx, cy = np.random.randint(W), np.random.randint(H)
    return (np.clip(cx - cut_w // 2, 0, W), np.clip(cy - cut_h // 2, 0, H),
            np.clip(cx + cut_w // 2, 0, W), np.clip(cy + cut_h // 2, 0, H))


def cutmix_data(x, y, alpha=0.2):
    """Paste a box from a shuffled copy, in place (reference :33-47; the x-extent indexes dim 2 as there, :42)."""
    lam = np.random.beta(alpha, alpha) if alpha > 0 else 1.0
    batch_size, _, H, W = x.size()
    idx = D.agree(torch.randperm(batch_size, device=x.device))
    bbx1, bby1, bbx2, bby2 = rand_bbox(H, W, lam)
    x[:, :, bbx1:bbx2, bby1:bby2] = x[idx, :, bbx1:bbx2, bby1:bby2]
    lam = 1 - ((bbx2 - bbx1) * (bby2 - bby1) / (H * W))
    return x, y, y[idx], lam


def mixup_criterion(criterion, pred, y_a, y_b, lam):
    """lam * L(pred, y_a) + (1 - lam) * L(pred, y_b) (reference :50-54)."""
    return lam * criterion(pred, y_a) + (1 - lam) * criterion(pred, y_b)


# ---- lazily graphed steps ------------------------------------------------------------------------------------------
# The reference wraps its model in torch.compile(mode="reduce-overhead") (main.py:284), i.e. CUDA-graphed code. The
# libsfcvit kernels are launched through a C ABI that no tracing compiler sees, so the loops below get the same effect
# themselves: the first full-size batch of an epoch function captures forward + loss + backward ONCE per (model,
# criterion, batch shape) and every later step replays it (no Python / ctypes / tensor-map encoding per kernel). The
# caller's optimizer, scheduler and gradient clipping stay what main.py passes in. SFC_TRAIN_GRAPH=0 turns it off; a
# ragged last batch, CPU tensors or a failed capture fall back to launching the same kernels eagerly.
_GRAPHS = weakref.WeakKeyDictionary()      # model -> {(criterion id, shapes, dtypes, autocast): GraphedStep | False}; dies with the model


def _graphed_step(model, criterion, images, targets, autocast_dtype):
    if os.environ.get("SFC_TRAIN_GRAPH", "1") == "0" or not images.is_cuda:
        return None
    per_model = _GRAPHS.setdefault(model, {})
    key = (id(criterion), tuple(images.shape), images.dtype, tuple(targets.shape), targets.dtype, autocast_dtype)
    ent = per_model.get(key)
    if ent is not None and ent and ent.criterion is not criterion:     # the id was recycled by another criterion object
        ent = None
    if ent is None:
        if sum(1 for v in per_model.values() if v) >= 2:              # at most two captured shapes per model (activations are pinned)
            return None
        from .graphs import GraphedStep
        try:
            ent = GraphedStep(model, criterion, images, targets, autocast_dtype=autocast_dtype)
        except Exception as e:                                     # e.g. a criterion with a host sync: stay eager, say so once
            import warnings
            warnings.warn(f"sfcvit: step not capturable in a CUDA graph ({type(e).__name__}: {e}); launching eagerly")
            ent = False
        per_model[key] = ent
    return ent or None


def _has_graphs(model):
    return any(bool(v) for v in _GRAPHS.get(model, {}).values())


def _step(model, criterion, optimizer, images, targets, full_batch, autocast_dtype):
    """forward + loss + backward of one batch: a replay of the captured step when the batch has the captured shape, the
    same kernels launched eagerly otherwise. Returns (outputs, loss); gradients are in .grad either way."""
    step = _graphed_step(model, criterion, images, targets, autocast_dtype) if images.size(0) == full_batch else None
    if step is not None:
        loss = step(images, targets)                               # gradients are (re)written into the static .grad tensors
        return step.logits, loss
    optimizer.zero_grad(set_to_none=not _has_graphs(model))       # keep captured .grad tensors alive: zero them in place
    if autocast_dtype is None:
        outputs = model(images)
        loss = criterion(outputs, targets)
    else:
        with torch.amp.autocast(device_type="cuda", dtype=autocast_dtype):
            outputs = model(images)
            loss = criterion(outputs, targets)
    loss.backward()
    return outputs, loss


_NUM_CLASSES = weakref.WeakKeyDictionary()


def _num_classes(model, images):
    n = _NUM_CLASSES.get(model)
    if n is None:
        was = model.training
        model.eval()
        with torch.no_grad(), torch.amp.autocast(device_type="cuda", dtype=torch.bfloat16):
            n = int(model(images[:1]).size(1))
        model.train(was)
        _NUM_CLASSES[model] = n
    return n


def _clip_grad_norm(params, max_norm):
    """torch.nn.utils.clip_grad_norm_ (reference :165). The reference asks for foreach=False (one norm kernel per
    parameter, ~3 launches x 150 tensors); the multi-tensor form computes the same norm in a handful of launches."""
    params = [p for p in params if p.grad is not None]
    try:
        return torch.nn.utils.clip_grad_norm_(params, max_norm, foreach=True)
    except RuntimeError:
        return torch.nn.utils.clip_grad_norm_(params, max_norm, foreach=False)


def _bar(loader, desc):
    rank, _ = D.world()
    return tqdm(loader, desc=desc, leave=False, disable=rank != 0)


def _reduce_epoch(device, *sums):
    return D.allreduce_scalars(list(sums), device)


def train(model, train_loader, criterion, optimizer, device):
    """One epoch, plain cross-entropy (reference :57-77)."""
    model.train()
    rank, ws = D.world()
    D.sync_module(model)
    total_loss, correct, seen = 0.0, 0.0, 0
    full_batch = None
    for images, labels in _bar(train_loader, "Training"):
        images, labels = D.shard(images.to(device), rank, ws), D.shard(labels.to(device), rank, ws)
        if full_batch is None:
            full_batch = images.size(0)
        outputs, loss = _step(model, criterion, optimizer, images, labels, full_batch, None)
        D.allreduce_gradients(model.parameters(), ws)
        optimizer.step()
        total_loss += loss.item() * images.size(0)
        correct += (outputs.argmax(dim=1) == labels).sum().item()
        seen += images.size(0)
    total_loss, correct, seen = _reduce_epoch(device, total_loss, correct, seen)
    n = len(train_loader.dataset) if ws == 1 else max(seen, 1)
    return total_loss / n, correct / n


def evaluate(model, test_loader, criterion, device):
    """Evaluation under bf16 autocast (reference :80-99; device_type is hard-coded to "cuda" there too)."""
    model.eval()
    rank, ws = D.world()
    D.sync_module(model)
    total_loss, correct, seen = 0.0, 0.0, 0
    with torch.no_grad():
        for images, labels in _bar(test_loader, "Evaluating"):
            images, labels = D.shard_all(images.to(device), rank, ws), D.shard_all(labels.to(device), rank, ws)
            if images.size(0) == 0:
                continue
            with torch.amp.autocast(device_type="cuda", dtype=torch.bfloat16):
                outputs = model(images)
                loss = criterion(outputs, labels)
            total_loss += loss.item() * images.size(0)
            correct += (outputs.argmax(dim=1) == labels).sum().item()
            seen += images.size(0)
    total_loss, correct, seen = _reduce_epoch(device, total_loss, correct, seen)
    n = len(test_loader.dataset) if ws == 1 else max(seen, 1)
    return total_loss / n, correct / n


def train_with_scheduler(model, train_loader, criterion, optimizer, scheduler, device):
    """One epoch with a per-step scheduler whose step() returns the rate (reference :102-130)."""
    model.train()
    rank, ws = D.world()
    D.sync_module(model)
    total_loss, correct, seen = 0.0, 0.0, 0
    full_batch = None
    bar = _bar(train_loader, "Training")
    for images, labels in bar:
        images, labels = D.shard(images.to(device), rank, ws), D.shard(labels.to(device), rank, ws)
        if full_batch is None:
            full_batch = images.size(0)
        outputs, loss = _step(model, criterion, optimizer, images, labels, full_batch, torch.bfloat16)
        D.allreduce_gradients(model.parameters(), ws)
        optimizer.step()
        current_lr = scheduler.step()
        total_loss += loss.item() * images.size(0)
        correct += (outputs.argmax(dim=1) == labels).sum().item()
        seen += images.size(0)
        bar.set_postfix(loss=f"{loss.item():.4f}", lr=f"{current_lr:.6f}")
    total_loss, correct, seen = _reduce_epoch(device, total_loss, correct, seen)
    n = len(train_loader.dataset) if ws == 1 else max(seen, 1)
    return total_loss / n, correct / n


def train_with_mixup_or_cutmix(model, train_loader, criterion, optimizer, scheduler, device, mixup_alpha=0.2,
                               cutmix_alpha=1.0, mix_prob=0.5):
    """One epoch of soft-target training with mixup or cutmix per step, grad-norm clip at 1.0 (reference :133-178).
    Under data parallelism the augmentation is applied to the FULL batch with the rank-shared host RNG and the batch
    is sharded afterwards, so the union over ranks equals the single-process batch."""
    model.train()
    rank, ws = D.world()
    D.sync_module(model)
    total_loss, total_correct, total_samples = 0.0, 0.0, 0
    full_batch = None
    bar = _bar(train_loader, "Training")
    for images, labels in bar:
        with span("batch prep"):
            images, labels = images.to(device), labels.to(device)
            if full_batch is None:
                full_batch = D.shard(images, rank, ws).size(0)      # the loader's batch size (per rank): the shape that is captured
            if np.random.rand() < mix_prob:
                images, y_a, y_b, lam = mixup_data(images, labels, alpha=mixup_alpha)
            else:
                images, y_a, y_b, lam = cutmix_data(images, labels, alpha=cutmix_alpha)
            images, y_a, y_b = D.shard(images, rank, ws), D.shard(y_a, rank, ws), D.shard(y_b, rank, ws)
            num_classes = _num_classes(model, images)
            soft_targets = lam * F.one_hot(y_a, num_classes).float() + (1 - lam) * F.one_hot(y_b, num_classes).float()
        with span("forward+backward"):
            outputs, loss = _step(model, criterion, optimizer, images, soft_targets, full_batch, torch.bfloat16)
        with span("allreduce"):
            D.allreduce_gradients(model.parameters(), ws)
        with span("clip+optimizer"):
            _clip_grad_norm(model.parameters(), 1.0)
            optimizer.step()
            scheduler.step()
        preds = outputs.argmax(dim=1)
        total_correct += (lam * (preds == y_a).float() + (1 - lam) * (preds == y_b).float()).sum().item()
        total_loss += loss.item() * images.size(0)
        total_samples += images.size(0)
    bar.close()
    total_loss, total_correct, total_samples = _reduce_epoch(device, total_loss, total_correct, total_samples)
    return total_loss / total_samples, total_correct / total_samples
