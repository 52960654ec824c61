"""CUDA-graph capture of one forward + loss + backward of a libsfcvit-backed model (an addition to the reference API).

The reference wraps its model in ``torch.compile(mode="reduce-overhead")`` (main.py:284), i.e. CUDA-graphed Inductor
code. The B200 path has no tracing compiler: its kernels are launched through the C ABI, so the equivalent is to capture
the ~320 launches of a step once and replay them (no Python / ctypes / tensor-map encoding on the critical path).
The training loops of src/training/train.py build one lazily per (model, criterion, batch shape).

    step = GraphedStep(model, criterion, example_images, example_targets)
    loss = step(images, targets)        # copies into the static buffers, replays, returns the static loss tensor
    optimizer.step()                    # p.grad tensors are static; do NOT zero them to None between replays

or, with a FusedAdamW (src/training/optim.py), the WHOLE step in one graph — backward's wgrad kernels write into the
optimizer's flat gradient bucket, the data-parallel all-reduce of the bucket (one NCCL call after backward by default;
bucket ranges on a side stream during backward with FusedAdamW(overlap=True)) is captured, and grad-norm + clip + AdamW
follow in the same graph:

    step = GraphedStep(model, criterion, example_images, example_targets, optimizer=fused_adamw)
    loss = step(images, targets)        # advance() (lr / bias corrections -> device) + one replay; no optimizer.step()

Construct the optimizer BEFORE the GraphedStep: FusedAdamW moves the parameters into its flat buffer at construction, and
a graph captured earlier would keep reading the old storage — GraphedStep checks the parameter addresses on every call.

Drop every reference to losses / outputs of earlier EAGER passes before constructing it: a live autograd graph keeps
AccumulateGrad nodes bound to the default stream, which stream capture cannot synchronise with.

Dropout: seeds drawn on the host are frozen into the graph, so the graph increments a device-side epoch counter that
every mask-drawing kernel mixes into its seed (``sfc_set_dropout_epoch_ptr``): each replay draws fresh masks, and the
forward / backward of one replay agree.
"""
import torch

from sfcvit import _lib, functional as SF

_epoch = {}


def dropout_epoch(device):
    """The process-wide device counter (created on first use and registered with the library)."""
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _epoch.get(key)
    if t is None:
        if _epoch:
            raise RuntimeError("the dropout epoch counter is process-wide: one CUDA device per process")
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _lib.load().sfc_set_dropout_epoch_ptr(t.data_ptr())
        _epoch[key] = t
    return t


class GraphedStep:
    def __init__(self, model, criterion, example_images, example_targets, warmup=3, optimizer=None, autocast_dtype=None):
        if not example_images.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA tensors (no CPU fallback)")
        if optimizer is not None and not (hasattr(optimizer, "advance") and hasattr(optimizer, "launch")):
            raise TypeError("GraphedStep(optimizer=...) needs a FusedAdamW (advance() / launch()); step other optimizers eagerly")
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.autocast_dtype = autocast_dtype      # run forward + loss under torch.amp.autocast (the reference loops do, train.py:155)
        self.images = example_images.clone()
        self.targets = example_targets.clone()
        self.epoch = dropout_epoch(example_images.device)
        params = [p for p in model.parameters() if p.requires_grad]
        # warm-up on a side stream (lazy initialisations, cudaFuncSetAttribute, workspace growth) as torch recommends
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for p in params:
                    p.grad = None
                SF.release_grad_claims()
                self._forward_loss()[1].backward()
                if optimizer is not None:
                    # a real step with zero learning rate would still decay; the warm-up only has to teach the optimizer
                    # which parameters receive gradients and grow every workspace, so it launches nothing that updates
                    optimizer._learn_expected(tuple(p.grad is not None for fb in optimizer._flat for (p, _o, _k) in fb["views"]))
                    optimizer._pending = None
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for p in params:
            p.grad = None                         # backward inside the capture allocates the static .grad tensors
        SF.release_grad_claims()
        SF.clear_weight_caches()                  # cached weight conversions must be RE-RUN inside the graph
        self.graph = torch.cuda.CUDAGraph()
        from sfcvit import ops
        l0 = ops.LAUNCHES
        if optimizer is not None:
            optimizer.last_num_buckets = optimizer.last_overlapped_buckets = 0
        with torch.cuda.graph(self.graph):
            self.epoch.add_(1)
            self.logits, self.loss = self._forward_loss()
            self.loss.backward()
            if optimizer is not None:
                optimizer.launch()
        self.captured_launches = ops.LAUNCHES - l0   # libsfcvit kernels per replay
        SF.clear_weight_caches()                  # entries created during capture alias graph-private memory
        self._param_ptrs = [p.data_ptr() for p in params]
        self._params = params

    def _forward_loss(self):
        if self.autocast_dtype is None:
            logits = self.model(self.images)
            return logits, self.criterion(logits, self.targets)
        with torch.amp.autocast(device_type="cuda", dtype=self.autocast_dtype):
            logits = self.model(self.images)
            return logits, self.criterion(logits, self.targets)

    def __call__(self, images, targets=None):
        if [p.data_ptr() for p in self._params] != self._param_ptrs:
            raise RuntimeError("GraphedStep: a parameter's storage moved after capture (construct FusedAdamW / load "
                               "checkpoints BEFORE building the GraphedStep); the graph would read stale weights")
        if images is not self.images:
            self.images.copy_(images, non_blocking=True)
        if targets is not None and targets is not self.targets:
            self.targets.copy_(targets, non_blocking=True)
        if self.optimizer is not None:
            self.optimizer.advance()
        self.graph.replay()
        SF.bump_weights_epoch()                   # replays update / read parameters behind torch's version counters
        return self.loss
