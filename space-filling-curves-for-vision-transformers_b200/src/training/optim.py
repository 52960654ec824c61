"""FusedAdamW — drop-in for `optim.AdamW(model.parameters(), ...)` (reference main.py:288-289) that also absorbs
`torch.nn.utils.clip_grad_norm_(params, max_norm)` (reference train.py:165) and, under torchrun, the data-parallel
gradient all-reduce: parameters live in ONE flat HBM buffer per dtype (the nn.Parameters become views, names/shapes/
state_dict unchanged), gradients are packed into a flat bucket, all-reduced in place over NCCL, and a single pass of
kernel K6 (csrc/optim.cu) computes ||g||, derives the clip coefficient on device and applies the AdamW update.
An addition to the reference API, used by bench.py; main.py keeps working with torch's own optimizer."""
import torch
import torch.distributed as dist

from sfcvit import ops
from . import distributed as D


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0,
                 state_dtype=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._state_dtype = state_dtype
        self._flat = None          # built lazily at the first step (needs to know which params receive gradients)
        self._stats = None
        self.last_num_buckets = 0

    def _build(self):
        self._flat = []
        for gi, group in enumerate(self.param_groups):
            by_dtype = {}
            for p in group["params"]:
                if p.grad is None:
                    continue                     # torch's AdamW skips grad-less params too (e.g. MixerBlock.token_mix*)
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamW: parameters must live on a CUDA device (no CPU fallback)")
                by_dtype.setdefault(p.dtype, []).append(p)
            for dtype, plist in by_dtype.items():
                n = sum(p.numel() for p in plist)
                dev = plist[0].device
                flat_p = torch.empty(n, dtype=dtype, device=dev)
                sdt = self._state_dtype or dtype
                flat_m = torch.zeros(n, dtype=sdt, device=dev)
                flat_v = torch.zeros(n, dtype=sdt, device=dev)
                flat_g = torch.empty(n, dtype=dtype, device=dev)
                off = 0
                views = []
                for p in plist:
                    k = p.numel()
                    flat_p[off:off + k].copy_(p.data.reshape(-1))
                    p.data = flat_p[off:off + k].view_as(p)                # the Parameter now aliases the flat buffer
                    st = self.state[p]
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = flat_m[off:off + k].view_as(p)
                    st["exp_avg_sq"] = flat_v[off:off + k].view_as(p)
                    views.append((p, off, k))
                    off += k
                self._flat.append(dict(group=gi, p=flat_p, m=flat_m, v=flat_v, g=flat_g, views=views, step=0))
        if self._flat:
            self._stats = torch.zeros(1, dtype=torch.float32, device=self._flat[0]["p"].device)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._flat is None:
            self._build()
        if not self._flat:
            return loss
        rank, world = D.world()
        # 1) pack gradients into the flat buckets
        for fb in self._flat:
            g = fb["g"]
            grads = [p.grad if p.grad is not None else None for (p, _o, _k) in fb["views"]]
            if all(x is not None for x in grads):
                torch.cat([x.reshape(-1) for x in grads], out=g)
            else:
                g.zero_()
                for (p, off, k), x in zip(fb["views"], grads):
                    if x is not None:
                        g[off:off + k].copy_(x.reshape(-1))
        # 2) data-parallel sum over ranks, in place on the flat bucket (NCCL over NVLink)
        self.last_num_buckets = 0
        if world > 1:
            for fb in self._flat:
                g = fb["g"]
                chunk = max(1, D.BUCKET_BYTES // g.element_size())
                for s in range(0, g.numel(), chunk):
                    dist.all_reduce(g[s:s + chunk], op=dist.ReduceOp.SUM)
                    self.last_num_buckets += 1
        # 3) global gradient norm (sum of squares over all buckets), stays on device
        max_norm = max(float(gr["max_grad_norm"]) for gr in self.param_groups)
        if max_norm > 0:
            self._stats.zero_()
            for fb in self._flat:
                ops.grad_sumsq(fb["g"], self._stats)
        # 4) fused clip + AdamW
        for fb in self._flat:
            gr = self.param_groups[fb["group"]]
            fb["step"] += 1
            b1, b2 = gr["betas"]
            ops.adamw_step(fb["p"], fb["g"], fb["m"], fb["v"], lr=gr["lr"], beta1=b1, beta2=b2, eps=gr["eps"],
                           weight_decay=gr["weight_decay"], step=fb["step"], grad_scale=1.0 / world,
                           max_norm=float(gr["max_grad_norm"]), stats=self._stats if max_norm > 0 else None)
            for (p, _o, _k) in fb["views"]:
                self.state[p]["step"] += 1
        return loss

    def grad_norm(self):
        """Global gradient norm of the last step (device -> host sync; diagnostics only)."""
        _, world = D.world()
        return float(self._stats.sqrt().item()) / world if self._stats is not None else 0.0
