"""FusedAdamW — drop-in for `optim.AdamW(model.parameters(), ...)` (reference main.py:288-289) that also absorbs
`torch.nn.utils.clip_grad_norm_(params, max_norm)` (reference train.py:165) and, under torchrun, the data-parallel
gradient all-reduce.

Layout: parameters, gradients and both Adam moments live in ONE flat HBM buffer per (param group, dtype) each, built
EAGERLY at construction; the nn.Parameters become views (names / shapes / state_dict unchanged) and every parameter's
gradient slice is registered with sfcvit.functional so that the wgrad GEMMs and reductions write gradients straight into
the bucket (no packing copy). Gradients produced any other way are copied into their slice at step().

Data parallel: by default ONE all-reduce of the whole bucket after backward (no packing copy: the gradients are already
there). With overlap=True the bucket is cut into `comm_buckets` contiguous ranges; a post-accumulate-grad hook per
parameter counts a range down and, when its last gradient has landed, all-reduces the range on a side stream while
backward continues on the main stream (reverse registration order == the order backward produces gradients). On B200
with this library's persistent one-CTA-per-SM kernels that is measured SLOWER (2 GPUs: 32.86 vs 32.49 ms per step,
profiles/r2_bench_2gpu_*.json): NCCL's CTAs find no free SM until a kernel ends and then hold up the next kernel's wave.
step() joins the side stream, reduces whatever is left, then runs kernel K6 (csrc/optim.cu): deterministic sum of squares -> clip coefficient on device ->
AdamW with the 1/world averaging folded in. Learning rate and bias corrections reach the kernel through a small device
block written in stream order, so the whole step() can sit inside a CUDA graph (src/training/graphs.py) and still follow
a host-side scheduler.

An addition to the reference API, used by bench.py; main.py keeps working with torch's own optimizer."""
import math

import torch
import torch.distributed as dist

from sfcvit import functional as SF
from sfcvit import ops
from . import distributed as D

_ALIGN = 64            # elements: every parameter starts on a 128-byte (bf16) / 256-byte (fp32) boundary of the bucket


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0,
                 state_dtype=None, comm_buckets=1, overlap=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._state_dtype = state_dtype
        self._comm_buckets = max(1, int(comm_buckets))
        self._overlap = bool(overlap)
        self._step_count = 0
        self._step_t = torch.tensor(0.0)          # shared by every parameter's state["step"]
        self._comm_stream = None
        self._hooks = []
        self.last_num_buckets = 0
        self.last_overlapped_buckets = 0
        self.skip_comm = False                    # measurement only (bench.py: exposed all-reduce time): no collective at all
        self._build()

    # ------------------------------------------------------------------ construction
    def _build(self):
        self._flat = []
        dev = None
        for gi, group in enumerate(self.param_groups):
            by_dtype = {}
            for p in group["params"]:
                if not p.requires_grad:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamW: parameters must live on a CUDA device (no CPU fallback)")
                by_dtype.setdefault(p.dtype, []).append(p)
            for dtype, plist in by_dtype.items():
                dev = plist[0].device
                offs, n = [], 0
                for p in plist:
                    offs.append(n)
                    n += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                sdt = self._state_dtype or dtype
                flat_p = torch.zeros(n, dtype=dtype, device=dev)       # padding stays 0 under AdamW (g = m = v = p = 0)
                flat_g = torch.zeros(n, dtype=dtype, device=dev)
                flat_m = torch.zeros(n, dtype=sdt, device=dev)
                flat_v = torch.zeros(n, dtype=sdt, device=dev)
                views = []
                with torch.no_grad():
                    for p, off in zip(plist, offs):
                        k = p.numel()
                        flat_p[off:off + k].copy_(p.data.reshape(-1))
                        p.data = flat_p[off:off + k].view(p.shape)     # the Parameter now aliases the flat buffer
                        st = self.state[p]
                        st["step"] = self._step_t
                        st["exp_avg"] = flat_m[off:off + k].view(p.shape)
                        st["exp_avg_sq"] = flat_v[off:off + k].view(p.shape)
                        SF.register_grad_buffer(p, flat_g, off)
                        views.append((p, off, k))
                fb = dict(group=gi, p=flat_p, m=flat_m, v=flat_v, g=flat_g, views=views, n=n,
                          hyper=torch.zeros(4, dtype=torch.float32, device=dev), runs=None, runs_key=None)
                fb["ranges"] = self._cut_ranges(views, n)
                self._flat.append(fb)
        SF.bump_weights_epoch()
        self._stats = torch.zeros(1, dtype=torch.float32, device=dev) if self._flat else None
        # per-parameter bookkeeping for the overlapped all-reduce
        self._where = {}                                   # id(param) -> (flat index, range index, offset, numel)
        for fi, fb in enumerate(self._flat):
            for (p, off, k) in fb["views"]:
                ri = next(i for i, (s, e) in enumerate(fb["ranges"]) if s <= off < e)
                self._where[id(p)] = (fi, ri, off, k)
        self._expected = None                              # learned at the first step: which parameters receive gradients
        self._pending = None
        self._launched = set()
        for fb in self._flat:
            for (p, _o, _k) in fb["views"]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _cut_ranges(self, views, n):
        """Contiguous [start, end) ranges of roughly equal size whose cuts fall on parameter boundaries."""
        target = max(1, n // self._comm_buckets)
        cuts, last = [0], 0
        for (_p, off, _k) in views:
            if off - last >= target and len(cuts) < self._comm_buckets:
                cuts.append(off)
                last = off
        cuts.append(n)
        return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]

    def close(self):
        """Detach from the parameters: remove the hooks and the gradient-slice registrations (the parameters keep
        aliasing the flat buffer). Call before dropping an optimizer whose model lives on."""
        for h in self._hooks:
            h.remove()
        self._hooks = []
        SF.unregister_grad_buffers([p for fb in self._flat for (p, _o, _k) in fb["views"]])

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ gradient intake / overlapped all-reduce
    def _grad_into_bucket(self, p):
        fi, _ri, off, k = self._where[id(p)]
        g = self._flat[fi]["g"]
        if p.grad.data_ptr() != g.data_ptr() + off * g.element_size():
            g[off:off + k].copy_(p.grad.reshape(-1))       # produced outside the registered slice (generic autograd path)

    def _on_grad(self, p):
        """post-accumulate-grad hook: count the parameter's range down; launch its all-reduce when complete."""
        _, world = D.world()
        if world == 1 or not self._overlap or self._expected is None or self.skip_comm:
            return
        key = self._where.get(id(p))
        if key is None or id(p) not in self._expected:
            return
        if self._pending is None:
            self._begin_backward()
        fi, ri, _off, _k = key
        self._grad_into_bucket(p)
        self._pending[(fi, ri)] -= 1
        if self._pending[(fi, ri)] == 0:
            self._launch_range(fi, ri, overlapped=True)

    def _begin_backward(self):
        self._pending = dict(self._expected_counts)
        self._launched = set()
        self.last_num_buckets = 0                 # counters describe the step that is starting
        self.last_overlapped_buckets = 0

    def _launch_range(self, fi, ri, overlapped):
        fb = self._flat[fi]
        s, e = fb["ranges"][ri]
        if overlapped and fb["g"].is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=fb["g"].device)
            cur = torch.cuda.current_stream(fb["g"].device)
            self._comm_stream.wait_stream(cur)             # the range's gradients are complete in stream order
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(fb["g"][s:e], op=dist.ReduceOp.SUM)
            self.last_overlapped_buckets += 1
        else:
            dist.all_reduce(fb["g"][s:e], op=dist.ReduceOp.SUM)
        self._launched.add((fi, ri))
        self.last_num_buckets += 1

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        SF.release_grad_claims()

    # ------------------------------------------------------------------ the step
    def advance(self):
        """Host part of a step: count it and hand lr / bias corrections to the device block the kernel reads. Stream
        ordered (values travel as kernel arguments) — call it before replaying a graph that holds launch()."""
        self._step_count += 1
        self._step_t.fill_(float(self._step_count))
        t = self._step_count
        for fb in self._flat:
            gr = self.param_groups[fb["group"]]
            b1, b2 = gr["betas"]
            ops.store_f32x4(fb["hyper"], gr["lr"], 1.0 - b1 ** t, math.sqrt(1.0 - b2 ** t), float(t))

    @torch.no_grad()
    def launch(self):
        """Device part of a step (capturable in a CUDA graph): gradient intake, outstanding all-reduces, grad norm,
        clip + AdamW."""
        if not self._flat:
            return
        _, world = D.world()
        active_key = []
        for fb in self._flat:
            for (p, off, k) in fb["views"]:
                has = p.grad is not None
                active_key.append(has)
                if has and self._pending is None:
                    self._grad_into_bucket(p)              # ranges not handled by the hooks
        active_key = tuple(active_key)
        if world > 1 and not self.skip_comm:
            if self._pending is not None:                  # hooks ran during this backward: reduce what they left
                for fi, fb in enumerate(self._flat):
                    for ri in range(len(fb["ranges"])):
                        if (fi, ri) not in self._launched:
                            for (p, off, k) in fb["views"]:
                                if p.grad is not None and self._where[id(p)][1] == ri and id(p) not in self._expected:
                                    self._grad_into_bucket(p)
                            self._launch_range(fi, ri, overlapped=False)
                if self._comm_stream is not None:
                    torch.cuda.current_stream().wait_stream(self._comm_stream)
            else:
                for fi, fb in enumerate(self._flat):
                    for ri in range(len(fb["ranges"])):
                        self._launch_range(fi, ri, overlapped=False)
            self._pending = None
        if self._expected is None or self._expected_key != active_key:
            self._learn_expected(active_key)
        max_norm = max(float(gr["max_grad_norm"]) for gr in self.param_groups)
        if max_norm > 0:
            self._stats.zero_()
            for fb in self._flat:
                ops.grad_sumsq(fb["g"], self._stats)
        idx = 0
        for fb in self._flat:
            gr = self.param_groups[fb["group"]]
            key = active_key[idx:idx + len(fb["views"])]
            idx += len(fb["views"])
            if fb["runs_key"] != key:
                fb["runs"], fb["runs_key"] = self._runs(fb, key), key
            b1, b2 = gr["betas"]
            for (s, e) in fb["runs"]:
                ops.adamw_step(fb["p"][s:e], fb["g"][s:e], fb["m"][s:e], fb["v"][s:e], lr=gr["lr"], beta1=b1, beta2=b2,
                               eps=gr["eps"], weight_decay=gr["weight_decay"], step=0, grad_scale=1.0 / world,
                               max_norm=float(gr["max_grad_norm"]), stats=self._stats if max_norm > 0 else None,
                               hyper=fb["hyper"])
        SF.bump_weights_epoch()

    def _runs(self, fb, key):
        """Maximal contiguous [start, end) element runs of parameters that received a gradient: torch's AdamW skips
        grad-less parameters (e.g. MixerBlock.token_mix*, vit.py:269-271) entirely — no weight decay either."""
        runs, cur = [], None
        views = fb["views"]
        for i, ((p, off, k), has) in enumerate(zip(views, key)):
            end = views[i + 1][1] if i + 1 < len(views) else fb["n"]
            if has:
                cur = [off, end] if cur is None else [cur[0], end]
            elif cur is not None:
                runs.append(tuple(cur))
                cur = None
        if cur is not None:
            runs.append(tuple(cur))
        return runs

    def _learn_expected(self, active_key):
        self._expected_key = active_key
        self._expected, self._expected_counts = set(), {}
        idx = 0
        for fi, fb in enumerate(self._flat):
            for ri in range(len(fb["ranges"])):
                self._expected_counts[(fi, ri)] = 0
            for (p, off, k) in fb["views"]:
                if active_key[idx]:
                    self._expected.add(id(p))
                    self._expected_counts[(fi, self._where[id(p)][1])] += 1
                else:
                    fb["g"][off:off + k].zero_()           # never written: keep it out of the norm
                idx += 1
        for key in [k for k, v in self._expected_counts.items() if v == 0]:
            del self._expected_counts[key]                 # a range without gradients never completes: step() reduces it

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._pending is None:                 # no hook ran during this backward: nothing was reduced yet
            self.last_num_buckets = 0
            self.last_overlapped_buckets = 0
        self.advance()
        self.launch()
        return loss

    # ------------------------------------------------------------------ checkpointing (reference main.py:317-330 saves
    # optimizer.state_dict()): loaded moments go back INTO the flat buffers and the views are re-established
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        step = 0
        with torch.no_grad():
            for fb in self._flat:
                for (p, off, k) in fb["views"]:
                    st = self.state.get(p, {})
                    for name, flat in (("exp_avg", fb["m"]), ("exp_avg_sq", fb["v"])):
                        t = st.get(name)
                        view = flat[off:off + k].view(p.shape)
                        if t is not None and t.data_ptr() != view.data_ptr():
                            view.copy_(t.to(view.dtype))
                        st[name] = view
                    s = st.get("step")
                    if s is not None:
                        step = max(step, int(float(s)))
                    st["step"] = self._step_t
                    self.state[p] = st
        self._step_count = step
        self._step_t.fill_(float(step))

    def grad_norm(self):
        """Global gradient norm of the last step (device -> host sync; diagnostics only)."""
        _, world = D.world()
        return float(self._stats.sqrt().item()) / world if self._stats is not None else 0.0
