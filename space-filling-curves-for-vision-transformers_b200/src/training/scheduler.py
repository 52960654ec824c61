"""Linear-warmup + cosine learning-rate schedule — mirror of the reference's src/training/scheduler.py:4-50
(host scalar math; `step()` applies the rate for the current step, advances, and returns the rate, which
train_with_scheduler prints)."""
import math


class WarmupCosineScheduler:
    def __init__(self, optimizer, warmup_steps, total_steps, min_lr=1e-6, base_lr=None):
        self.optimizer = optimizer
        self.warmup_steps = warmup_steps
        self.total_steps = total_steps
        self.min_lr = min_lr
        self.base_lr = optimizer.param_groups[0]["lr"] if base_lr is None else base_lr
        self.current_step = 0

    def _lr_at(self, s):
        if s < self.warmup_steps:
            return self.base_lr * (s / max(1, self.warmup_steps))
        progress = (s - self.warmup_steps) / max(1, (self.total_steps - self.warmup_steps))
        return self.min_lr + 0.5 * (self.base_lr - self.min_lr) * (1 + math.cos(math.pi * min(1.0, progress)))

    def step(self):
        lr = self._lr_at(self.current_step)
        for group in self.optimizer.param_groups:
            group["lr"] = lr
        self.current_step += 1
        return lr
