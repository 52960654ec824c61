"""SoftTargetCrossEntropy — the criterion the reference defines in main.py:45-51 (not importable from there without
running the script), provided for callers of train_with_mixup_or_cutmix: -(targets * log_softmax(logits)).sum(-1).mean().
One libsfcvit kernel forward (row log-sum-exp, target-weighted sum and a fixed-order batch mean) and one backward
(csrc/softce.cu) instead of six ATen launches each way; CUDA tensors only."""
import torch.nn as nn

from sfcvit import functional as SF


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, inputs, targets):
        return SF.soft_target_cross_entropy(inputs, targets)
