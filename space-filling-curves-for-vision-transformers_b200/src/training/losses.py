"""SoftTargetCrossEntropy — the criterion the reference defines in main.py:45-51 (not importable from there without
running the script), provided for callers of train_with_mixup_or_cutmix: -(targets * log_softmax(logits)).sum(-1).mean().
Logits are [B, num_classes] (tiny): plain torch ops on the kernel path's output tensor."""
import torch.nn as nn
import torch.nn.functional as F


class SoftTargetCrossEntropy(nn.Module):
    def forward(self, inputs, targets):
        return -(targets * F.log_softmax(inputs.float(), dim=-1)).sum(dim=-1).mean()
