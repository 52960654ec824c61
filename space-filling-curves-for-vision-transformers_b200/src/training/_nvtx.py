"""Optional NVTX ranges around the phases of a training step (tracing). Off by default; `SFC_NVTX=1` turns
them on, so that an Nsight Systems / `ncu --nvtx` timeline shows "batch prep", "forward+backward" (one graph replay when
the step is captured), "allreduce", "clip+optimizer" per step. No effect on what is launched."""
import contextlib
import os

import torch

_ON = os.environ.get("SFC_NVTX") == "1"


def enabled():
    return _ON and torch.cuda.is_available()


@contextlib.contextmanager
def span(name):
    if not enabled():
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()
