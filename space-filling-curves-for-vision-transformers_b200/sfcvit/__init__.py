"""sfcvit — Python runtime over libsfcvit.so (ctypes): device-pointer marshalling, op wrappers and
autograd functions used by the ``src.*`` mirror of the reference API. No CPU fallback: every op raises
if the CUDA library is missing or a tensor is not on a CUDA device."""
from . import _lib  # noqa: F401
