"""``src`` — drop-in mirror of the reference's Python API (src.curves, src.tokenizers, src.models, src.training) so that
the reference's main.py / run_vit.sh drive the B200-native implementation unchanged (main.py:25-42 import paths).

This top-level package is a path shim: the implementation lives in
``space-filling-curves-for-vision-transformers_b200/src`` next to the CUDA sources and the ``sfcvit`` runtime."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PK = os.path.join(_ROOT, "space-filling-curves-for-vision-transformers_b200")
if _PK not in sys.path:
    sys.path.insert(0, _PK)
__path__ = [os.path.join(_PK, "src")]
