"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference on CPU.

Only usable in the build container (the GPU box has no /root/reference); used by
tests/golden/gen_golden.py to produce the committed golden vectors and by optional
``-m "not gpu"`` tests that re-validate the oracle when the reference is present.

Shims, none of which touches reference arithmetic:
  1. quantization/observer.py:6,8 import seaborn / matplotlib.pyplot (unused) -> empty stub modules;
  2. observer.py:81,95,425-426,481,494 hard-code ``.cuda()`` -> identity on this CPU-only process.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OSQ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "quant_transformer", "quantization"))


def load():
    """Returns the reference ``quant_transformer.quantization`` package (CPU-shimmed)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    for name in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2 (CPU-only process)
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import quant_transformer.quantization as q  # noqa: E402
    import quant_transformer.quantization.fake_quant  # noqa: F401,E402
    import quant_transformer.quantization.observer  # noqa: F401,E402
    import quant_transformer.quantization.quantized_module  # noqa: F401,E402
    import quant_transformer.quantization.state  # noqa: F401,E402
    import quant_transformer.quantization.util_quant  # noqa: F401,E402
    return q


class QConfig:
    """Attribute-style stand-in for the EasyDict node (easydict is not installed)."""

    def __init__(self, quantizer, observer, bit, symmetric, ch_axis):
        self.quantizer, self.observer, self.bit, self.symmetric, self.ch_axis = quantizer, observer, bit, symmetric, ch_axis
