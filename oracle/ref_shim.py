"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference on CPU.

Where from: /root/reference in the build container; on the GPU box the verbatim copy that ``oracle/make_ref.py``
staged under the git-ignored ``oracle/_ref/`` (it travels with the snapshot like the built .so files).  Used by
tests/golden/gen_golden.py (golden vectors), by the model-level drop-in tests, and by ``bench.py --impl reference``.

Shims, none of which touches reference arithmetic:
  1. quantization/observer.py:6,8 import seaborn / matplotlib.pyplot (unused) -> empty stub modules;
  2. observer.py:81,95,425-426,481,494 and gamma_migration.py:67 hard-code ``.cuda()`` -> identity while the
     reference runs as the CPU oracle (``cpu_only()`` context; permanent in a process without CUDA);
  3. transformers-4.18 symbols the model files import (``compat_transformers()``).
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

from . import make_ref

REFERENCE_ROOT = make_ref.root() or os.environ.get("OSQ_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return make_ref.root() is not None


def _stub_plot_modules():
    for name in ("seaborn", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))


@contextlib.contextmanager
def cpu_only():
    """``tensor.cuda()`` / ``module.cuda()`` are identities inside this block (shim 2): lets the reference run as the
    CPU oracle in a process that also has a GPU."""
    import torch
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def compat_transformers():
    """transformers 4.18 symbols the reference model files import (shim 3)."""
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for name in ("apply_chunking_to_forward", "prune_linear_layer", "find_pruneable_heads_and_indices"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name, lambda *a, **k: (set(), None)))
    if "transformers.generation_utils" not in sys.modules:
        g = types.ModuleType("transformers.generation_utils")
        g.GenerationMixin = transformers.generation.GenerationMixin
        sys.modules["transformers.generation_utils"] = g


def purge():
    """Forget every ``quant_transformer*`` module so the package can be re-imported bound to another backend."""
    for name in [n for n in sys.modules if n == "quant_transformer" or n.startswith("quant_transformer.")]:
        del sys.modules[name]


def load():
    """Returns the reference ``quant_transformer.quantization`` package (CPU-shimmed)."""
    root = make_ref.root()
    if root is None:
        raise RuntimeError("reference tree not present (neither /root/reference nor a staged oracle/_ref)")
    import torch

    _stub_plot_modules()
    if root not in sys.path:
        sys.path.insert(0, root)
    cur = sys.modules.get("quant_transformer.quantization")
    if cur is not None and not (getattr(cur, "__file__", "") or "").startswith(root):
        purge()  # the name is currently aliased to another backend (install_as_reference_backend)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # shim 2 (CPU-only process)
        torch.nn.Module.cuda = lambda self, *a, **k: self
    import quant_transformer.quantization as q  # noqa: E402
    import quant_transformer.quantization.fake_quant  # noqa: F401,E402
    import quant_transformer.quantization.observer  # noqa: F401,E402
    import quant_transformer.quantization.quantized_module  # noqa: F401,E402
    import quant_transformer.quantization.state  # noqa: F401,E402
    import quant_transformer.quantization.util_quant  # noqa: F401,E402
    if not (getattr(q, "__file__", "") or "").startswith(root):
        raise RuntimeError("quant_transformer.quantization resolved to %s, not the reference: call ref_shim.purge() first"
                           % getattr(q, "__file__", None))
    return q


class QConfig:
    """Attribute-style stand-in for the EasyDict node (easydict is not installed)."""

    def __init__(self, quantizer, observer, bit, symmetric, ch_axis):
        self.quantizer, self.observer, self.bit, self.symmetric, self.ch_axis = quantizer, observer, bit, symmetric, ch_axis
