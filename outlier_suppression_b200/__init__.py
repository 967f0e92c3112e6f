"""outlier_suppression_b200 -- B200-native (sm_100a) fake-quantize / observer / fused
fake-quant+Linear hot path of wimh966/outlier_suppression.

    outlier_suppression_b200.quantization   drop-in for quant_transformer.quantization
    outlier_suppression_b200.ops            torch-facing wrappers over the C ABI (include/osq.h)
    outlier_suppression_b200.dist           rank-sharded calibration with one packed all-reduce
    outlier_suppression_b200.install_as_reference_backend()
                                            make `import quant_transformer.quantization` resolve here
"""
import sys

__version__ = "0.1.0"


def install_as_reference_backend():
    """Aliases this package's quantization modules under the reference's import path so that
    quant_transformer/model/*.py and solver/*.py (which `from quant_transformer.quantization import ...`)
    run unchanged on the B200 path.  Call before importing quant_transformer.model."""
    import importlib
    import types

    from . import quantization as q

    pkg = sys.modules.get("quant_transformer")
    if pkg is None:
        try:
            pkg = importlib.import_module("quant_transformer")
        except ImportError:
            pkg = types.ModuleType("quant_transformer")
            pkg.__path__ = []
            sys.modules["quant_transformer"] = pkg
    sys.modules["quant_transformer.quantization"] = q
    pkg.quantization = q
    for sub in ("fake_quant", "observer", "quantized_module", "state", "util_quant"):
        sys.modules["quant_transformer.quantization." + sub] = importlib.import_module(__name__ + ".quantization." + sub)
    return q
