"""Drop-in replacement for the reference's ``quant_transformer.quantization`` package
(quantization/__init__.py:1-4 exports the same four groups of names)."""
from .quantized_module import Quantizer  # noqa: F401
from .quantized_module import QuantizedModule  # noqa: F401
from .state import enable_calibration_quantization, enable_calibration_woquantization, \
    enable_quantization, disable_all  # noqa: F401
