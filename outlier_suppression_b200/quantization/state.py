"""Model-wide observer / fake-quant togglers (quantization/state.py:7-69): same names, same
substring matching on the module path, same logger."""
import logging

from .fake_quant import LSQFakeQuantize, LSQPlusFakeQuantize, QuantizeBase
from .observer import ObserverBase

logger = logging.getLogger("transformer")


def _drop_weight_caches(model):
    # weights may have been rewritten through `.data` (gamma migration) since the last forward
    for m in model.modules():
        inv = getattr(m, "invalidate_packed", None)
        if inv is not None:
            inv()
    # query | key | value style siblings share one fused launch from here on (quantized_module.QLinearGroup)
    from .quantized_module import group_sibling_linears
    group_sibling_linears(model)
    # dense -> GELU -> quantizer blocks get the fused output stage in the quantized state (quantization/fusion.py)
    from .fusion import fuse_ffn_activation, fuse_layernorm_output, fuse_self_attention
    fuse_ffn_activation(model)
    fuse_layernorm_output(model)
    fuse_self_attention(model)


def _apply(model, quantizer_type, except_quantizer, observer_on, fq_on, lsq_observer_off=False):
    _drop_weight_caches(model)
    for name, sub in model.named_modules():
        if not isinstance(sub, QuantizeBase):
            continue
        if (quantizer_type not in name) or (except_quantizer is not None and name in except_quantizer):
            logger.debug("The except_quantizer is {}".format(name))
            sub.disable_observer()
            sub.disable_fake_quant()
            continue
        if observer_on and not (lsq_observer_off and isinstance(sub, (LSQFakeQuantize, LSQPlusFakeQuantize))):
            sub.enable_observer()
        else:
            sub.disable_observer()
            if observer_on:
                logger.info("Extrally disable observer for LSQ/LSQPlusFakeQuantize during training!")
        sub.enable_fake_quant() if fq_on else sub.disable_fake_quant()


def enable_calibration_woquantization(model, quantizer_type="fake_quant", except_quantizer=None):
    logger.info("Enable observer and Disable quantize for {}".format(quantizer_type))
    _apply(model, quantizer_type, except_quantizer, observer_on=True, fq_on=False)


def enable_calibration_quantization(model, quantizer_type="fake_quant", except_quantizer=None):
    logger.info("Enable observer and Enable quantize for {}".format(quantizer_type))
    _apply(model, quantizer_type, except_quantizer, observer_on=True, fq_on=True, lsq_observer_off=True)


def enable_quantization(model, quantizer_type="fake_quant", except_quantizer=None):
    logger.info("Disable observer and Enable quantize.")
    _apply(model, quantizer_type, except_quantizer, observer_on=False, fq_on=True)


def disable_all(model):
    logger.info("Disable observer and disable quantize.")
    _drop_weight_caches(model)
    for _, sub in model.named_modules():
        if isinstance(sub, QuantizeBase):
            sub.disable_observer()
            sub.disable_fake_quant()


def set_observer_name(model):
    logger.info("set name for obsever")
    for name, sub in model.named_modules():
        if isinstance(sub, ObserverBase):
            sub.set_name(name)
