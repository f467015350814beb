from .functional import float_quantize, quantizer, quant_gemm  # noqa: F401
from .modules import Quantizer, Quant_Linear, Quant_Conv, Quant_LinearFunction  # noqa: F401

__all__ = ["float_quantize", "quantizer", "Quantizer", "quant_gemm", "Quant_Linear", "Quant_Conv"]
