from .functional import float_quantize, quantizer, quant_gemm, set_default_gemm_mode, tile_gemm  # noqa: F401
from .modules import Quantizer, Quant_Linear, Quant_Conv, Quant_LinearFunction, Quant_ConvTileFunction  # noqa: F401

__all__ = ["float_quantize", "quantizer", "Quantizer", "quant_gemm", "Quant_Linear", "Quant_Conv", "set_default_gemm_mode"]
