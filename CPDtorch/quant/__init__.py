from cpd_b200.quant import (float_quantize, quantizer, Quantizer, quant_gemm, Quant_Linear, Quant_Conv)  # noqa: F401

__all__ = ["float_quantize", "quantizer", "Quantizer", "quant_gemm", "Quant_Linear", "Quant_Conv"]
