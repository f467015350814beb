"""Import-path compatibility: ``from CPDtorch.quant.quant_function import float_quantize, quantizer, quant_gemm``
(reference: CPDtorch/quant/quant_function.py:21).  The implementations live in cpd_b200.quant.functional."""
from cpd_b200.quant.functional import float_quantize, quant_gemm, quantizer

__all__ = ["float_quantize", "quantizer", "quant_gemm"]
