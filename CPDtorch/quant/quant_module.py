"""Import-path compatibility: ``from CPDtorch.quant.quant_module import Quantizer, Quant_Linear, Quant_Conv``
(reference: CPDtorch/quant/quant_module.py:10).  The implementations live in cpd_b200.quant.modules."""
from cpd_b200.quant.modules import Quant_Conv, Quant_Linear, Quant_LinearFunction, Quantizer

__all__ = ["Quantizer", "Quant_Linear", "Quant_Conv", "Quant_LinearFunction"]
