"""cpd_b200 — Blackwell-native custom-precision distributed training (capabilities of drcut/CPD)."""
from .config import PrecisionConfig  # noqa: F401
from .quant import float_quantize, quantizer, Quantizer, quant_gemm, Quant_Linear, Quant_Conv  # noqa: F401
from .parallel import (dist_init, DistModule, sum_gradients, normal_sum_gradients, kahan_sum_gradients,  # noqa: F401
                       broadcast_params, GradReducer, ModelArena, EmulateNode, discover_topology)

__version__ = "0.1.0"
