"""Import-compatible façade: ``from CPDtorch.quant import float_quantize`` and
``from CPDtorch.utils.dist_util import sum_gradients, dist_init, DistModule`` work as in
drcut/CPD, served by the B200-native ``cpd_b200`` package."""
from . import quant, utils  # noqa: F401
