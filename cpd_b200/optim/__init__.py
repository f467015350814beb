from .fused import FusedSGD, bn_bias_no_decay  # noqa: F401
from .schedules import WarmupStepLR, PiecewiseLinear, IterLRScheduler, linear_warmup  # noqa: F401
