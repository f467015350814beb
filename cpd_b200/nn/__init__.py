from .fused_bn import BatchNormAct2d, fuse_bn_act  # noqa: F401
