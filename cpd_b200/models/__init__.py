from .resnet_cifar import resnet18_cifar, resnet34_cifar, resnet18_cifar_quant, ResNetCifar  # noqa: F401
from .resnet_imagenet import resnet18, resnet34, resnet50, resnet101, ResNet  # noqa: F401
from .fcn import FCN, FCNHead, fcn_r50_d8, fcn_r18_d8  # noqa: F401
from .davidnet import davidnet, GraphNet, graph_from_nested, rel, Identity, Concat, Correct, Scale, Add, batch_norm  # noqa: F401

_REGISTRY = {
    "resnet18_cifar": resnet18_cifar, "res_cifar": resnet18_cifar, "resnet34_cifar": resnet34_cifar,
    "resnet18_cifar_quant": resnet18_cifar_quant,
    "resnet18": resnet18, "resnet34": resnet34, "resnet50": resnet50, "resnet101": resnet101,
    "davidnet": davidnet,
    "fcn_r50_d8": fcn_r50_d8, "fcn_r18_d8": fcn_r18_d8,
}


def create_model(name: str, **kw):
    try:
        return _REGISTRY[name](**kw)
    except KeyError:
        raise ValueError(f"unknown model {name!r}; available: {sorted(_REGISTRY)}") from None
