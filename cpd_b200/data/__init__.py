from .synthetic import SyntheticImageDataset, cifar_shape, imagenet_shape, PinnedBatchStream, Batches  # noqa: F401
from .samplers import GivenIterationSampler, DistributedGivenIterationSampler, DistributedSampler  # noqa: F401
from .augment import normalise, pad, transpose, Crop, FlipLR, Cutout, Transform  # noqa: F401
