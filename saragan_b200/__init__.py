"""saragan_b200 -- B200-native (sm_100a) implementation of saraGAN's data-parallel hot path:
the 3D progressive-GAN train step behind the reference's ``pgan_pytorch/network.py`` /
``loss.py`` / ``train.py`` API.  See DESIGN.md and INTEGRATION.md.
"""
from . import config  # noqa: F401
from . import metrics, network_dict  # noqa: F401  (pgan_pytorch/metrics and network_dict.py drop-ins)
from .config import set_precision, use_precision  # noqa: F401
from .loss import compute_gradient_penalty, wasserstein_loss  # noqa: F401
from .network import (ChannelNormalization, Discriminator, DiscriminatorBlock,  # noqa: F401
                      EqualizedConv3d, EqualizedLinear, FromRGB, Generator, GeneratorBlock,
                      MinibatchStandardDeviation, ToRGB, num_filters)
from .optim import FusedAdam  # noqa: F401
from .train import make_optimizers, train_epoch, train_step  # noqa: F401

__version__ = "0.1.0"
