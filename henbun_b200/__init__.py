"""henbun_b200: B200-native Monte-Carlo ELBO hot path of Henbun behind Henbun's Python surface
(Henbun/__init__.py exposes the same module names).

    import henbun_b200 as hb
    import henbun_b200.tf as tf      # tf-shaped namespace for objectives written for the reference

The compute path is libhenbun_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/henbun_b200.h).  There is no CPU fallback: evaluating anything without the built library and
a CUDA device raises.
"""
from . import _settings
from ._settings import settings
from . import transforms, densities, priors
from . import tf_wraps
from . import param, model, variationals
from . import nn
from . import gp
from . import train
from . import tf  # noqa: F401

# names used by BASELINE.json's wording
param.Param = param.Variable

__version__ = "0.1.0"
