"""The callers that turn the inner loop into the multi-filter workload: NESS (reference
inference/sequential/ness.py, kernels/online.py, kernels/jittering.py) and SMC2 (reference inference/sequential/smc2.py) with its particle Metropolis-Hastings rejuvenation kernel (inference/sequential/kernels/mh.py,
inference/batch/mcmc/utils.py) on top of the resident batch of filters - the state particles never leave the device."""
from .prior import Exponential, LogNormal, Normal, ParameterContext  # noqa: F401
from .smc2 import SMC2, ShardedSMC2, SMC2State  # noqa: F401
from .ness import NESS, FixedWidthNESS, NonShrinkingKernel, ShrinkingKernel, LiuWestShrinkage, ConstantKernel  # noqa: F401
