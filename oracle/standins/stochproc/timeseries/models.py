from math import sqrt

import torch
from torch.distributions import Normal

from .process import AffineProcess


def _ar_mean_scale(x, alpha, beta, sigma):
    return alpha + beta * x.value, sigma


def _ar_init(alpha, beta, sigma):
    return Normal(alpha, sigma / (1.0 - beta**2.0).sqrt())


class AR(AffineProcess):
    """AR(1): ``x_t = alpha + beta x_{t-1} + sigma eps`` with the stationary initial law."""

    def __init__(self, alpha, beta, sigma):
        super().__init__(_ar_mean_scale, (alpha, beta, sigma), Normal(0.0, 1.0), _ar_init)


class RandomWalk(AffineProcess):
    def __init__(self, sigma, initial_mean=0.0):
        super().__init__(
            lambda x, s: (x.value, s), (sigma,), Normal(0.0, 1.0), lambda s: Normal(torch.tensor(initial_mean), s)
        )
