"""TEST INFRASTRUCTURE - CPU restatement of the theta-level arithmetic of the reference's SMC2.  Only ``tests/`` may
import this module.  Every function cites the reference lines it follows (paths relative to /root/reference/pyfilter/); pinned against
the UNMODIFIED reference's own functions by ``tests/test_oracle_pinned.py::test_smc2_oracle_*`` when /root/reference is present, and by
the golden vectors ``tests/golden/smc2_theta.npz`` written by ``oracle/make_golden_smc2.py`` from the live reference."""
import math

import torch


def get_ess(w: torch.Tensor) -> torch.Tensor:
    """``utils.py:8-20`` on log-weights (``SequentialAlgorithmState.append``, inference/sequential/state.py:43-44)."""
    W = normalize(w)
    return 1.0 / (W * W).sum(0)


def normalize(w: torch.Tensor) -> torch.Tensor:
    """``utils.py:49-64``."""
    w = w.clone().nan_to_num_(-float("inf"), posinf=-float("inf"))
    m = w.max(0).values
    e = (w - m).exp()
    s = e.sum(0)
    out = e / s
    if out.dim() == 1:
        return out if float(s) > 0 else torch.full_like(out, 1.0 / out.shape[0])
    return out


def calc_mean_chol(x: torch.Tensor, w: torch.Tensor):
    """``inference/utils.py:42-58``."""
    mean = w @ x
    centralized = x - mean
    cov = (w * centralized.t()).matmul(centralized)
    chol, info = torch.linalg.cholesky_ex(cov)
    if (info > 0).any():
        chol = cov.diag().sqrt().diag()
    return mean, chol


def construct_mvn(x: torch.Tensor, w: torch.Tensor, scale: float = 1.0):
    """``inference/utils.py:61-76`` (``quasi_engine=None``): ``MultivariateNormal(mean, scale_tril=scale * chol)``."""
    mean, chol = calc_mean_chol(x, w)
    return torch.distributions.MultivariateNormal(mean, scale_tril=scale * chol, validate_args=False)


def symmetric_proposal(values: torch.Tensor, normalized_weights: torch.Tensor):
    """``SymmetricMH.build`` (inference/batch/mcmc/proposals/symmetric_mh.py:13-23): scale 1.1."""
    return construct_mvn(values, normalized_weights, scale=1.1)


def run_pmmh_acceptance(values_old, values_new, kernel, new_kernel, prior_old, prior_new, ll_old, ll_new, u):
    """``run_pmmh`` (inference/batch/mcmc/utils.py:56-67) from the quantities it compares: ``diff_logl``, ``diff_prior``,
    ``diff_prop = new_kernel.log_prob(theta) - kernel.log_prob(theta*)``; accept where ``log u < sum``."""
    diff_logl = ll_new - ll_old
    diff_prior = prior_new - prior_old
    diff_prop = new_kernel.log_prob(values_old) - kernel.log_prob(values_new)
    log_acc = diff_prop + diff_prior + diff_logl
    return u.log() < log_acc, log_acc


def smc2_needs_rejuvenation(w: torch.Tensor, threshold: float) -> bool:
    """``SMC2._step`` (inference/sequential/smc2.py:59-63) with a ``ConstantThreshold``."""
    any_nans = not bool(torch.isfinite(w).all())
    return bool(get_ess(w) < threshold * w.shape[0]) or any_nans


def normal_unconstrained_log_prob(u: torch.Tensor, loc: float, scale: float) -> torch.Tensor:
    """``PriorMixin.eval_prior(x, constrained=False)`` (inference/prior.py:81-90) for ``Normal`` (identity bijection) and for
    ``LogNormal`` (``exp`` bijection: the unconstrained prior of ``log theta`` is ``Normal(loc, scale)``)."""
    return -((u - loc) ** 2) / (2.0 * scale * scale) - math.log(scale) - 0.5 * math.log(2.0 * math.pi)


# ---- NESS: the jittering kernels (inference/sequential/kernels/jittering.py) -----------------------------------------------------------
EPS = math.sqrt(torch.finfo(torch.float32).eps)   # constants.py


def robust_var(x: torch.Tensor, w: torch.Tensor, mean: torch.Tensor = None) -> torch.Tensor:
    """``jittering.py:49-83``."""
    sort, sort_indices = x.sort(0)
    cumulative_weights = w[sort_indices].cumsum(0)
    low_indices = (cumulative_weights - 0.25).abs().argmin(0)
    high_indices = (cumulative_weights - 0.75).abs().argmin(0)
    iqr = (sort[high_indices].diag() - sort[low_indices].diag()) / 1.349
    iqr2 = iqr**2
    w = w.unsqueeze(-1)
    if mean is None:
        mean = (w * x).sum(0)
    var = (w * (x - mean) ** 2).sum(0)
    mask = iqr2 <= var
    if mask.any():
        var[mask] = iqr2[mask]
    return var


def jitter_fit(kind: str, x: torch.Tensor, w: torch.Tensor, indices: torch.Tensor, a: float = 0.98, scale: float = 0.1):
    """``fit`` of ShrinkingKernel (jittering.py:148-158), NonShrinkingKernel (:166-173), LiuWestShrinkage (:197-203), ConstantKernel
    (:222-225): the location and the scale of the jitter; ``jitter`` (:117-134) adds ``max(scale, EPS) * eps`` to the location."""
    ess = 1.0 / (w * w).sum()
    bw_fac = (1.59 * ess ** (-1 / 3)).clamp(EPS, 1 - EPS)
    if kind == "shrinking":
        mean = (w.unsqueeze(-1) * x).sum(0)
        var = robust_var(x, w, mean)
        beta = (1.0 - bw_fac**2).sqrt()
        return (mean + beta * (x - mean))[indices], bw_fac * var.sqrt()
    if kind == "nonshrinking":
        return x[indices], bw_fac * robust_var(x, w).sqrt()
    if kind == "liuwest":
        mean = (w.unsqueeze(-1) * x).sum(0)
        var = robust_var(x, w, mean)
        return (x * a + (1 - a) * mean)[indices], math.sqrt(1 - a**2) * var.sqrt()
    if kind == "constant":
        return x[indices], torch.as_tensor(scale)
    raise ValueError(kind)
