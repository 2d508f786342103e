"""pyfilter_b200 - a B200-native Sequential-Monte-Carlo inner loop behind pyfilter's filter / proposal / resampler surface.

Host-side mirror of the reference's interface for the hot path only:

    pyfilter_b200.resampling.{systematic, multinomial}        <->  pyfilter.resampling            (resampling.py)
    pyfilter_b200.utils.{normalize, get_ess}                  <->  pyfilter.utils                 (utils.py)
    pyfilter_b200.filters.particle.{SISR, APF, proposals}     <->  pyfilter.filters.particle      (filters/particle/*.py)
    pyfilter_b200.filters.FilterResult                        <->  pyfilter.filters.FilterResult  (filters/result.py)
    pyfilter_b200.timeseries                                   <->  the stochproc model objects the reference is handed

All arithmetic runs in hand-written sm_100a CUDA kernels reached through the C ABI of ``libsmcb200.so``
(``include/smcb200.h``); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import utils, resampling, timeseries, filters  # noqa: F401

__version__ = "0.1.0"
