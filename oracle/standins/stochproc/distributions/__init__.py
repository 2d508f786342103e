"""Empty stand-in: only the reference's tests/examples use ``stochproc.distributions``."""
