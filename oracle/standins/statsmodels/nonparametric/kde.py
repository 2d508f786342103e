class KDEUnivariate: pass
