from scipy.special import erf, erfc, gammaln  # noqa: F401
