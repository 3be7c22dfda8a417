from scipy.stats import norm, multivariate_normal  # noqa: F401
