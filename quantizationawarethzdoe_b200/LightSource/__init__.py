from .Gaussian_beam import Guassian_beam  # noqa: F401
