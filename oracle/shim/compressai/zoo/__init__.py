from .image import bmshj2018_factorized, bmshj2018_hyperprior, model_architectures  # noqa: F401
