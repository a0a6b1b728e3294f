from v3d_b200.encoder import DiagonalGaussianRegularizer  # noqa: F401  (reference: regularizers/__init__.py:13-32)
