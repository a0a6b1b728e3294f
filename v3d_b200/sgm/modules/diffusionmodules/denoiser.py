from v3d_b200.sampling import Denoiser  # noqa: F401  (reference: denoiser.py:12-39)
