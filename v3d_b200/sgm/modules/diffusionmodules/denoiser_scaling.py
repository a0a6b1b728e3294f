from v3d_b200.sampling import VScalingWithEDMcNoise  # noqa: F401  (reference: denoiser_scaling.py:51-59)
