from v3d_b200.sampling import EulerEDMSampler, HeunEDMSampler  # noqa: F401  (reference: sampling.py:24-133,214-237)
