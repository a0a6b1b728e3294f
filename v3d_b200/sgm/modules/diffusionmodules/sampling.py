from v3d_b200.sampling import EulerEDMSampler  # noqa: F401  (reference: sampling.py:24-133,214-218)
