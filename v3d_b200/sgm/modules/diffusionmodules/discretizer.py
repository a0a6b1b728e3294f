from v3d_b200.sampling import EDMDiscretization  # noqa: F401  (reference: discretizer.py:18-39)
