from v3d_b200.sampling import IdentityGuider, LinearPredictionGuider  # noqa: F401  (reference: guiders.py:60-101)
