from v3d_b200.engine import DiffusionEngine  # noqa: F401  (reference: video_diffusion.py:34-210, inference side)
