from v3d_b200.decoder import AutoencodingEngine  # noqa: F401  (reference: autoencoder.py:128-212, decode side)
