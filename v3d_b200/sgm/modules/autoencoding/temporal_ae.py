from v3d_b200.decoder import VideoDecoder  # noqa: F401  (reference: temporal_ae.py:293-349)
