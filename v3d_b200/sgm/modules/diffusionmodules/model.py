from v3d_b200.encoder import Encoder  # noqa: F401  (reference: model.py:463-601)
