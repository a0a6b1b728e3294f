from v3d_b200.sampling import (CentralPredictionGuider, IdentityGuider, LinearPredictionGuider,  # noqa: F401
                               VanillaCFG)  # reference: guiders.py:23-42,60-101,104-146
