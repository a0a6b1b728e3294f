"""CPU oracle for the V3D denoising hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32 restatement of the reference's algorithm for the path named in BASELINE.json
(EulerEDMSampler -> VideoUNet -> VideoDecoder), written functionally over a reference-keyed state_dict.
Every function cites the reference file:line it follows (paths relative to /root/reference).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package, and only as the checker / timed baseline.  Nothing under `v3d_b200/` imports it.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against *the reference modules themselves*, imported in the build container through
`oracle/reference_shim.py`; `oracle/make_golden.py` asserts oracle == reference (fp32, <= 2e-4 max-abs
relative to output scale) on every fixture it writes to tests/golden/.  Parity on the real pretrained
checkpoints is unpinned (weights are not available offline): synthetic seeded weights are the vehicle.
"""
