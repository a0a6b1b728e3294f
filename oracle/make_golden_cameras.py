"""Golden camera paths from the REAL reference (TEST INFRASTRUCTURE; run in the build container, where /root/reference
exists):  recon/utils/camera_utils.py:get_uniform_poses (:130-151, through get_c2w_from_up_and_look_at :102-127) - the
frame <-> camera convention every consumer of the generated views assumes (recon/scene/dataset_readers.py:447-477 with
recon/arguments/__init__.py:64-67: 18 frames, radius 2, elevation 0, fov 60).  The module's own imports (mediapy, the
3DGS scene package) are irrelevant to these two pure-numpy functions and are stubbed.

    python oracle/make_golden_cameras.py   ->  tests/golden/cameras.npz
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REF = Path("/root/reference/recon/utils/camera_utils.py")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "cameras.npz"


def load_reference():
    stubs = {"mediapy": ["read_video", "write_video"], "scene": [], "scene.cameras": ["Camera"], "utils": [],
             "utils.general_utils": ["PILtoTorch"], "utils.graphics_utils": ["fov2focal"]}
    saved = {k: sys.modules.get(k) for k in stubs}
    try:
        for name, attrs in stubs.items():
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, None)
            sys.modules[name] = m
        spec = importlib.util.spec_from_file_location("_ref_camera_utils", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main():
    ref = load_reference()
    cases = {"t18_r2_e0": (18, 2.0, 0.0, False), "t24_r2_e0": (24, 2.0, 0.0, False),
             "t18_r1p5_e15_gl": (18, 1.5, 15.0, True), "t14_r2_em10": (14, 2.0, -10.0, False)}
    out = {}
    for name, (t, r, e, gl) in cases.items():
        out[name] = ref.get_uniform_poses(t, r, e, opengl=gl).astype(np.float32)
        out[name + "_args"] = np.array([t, r, e, float(gl)], dtype=np.float64)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items() if not k.endswith("_args")})


if __name__ == "__main__":
    main()
