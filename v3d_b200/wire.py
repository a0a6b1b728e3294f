"""The step right after the hot path: the wire format of the generated views and the frame <->
camera convention their consumers assume.

  * `frames_to_u8`: `clamp((x + 1) / 2, 0, 1) * 255 -> uint8`, `t c h w -> t h w c`
    (scripts/pub/V3D_512.py:286-303) in one kernel (`v3d_frames_nchw_to_u8`), so only 14 MB per image cross PCIe;
  * `write_video`: the mp4 `sample_one` saves (`mediapy.write_video(path, frames, fps=3)`, V3D_512.py:304-306), through
    OpenCV when it is installed;
  * `orbit_poses` / `camera_infos`: frame t of the T generated views is the camera at azimuth 360 t / T on a circle of
    radius 2 around the object, elevation 0, looking at the origin, z up, 60 degree field of view - what
    recon/utils/camera_utils.py:130-151 (`get_uniform_poses`) and recon/scene/dataset_readers.py:447-477
    (`constructVideoNVSInfo`, defaults recon/arguments/__init__.py:64-67) hand to the 3-D reconstruction.  Host-side
    numpy; pinned against the reference's own function in tests/golden/cameras.npz.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops


def frames_to_u8(frames: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[T,3,H,W] fp32 in [-1,1] (decode_first_stage's output) -> [T,H,W,3] uint8 on the same device."""
    t, c, h, w = frames.shape
    assert c == 3
    if out is None:
        out = torch.empty(t, h, w, 3, device=frames.device, dtype=torch.uint8)
    return ops.frames_nchw_to_u8(frames.float().contiguous(), out)


def write_video(path: str, frames_u8, fps: int = 3) -> str:
    """frames_u8: [T,H,W,3] uint8 RGB (tensor or array).  mp4 at 3 frames/s like the reference."""
    try:
        import cv2
    except ImportError as e:  # pragma: no cover
        raise RuntimeError("write_video needs OpenCV (cv2); the reference uses mediapy, which is not installed") from e
    arr = frames_u8.cpu().numpy() if torch.is_tensor(frames_u8) else np.asarray(frames_u8)
    assert arr.ndim == 4 and arr.shape[-1] == 3 and arr.dtype == np.uint8
    wr = cv2.VideoWriter(str(path), cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (arr.shape[2], arr.shape[1]))
    if not wr.isOpened():
        raise RuntimeError(f"cannot open {path} for writing")
    for f in arr:
        wr.write(np.ascontiguousarray(f[..., ::-1]))   # OpenCV wants BGR
    wr.release()
    return str(path)


def orbit_poses(num_frames: int = 18, radius: float = 2.0, elevation: float = 0.0, opengl: bool = False) -> np.ndarray:
    """Camera-to-world matrices [T,4,4] fp32 of the T views (OpenCV axes: x right, y down, z forward; `opengl` flips y
    and z).  The camera of frame t sits at azimuth 360 t / T, looks at the origin, world z is up."""
    az = np.deg2rad(np.linspace(0.0, 360.0, num_frames + 1)[:num_frames])
    el = np.full_like(az, np.deg2rad(elevation))
    dist = np.full_like(az, radius)
    pos = np.stack([dist * np.cos(el) * np.cos(az), dist * np.cos(el) * np.sin(az), dist * np.sin(el)], axis=-1)
    fwd = -pos / np.linalg.norm(pos, axis=-1, keepdims=True)               # z: towards the origin
    down = np.broadcast_to(np.array([0.0, 0.0, -1.0]), pos.shape)           # y starts as "minus up" ...
    right = np.cross(down, fwd)
    right = right / np.linalg.norm(right, axis=-1, keepdims=True)           # x
    down = np.cross(fwd, right)                                             # ... and is re-orthogonalised
    c2w = np.zeros((num_frames, 4, 4), dtype=np.float32)
    c2w[:, :3, 0], c2w[:, :3, 1], c2w[:, :3, 2], c2w[:, :3, 3] = right, down, fwd, pos
    c2w[:, 3, 3] = 1.0
    if opengl:
        c2w[:, :, 1:3] *= -1.0
    return c2w


def camera_infos(num_frames: int = 18, radius: float = 2.0, elevation: float = 0.0, fov: float = 60.0,
                 reso: int = 512) -> List[Dict]:
    """Per-frame camera records in the convention of constructVideoNVSInfo: R = transpose of the world-to-camera
    rotation, T = its translation, square images of `reso` pixels with FovX = FovY = fov degrees (radians stored)."""
    w2c = np.linalg.inv(orbit_poses(num_frames, radius, elevation))
    return [{"uid": i, "R": np.transpose(m[:3, :3]), "T": m[:3, 3], "FovX": float(np.deg2rad(fov)),
             "FovY": float(np.deg2rad(fov)), "width": reso, "height": reso, "image_name": i}
            for i, m in enumerate(w2c)]
