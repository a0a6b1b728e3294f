"""fp32 CPU restatement of the conditioning assembly of V3D_512 (TEST INFRASTRUCTURE).

scripts/pub/V3D_512.py:31-69 (get_batch), :247-262 (conditioner call + per-frame repeat) over GeneralConditioner
(sgm/modules/encoders/modules.py:85-206) with the embedders of scripts/pub/configs/V3D_512.yaml:59-86:
IdentityEncoder(cond_frames_without_noise) -> crossattn, ConcatTimestepEmbedderND(256) x {fps_id, motion_bucket_id,
cond_aug} -> vector (concatenated in that order), IdentityEncoder(cond_frames) -> concat.  The CLIP embedding and the
noised first-stage latent are inputs (they are computed before get_batch, V3D_512.py:238-243).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .ref_unet import timestep_embedding


def concat_timestep_embed(x: torch.Tensor, outdim: int = 256) -> torch.Tensor:
    """ConcatTimestepEmbedderND.forward (encoders/modules.py:937-953): every scalar embedded on its own
    (Timestep(outdim) = timestep_embedding, openaimodel.py:64-70), concatenated per row."""
    if x.ndim == 1:
        x = x[:, None]
    b, dims = x.shape
    emb = timestep_embedding(x.reshape(-1), outdim)
    return emb.reshape(b, dims * outdim)


def v3d_conditioning(clip_emb: torch.Tensor, cond_latent: torch.Tensor, fps_id: float, motion_bucket_id: float,
                     cond_aug: float, T: int) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(c, uc) exactly as the sampler receives them: vector [T, 768] (identical in c and uc: only the two frame
    embedders are forced to zero, V3D_512.py:255-258), crossattn [T, 1, 1024], concat [T, 4, h, w]; uc's crossattn and
    concat are zeros; both repeated "b ... -> (b t) ..." with b = 1 (V3D_512.py:259-262)."""
    one = lambda v: torch.tensor([v]).repeat(T)  # noqa: E731  get_batch: N = [1, T] -> prod(N) rows
    vec = torch.cat([concat_timestep_embed(one(fps_id)), concat_timestep_embed(one(motion_bucket_id)),
                     concat_timestep_embed(one(cond_aug))], dim=1)
    rep = lambda t: t[:1].unsqueeze(1).expand(1, T, *t.shape[1:]).reshape(T, *t.shape[1:]).clone()  # noqa: E731
    c = {"vector": vec, "crossattn": rep(clip_emb), "concat": rep(cond_latent)}
    uc = {"vector": vec.clone(), "crossattn": torch.zeros_like(c["crossattn"]),
          "concat": torch.zeros_like(c["concat"])}
    return c, uc
