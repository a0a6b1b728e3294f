"""__graft_entry__.smoke(): one tiny invocation of the hot path on cuda:0, checked against the CPU oracle.

(The oracle import below is the smoke check's checker role; the product modules never import it.)
A reduced-width V3D_512-shaped VideoUNet (model_channels 64) runs 2 Euler-EDM steps with CFG through the drop-in
sampler / denoiser / wrapper, then the reduced VideoDecoder decodes the latents; both are compared with the
oracle restatement on identical seeded weights and inputs.
"""
from __future__ import annotations

import time

import torch


def run(verbose: bool = True) -> dict:
    from oracle import ref_decoder, ref_sampling, ref_unet, synth  # checker only
    from v3d_b200 import engine, ops

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    T, hw, steps = 4, 16, 2
    cfg = engine.v3d_512_config(num_frames=T, num_steps=steps, min_cfg=1.5, max_cfg=3.5)
    cfg["network_config"]["params"]["model_channels"] = 64
    cfg["first_stage_config"]["params"]["decoder_config"]["params"]["ch"] = 64
    eng = engine.DiffusionEngine(**cfg)
    unet = eng.model.diffusion_model
    dec = eng.first_stage_model.decoder
    sd_u = synth.synth_state_dict(unet.param_shapes(), seed=11)
    sd_d = synth.synth_state_dict(dec.param_shapes(), seed=12)
    unet.load_state_dict(sd_u, strict=True)
    dec.load_state_dict(sd_d, strict=True)
    eng = eng.to(dev).eval()

    x, c, uc = synth.synth_inputs(T, hw)
    launches0 = ops.launch_count()
    t0 = time.time()
    frames = eng.sample_views(x.clone().to(dev), {k: v.to(dev) for k, v in c.items()},
                              {k: v.to(dev) for k, v in uc.items()}, num_frames=T)
    torch.cuda.synchronize()
    t_gpu = time.time() - t0
    launches = ops.launch_count() - launches0

    spec_u = ref_unet.UNetSpec(model_channels=64)
    spec_d = ref_decoder.DecoderSpec(ch=64)
    extra = {"image_only_indicator": torch.zeros(2, T), "num_video_frames": T}
    with torch.no_grad():
        z_ref = ref_sampling.euler_edm_sample(
            lambda i, s, cc: ref_sampling.denoiser(
                lambda xx, tt, cond, **kw: ref_unet.openai_wrapper(sd_u, spec_u, xx, tt, cond, **kw), i, s, cc, **extra),
            x.clone(), c, uc, steps, ref_sampling.guider_scale(1.5, 3.5, T), T)
        img_ref = ref_decoder.decode_first_stage(sd_d, spec_d, z_ref, n_samples_a_time=T)
    got = frames.float().cpu()
    rel = ((got - img_ref).norm() / img_ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(), img_ref.flatten(), dim=0).item()
    res = {"rel_l2": rel, "cosine": cos, "gpu_launches": launches, "gpu_seconds": t_gpu,
           "frames_shape": list(frames.shape)}
    if verbose:
        print("smoke:", res)
    assert launches > 0, "no native kernels were launched"
    assert torch.isfinite(got).all()
    # the parity tolerance of DESIGN.md section 4 (measured 2.3e-2 / 0.9997 on this composition: sampler + UNet + decode)
    assert rel <= 3e-2 and cos >= 0.999, f"smoke parity failed: {res}"
    return res
