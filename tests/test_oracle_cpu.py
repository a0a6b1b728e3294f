"""CPU tests of the oracle itself: the restatement must reproduce the committed outputs of the REAL reference
modules (tests/golden, written by oracle/make_golden.py) on this machine too."""
import json
from pathlib import Path

import pytest
import torch

from oracle import ref_decoder, ref_encoder, ref_sampling, ref_unet, synth

GOLD = Path(__file__).resolve().parent / "golden"
MANIFEST = json.loads((GOLD / "MANIFEST.json").read_text())


def _close(a, b, tol=2e-4):
    return ((a - b).abs().max() / b.abs().max()).item() <= tol


def _unet_inputs(T, hw):
    x, c, uc = synth.synth_inputs(T, hw)
    xin = torch.cat([torch.cat([x, x]), torch.cat([uc["concat"], c["concat"]])], 1)
    return x, c, uc, xin, torch.cat([uc["crossattn"], c["crossattn"]]), torch.cat([uc["vector"], c["vector"]])


@pytest.mark.parametrize("tag", ["unet_small", "unet_small_t18"])
def test_oracle_unet_vs_reference_golden(tag):
    m = MANIFEST[tag]
    gold = torch.load(GOLD / f"{tag}.pt")
    spec = ref_unet.UNetSpec(model_channels=m["model_channels"])
    sd = synth.synth_state_dict(ref_unet.unet_param_shapes(spec), seed=m["weight_seed"])
    T = m["T"]
    _, _, _, xin, ctx, y = _unet_inputs(T, m["latent_hw"])
    taps = {}
    with torch.no_grad():
        out = ref_unet.unet_forward(sd, spec, xin, gold["timesteps"], ctx, y, T, torch.zeros(2, T), taps=taps)
    assert _close(out, gold["out"])
    for k, v in gold.items():
        if k.startswith("tap:"):
            assert _close(taps[k[4:]][[0, T]][:, ::8], v.float(), tol=2e-3)  # fixture taps are fp16


def test_oracle_edm_sampler_vs_reference_golden():
    m, mu = MANIFEST["edm_small"], MANIFEST["unet_small"]
    gold = torch.load(GOLD / "edm_small.pt")
    spec = ref_unet.UNetSpec(model_channels=mu["model_channels"])
    sd = synth.synth_state_dict(ref_unet.unet_param_shapes(spec), seed=mu["weight_seed"])
    T = m["T"]
    x, c, uc = synth.synth_inputs(T, m["latent_hw"])
    extra = {"image_only_indicator": torch.zeros(2, T), "num_video_frames": T}
    with torch.no_grad():
        out = ref_sampling.euler_edm_sample(
            lambda i, s, cc: ref_sampling.denoiser(
                lambda xx, tt, cond, **kw: ref_unet.openai_wrapper(sd, spec, xx, tt, cond, **kw), i, s, cc, **extra),
            x.clone(), c, uc, m["num_steps"], ref_sampling.guider_scale(m["min_scale"], m["max_scale"], T), T)
    assert _close(out, gold["out"])


@pytest.mark.parametrize("tag", ["edm_small_heun", "edm_small_central", "edm_small_vanilla"])
def test_oracle_sampler_variants_vs_reference_golden(tag):
    """HeunEDMSampler, CentralPredictionGuider and VanillaCFG around the same denoiser."""
    m, mu = MANIFEST[tag], MANIFEST["unet_small"]
    gold = torch.load(GOLD / f"{tag}.pt")
    spec = ref_unet.UNetSpec(model_channels=mu["model_channels"])
    sd = synth.synth_state_dict(ref_unet.unet_param_shapes(spec), seed=mu["weight_seed"])
    T = m["T"]
    x, c, uc = synth.synth_inputs(T, m["latent_hw"])
    extra = {"image_only_indicator": torch.zeros(2, T), "num_video_frames": T}
    scale, nf = {"linear": (ref_sampling.guider_scale(m["min_scale"], m["max_scale"], T), T),
                 "central": (ref_sampling.central_guider_scale(m["min_scale"], m["max_scale"], T), T),
                 "vanilla": (ref_sampling.vanilla_scale(m["vanilla_scale"]), 1)}[m["guider"]]
    fn = ref_sampling.heun_edm_sample if m["sampler"] == "HeunEDMSampler" else ref_sampling.euler_edm_sample
    with torch.no_grad():
        out = fn(lambda i, s, cc: ref_sampling.denoiser(
            lambda xx, tt, cond, **kw: ref_unet.openai_wrapper(sd, spec, xx, tt, cond, **kw), i, s, cc, **extra),
            x.clone(), c, uc, m["num_steps"], scale, nf)
    assert _close(out, gold["out"])


@pytest.mark.parametrize("tag", ["decoder_small", "decoder_small_2videos", "decoder_full"])
def test_oracle_decoder_vs_reference_golden(tag):
    m = MANIFEST[tag]
    gold = torch.load(GOLD / f"{tag}.pt")
    spec = ref_decoder.DecoderSpec(ch=m["ch"])
    sd = synth.synth_state_dict(ref_decoder.decoder_param_shapes(spec), seed=m["weight_seed"])
    with torch.no_grad():
        out = ref_decoder.decoder_forward(sd, spec, gold["z"] / 0.18215, m["T"])
    assert _close(out, gold["out"])


@pytest.mark.parametrize("tag", ["encoder_small", "encoder_full"])
def test_oracle_encoder_vs_reference_golden(tag):
    """Encoder.forward + DiagonalGaussianDistribution.sample (noise explicit) vs the real reference."""
    m = MANIFEST[tag]
    gold = torch.load(GOLD / f"{tag}.pt")
    spec = ref_encoder.EncoderSpec(ch=m["ch"])
    sd = synth.synth_state_dict(ref_encoder.encoder_param_shapes(spec), seed=m["weight_seed"])
    with torch.no_grad():
        mom = ref_encoder.encoder_forward(sd, spec, gold["x"])
    assert tuple(mom.shape) == (m["B"], 8, m["image_hw"] // 8, m["image_hw"] // 8)
    assert _close(mom, gold["moments"])
    assert _close(ref_encoder.gaussian_sample(mom, gold["noise"]), gold["z"])
    assert torch.equal(ref_encoder.gaussian_mode(gold["moments"]), gold["moments"][:, :4])


def test_oracle_schedule_bit_exact():
    gold = torch.load(GOLD / "schedule.pt")
    for n in (1, 10, 25, 50):
        s = ref_sampling.edm_sigmas(n)
        assert torch.equal(s, gold[f"sigmas_{n}"])
        assert s.shape[0] == n + 1 and s[-1] == 0 and bool((s[:-1][1:] < s[:-1][:-1]).all() if n > 1 else True)
    assert torch.equal(ref_sampling.guider_scale(1.0, 3.5, 18), gold["guider_scale_18"])
    assert torch.equal(ref_sampling.central_guider_scale(1.0, 3.5, 18), gold["central_scale_18"])
    assert torch.equal(ref_sampling.central_guider_scale(1.0, 2.5, 25), gold["central_scale_25"])


def test_oracle_cfg_order_and_chunking():
    """Index paths: [uc; c] ordering, per-frame scale broadcast, decode chunk boundaries."""
    T = 3
    x = torch.arange(T * 2.0).reshape(T, 2)
    c = {"vector": torch.ones(T, 1), "crossattn": torch.ones(T, 1, 2), "concat": torch.ones(T, 2)}
    uc = {k: torch.zeros_like(v) for k, v in c.items()}
    x2, s2, c2 = ref_sampling.guider_prepare_inputs(x, torch.ones(T), c, uc)
    assert torch.equal(x2[:T], x) and torch.equal(x2[T:], x)
    assert c2["vector"][:T].sum() == 0 and c2["vector"][T:].sum() == T  # uc FIRST
    den = torch.cat([torch.zeros(T, 2), torch.ones(T, 2)])
    out = ref_sampling.guider_combine(den, torch.tensor([[1.0, 2.0, 3.0]]), T)
    assert torch.equal(out, torch.tensor([[1.0, 1.0], [2.0, 2.0], [3.0, 3.0]]))
    spec = ref_decoder.DecoderSpec(ch=32)
    sd = synth.synth_state_dict(ref_decoder.decoder_param_shapes(spec), seed=9)
    z = torch.randn(4, 4, 8, 8)
    with torch.no_grad():
        whole = ref_decoder.decode_first_stage(sd, spec, z, 4)
        halves = ref_decoder.decode_first_stage(sd, spec, z, 2)
        ref_halves = torch.cat([ref_decoder.decoder_forward(sd, spec, z[:2] / 0.18215, 2),
                                ref_decoder.decoder_forward(sd, spec, z[2:] / 0.18215, 2)])
    assert torch.allclose(halves, ref_halves, atol=1e-5) and not torch.allclose(whole, halves, atol=1e-3)


def test_clip_tower_oracle_matches_hf_transformers():
    """open_clip (the reference's dependency for the ViT-H/14 image tower) is absent offline; the
    oracle restatement (oracle/ref_clip.py) is pinned against an independent implementation of the same architecture,
    Hugging Face transformers' CLIPVisionModelWithProjection, on per-name seeded weights mapped name by name."""
    transformers = pytest.importorskip("transformers")
    from oracle import ref_clip, synth

    spec = ref_clip.ClipSpec(image_size=56, patch=14, width=128, layers=3, heads=2, mlp=512, embed_dim=96)
    sd = synth.synth_state_dict(ref_clip.clip_visual_param_shapes(spec), seed=31)
    cfg = transformers.CLIPVisionConfig(hidden_size=spec.width, intermediate_size=spec.mlp, projection_dim=spec.embed_dim,
                                        num_hidden_layers=spec.layers, num_attention_heads=spec.heads,
                                        image_size=spec.image_size, patch_size=spec.patch, hidden_act="gelu",
                                        layer_norm_eps=1e-5, attention_dropout=0.0)
    hf = transformers.CLIPVisionModelWithProjection(cfg).eval()
    missing, unexpected = hf.load_state_dict(ref_clip.to_hf_state_dict(sd, spec), strict=False)
    assert not missing and not unexpected
    g = torch.Generator().manual_seed(9)
    img = torch.randn(3, 3, 56, 56, generator=g)
    with torch.no_grad():
        ref = hf(pixel_values=img).image_embeds
        ora = ref_clip.clip_visual_forward(sd, spec, img)
    assert (ref - ora).abs().max() <= 2e-5 * ref.abs().max()


def test_clip_oracle_reproduces_golden_and_preprocess_properties():
    """The committed CLIP fixtures (outputs of the Hugging Face implementation) against the oracle, and closed-form
    properties of the restated kornia resize: a constant image stays constant, the output is 224 x 224, a 224 x 224
    input is only renormalised (no blur, identity interpolation)."""
    from oracle import ref_clip, synth

    m = MANIFEST["clip_small"]
    gold = torch.load(GOLD / "clip_small.pt")
    spec = ref_clip.ClipSpec(**m["spec"])
    sd = synth.synth_state_dict(ref_clip.clip_visual_param_shapes(spec), seed=m["weight_seed"])
    g = torch.Generator().manual_seed(m["x_seed"])
    x = torch.rand(m["B"], 3, m["image_hw"], m["image_hw"], generator=g) * 2.0 - 1.0
    pre = ref_clip.preprocess(x, spec.image_size)
    assert torch.allclose(pre[:, :, ::4, ::4], gold["pre_sub"], atol=1e-6)
    out = ref_clip.clip_visual_forward(sd, spec, pre)
    assert (out - gold["out"]).abs().max() <= 2e-5 * gold["out"].abs().max()
    emb = ref_clip.image_embedder_forward(sd, spec, x, n_cond_frames=1, n_copies=2)
    assert emb.shape == (2 * m["B"], 1, spec.embed_dim) and torch.equal(emb[0], emb[1])
    const = torch.full((1, 3, 300, 300), 0.25)
    pc = ref_clip.preprocess(const)
    assert pc.shape == (1, 3, 224, 224)
    want = (0.625 - torch.tensor(ref_clip.CLIP_MEAN)) / torch.tensor(ref_clip.CLIP_STD)
    assert torch.allclose(pc[0, :, 100, 100], want, atol=1e-5) and torch.allclose(pc[0, :, 0, 0], want, atol=1e-5)
    same = torch.rand(1, 3, 224, 224) * 2 - 1
    assert torch.allclose(ref_clip.preprocess(same),
                          ((same + 1) / 2 - torch.tensor(ref_clip.CLIP_MEAN).view(1, 3, 1, 1)) /
                          torch.tensor(ref_clip.CLIP_STD).view(1, 3, 1, 1), atol=1e-5)
