"""End-to-end host logic of the drop-in on CPU: StableVideoDiffusionControlNetPipeline.__call__ with PIL / token-id /
numpy inputs, running CLIP towers -> conditioning assembly -> VAE encode -> 2 Euler steps of UNet + GestureNet -> chunked
VAE decode through tests/fake_lib.py, against the same computation composed from the three oracles. Plus the CLIP
engine's schedule on its own (head_dim 80 padding, causal text tower) and the towers' module surface."""
import pytest
import torch

from oracle import clip_oracle as CO
from tests import fake_lib, pipeline_case as PC
from tests.common import (TINY_CLIP_TEXT, TINY_CLIP_VISION, TINY_CLIP_VISION_D80, clip_inputs, clip_text_sd,
                          clip_vision_sd, rel_l2)
from this_and_that_vdm_b200.clip_engine import ClipTowerEngine, assemble_conditioning

CAP = 3e-2


@pytest.fixture
def cpu_engines(monkeypatch):
    """Let the drop-in modules build their engines on CPU tensors (the emulation is installed by the test)."""
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from svd.clip_towers import _TowerBase
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from this_and_that_vdm_b200.engine import DenoiserEngine
    from this_and_that_vdm_b200.vae_engine import VaeEngine

    def cached(make):
        def get(self):
            if getattr(self, "_engine", None) is None:
                self._engine = make(self)
            return self._engine
        return get

    monkeypatch.setattr(UNetSpatioTemporalConditionModel, "_get_engine", cached(lambda m: DenoiserEngine(m, "unet")))
    monkeypatch.setattr(ControlNetModel, "_get_engine", cached(lambda m: DenoiserEngine(m, "controlnet")))
    monkeypatch.setattr(AutoencoderKLTemporalDecoder, "_get_engine", cached(lambda m: VaeEngine(m)))
    monkeypatch.setattr(_TowerBase, "_get_engine",
                        cached(lambda m: ClipTowerEngine(m.state_dict(), m._cfg, m._kind, "cpu")))


@pytest.mark.parametrize("cfg", [TINY_CLIP_VISION, TINY_CLIP_VISION_D80], ids=["d64", "d80"])
def test_vision_tower_schedule_vs_oracle_and_transformers_golden(cfg):
    sd = clip_vision_sd(cfg)
    px, _ = clip_inputs(cfg, TINY_CLIP_TEXT, n=2)
    with torch.no_grad(), fake_lib.installed():
        out = ClipTowerEngine(sd, cfg, "vision", "cpu").image_embeds(px)
    ref = CO.vision_image_embeds(sd, px, cfg["num_attention_heads"], cfg["hidden_act"])
    assert out.shape == ref.shape and out.dtype == torch.float32 and rel_l2(out, ref) < CAP
    gold = torch.load(PC.__file__.replace("pipeline_case.py", "golden/clip_golden.pt"))
    assert rel_l2(out, gold["vision" if cfg is TINY_CLIP_VISION else "vision_d80"]) < CAP


@pytest.mark.parametrize("cfg,kind", [(TINY_CLIP_VISION_D80, "vision"), (TINY_CLIP_TEXT, "text")], ids=["vision", "text"])
def test_batched_heads_equal_the_per_head_schedule(cfg, kind):
    """Block-diagonal batching of the heads (one scores GEMM / softmax / P V GEMM per group of heads) must reproduce
    the 5-launches-per-head schedule; the zero blocks contribute exact zeros."""
    sd = clip_vision_sd(cfg) if kind == "vision" else clip_text_sd(cfg)
    px, ids = clip_inputs(TINY_CLIP_VISION_D80, TINY_CLIP_TEXT, n=2)
    with torch.no_grad(), fake_lib.installed():
        eng = ClipTowerEngine(sd, cfg, kind, "cpu")
        run = (lambda: eng.image_embeds(px)) if kind == "vision" else (lambda: eng.last_hidden_state(ids))
        n0 = fake_lib.launch_count()
        a = run()
        n_batched = fake_lib.launch_count() - n0
        eng.batch_heads = False
        b = run()
        n_per_head = fake_lib.launch_count() - n0 - n_batched
    assert rel_l2(a, b) < 2e-3 and n_batched < n_per_head


def test_text_tower_schedule_is_causal_and_matches_oracle():
    sd = clip_text_sd(TINY_CLIP_TEXT)
    _, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT, n=2)
    ids2 = ids.clone()
    ids2[:, 50:] = (ids2[:, 50:] + 7) % TINY_CLIP_TEXT["vocab_size"]
    with torch.no_grad(), fake_lib.installed():
        eng = ClipTowerEngine(sd, TINY_CLIP_TEXT, "text", "cpu")
        out, out2 = eng.last_hidden_state(ids), eng.last_hidden_state(ids2)
    assert rel_l2(out, CO.text_last_hidden_state(sd, ids, 2)) < CAP
    assert torch.equal(out[:, :50], out2[:, :50]) and not torch.equal(out[:, 50:], out2[:, 50:])


def test_assembly_and_engine_errors():
    g = torch.Generator().manual_seed(0)
    emb, txt = torch.randn(1, 1024, generator=g), torch.randn(1, 77, 1024, generator=g)
    with fake_lib.installed():
        a = assemble_conditioning(emb, txt, True)
        b = assemble_conditioning(emb, None, False)
    assert rel_l2(a, CO.assemble(emb, txt, True)) < 1e-6 and float(a[0].abs().max()) == 0.0
    assert torch.equal(b, emb[:, None])
    from this_and_that_vdm_b200 import lib
    with fake_lib.installed():
        with pytest.raises(lib.TtvdmError, match="hidden_act"):
            ClipTowerEngine({}, dict(TINY_CLIP_TEXT, hidden_act="relu"), "text", "cpu")
        with pytest.raises(lib.TtvdmError, match="64"):
            ClipTowerEngine({}, dict(TINY_CLIP_TEXT, hidden_size=96), "text", "cpu")


def test_tower_modules_surface(tmp_path, cpu_engines):
    """from_pretrained / save_pretrained on the HF directory layout, HF key names, transformers-style outputs, and no
    CPU execution outside the emulation."""
    from svd.clip_towers import CLIPTextModel, CLIPVisionModelWithProjection
    vis = CLIPVisionModelWithProjection(TINY_CLIP_VISION)
    vis.load_state_dict(clip_vision_sd(TINY_CLIP_VISION))
    vis.save_pretrained(tmp_path / "image_encoder")
    vis2 = CLIPVisionModelWithProjection.from_pretrained(str(tmp_path), subfolder="image_encoder")
    assert set(vis2.state_dict()) == set(clip_vision_sd(TINY_CLIP_VISION))
    assert all(torch.equal(a, b) for a, b in zip(vis.state_dict().values(), vis2.state_dict().values()))
    assert vis2.config.projection_dim == 64 and vis2.dtype == torch.float32
    px, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT)
    txt = CLIPTextModel(TINY_CLIP_TEXT)
    txt.load_state_dict(clip_text_sd(TINY_CLIP_TEXT))
    with fake_lib.installed():
        out = vis2(px)
        hs = txt(ids)
    assert out.image_embeds.shape == (1, 64) and hs[0].shape == (1, 77, 128) and hs.last_hidden_state is hs[0]
    tr = pytest.importorskip("transformers")
    hf = tr.CLIPVisionModelWithProjection(tr.CLIPVisionConfig(**TINY_CLIP_VISION)).eval()
    wrapped = CLIPVisionModelWithProjection.from_hf(hf)
    with torch.no_grad(), fake_lib.installed():
        assert rel_l2(wrapped(px).image_embeds, hf(px).image_embeds) < CAP


def test_towers_refuse_cpu_without_emulation():
    from svd.clip_towers import CLIPVisionModelWithProjection
    vis = CLIPVisionModelWithProjection(TINY_CLIP_VISION)
    px, _ = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT)
    with pytest.raises(RuntimeError, match="CUDA"):
        vis(px)


def test_vgl_pipeline_end_to_end_vs_oracles(cpu_engines):
    mods, sds = PC.build("cpu")
    with torch.no_grad(), fake_lib.installed():
        n0 = fake_lib.launch_count()
        frames = PC.run_pipeline(mods, "cpu", output_type="pt")
        launches = fake_lib.launch_count() - n0
        pil = PC.run_pipeline(mods, "cpu", output_type="pil")
        ref, _ = PC.run_oracle(sds)
    assert len(frames) == 1 and frames[0].shape == (PC.FRAMES, 3, PC.H, PC.W) and launches > 3000
    want = (ref[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert rel_l2(frames[0], want) < 5e-2  # CLIP + VAE encode + 2 denoising steps + VAE decode chained in bf16 storage
    assert len(pil) == 1 and len(pil[0]) == PC.FRAMES and pil[0][0].size == (PC.W, PC.H)


def test_vl_pipeline_end_to_end_without_text_vs_oracles(cpu_engines):
    """StableVideoDiffusionPipeline (UNet only) with use_text=False: one context token, no LayerNorm, no GestureNet."""
    mods, sds = PC.build("cpu")
    with torch.no_grad(), fake_lib.installed():
        frames = PC.run_vl_pipeline(mods, "cpu", output_type="pt")
        ref = PC.run_vl_oracle(sds)
    want = (ref[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert frames[0].shape == want.shape and rel_l2(frames[0], want) < 5e-2


def test_vgl_pipeline_with_fp16_modules(cpu_engines):
    """The reference casts every module to fp16 (weight_dtype, test_code/inference.py:357-368); the drop-ins keep the
    caller's dtype at their boundaries and compute in bf16 / fp32 inside."""
    mods, _ = PC.build("cpu")
    for m in mods.values():
        m.half()
    sds = {k: {n: v.detach().float() for n, v in m.state_dict().items()} for k, m in mods.items()}
    with torch.no_grad(), fake_lib.installed():
        frames = PC.run_pipeline(mods, "cpu", output_type="pt")
        ref, _ = PC.run_oracle(sds, latent_dtype=torch.float16)
    want = (ref[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert frames[0].dtype == torch.float32 and rel_l2(frames[0], want) < 5e-2


def test_preprocess_image_tensor_inputs_follow_vae_image_processor():
    """tensor images are a documented input of __call__. diffusers' VaeImageProcessor.preprocess —
    which the reference calls at svd/pipeline_stable_video_diffusion_controlnet.py:541 — stacks / concatenates them,
    resizes to (height, width), maps [0,1] -> [-1,1] unless the tensor already has negative values, and passes 4-channel
    latents through untouched."""
    import torch
    from svd.pipeline_common import SVDPipelineBase as P
    pre = lambda im, h, w: P._preprocess_image(None, im, h, w)  # noqa: E731
    g = torch.Generator().manual_seed(0)
    x01 = torch.rand(1, 3, 16, 24, generator=g)
    out = pre(x01, 32, 48)
    assert out.shape == (1, 3, 32, 48)
    assert torch.allclose(out, 2.0 * torch.nn.functional.interpolate(x01, size=(32, 48)) - 1.0)
    assert float(out.min()) < 0 and float(out.max()) <= 1.0
    xs = x01 * 2 - 1                        # already in [-1, 1]: not normalised again
    assert torch.equal(pre(xs, 16, 24), xs)
    lst = [torch.rand(3, 16, 24, generator=g), torch.rand(3, 16, 24, generator=g)]   # list of 3-D tensors: stacked
    out = pre(lst, 16, 24)
    assert out.shape == (2, 3, 16, 24) and torch.allclose(out, 2 * torch.stack(lst) - 1)
    lat = torch.randn(2, 4, 8, 8, generator=g)   # latents: returned as they are
    assert torch.equal(pre(lat, 64, 64), lat)
