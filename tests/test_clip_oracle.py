"""Pins oracle/clip_oracle.py: against golden outputs of the `transformers` CLIP modules the reference calls and of the
reference's own `_resize_with_antialiasing` source (tests/golden/clip_golden.pt, made by make_clip_golden.py), against
the live library when it is importable, and the closed forms of the assembly tail of encode_clip."""
from pathlib import Path

import pytest
import torch

from oracle import clip_oracle as CO
from tests.common import (TINY_CLIP_TEXT, TINY_CLIP_VISION, TINY_CLIP_VISION_D80, clip_inputs, clip_text_sd,
                          clip_vision_sd, rel_l2)

GOLD = torch.load(Path(__file__).parent / "golden" / "clip_golden.pt")


@pytest.mark.parametrize("name,cfg", [("vision", TINY_CLIP_VISION), ("vision_d80", TINY_CLIP_VISION_D80)])
def test_vision_tower_vs_transformers_golden(name, cfg):
    px, _ = clip_inputs(cfg, TINY_CLIP_TEXT, n=2)
    out = CO.vision_image_embeds(clip_vision_sd(cfg), px, cfg["num_attention_heads"], cfg["hidden_act"])
    assert out.shape == GOLD[name].shape and rel_l2(out, GOLD[name]) < 2e-5


def test_text_tower_vs_transformers_golden():
    _, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT, n=2)
    out = CO.text_last_hidden_state(clip_text_sd(TINY_CLIP_TEXT), ids, TINY_CLIP_TEXT["num_attention_heads"])
    assert out.shape == GOLD["text"].shape and rel_l2(out, GOLD["text"]) < 2e-5


def test_towers_vs_live_transformers():
    tr = pytest.importorskip("transformers")
    cfg = TINY_CLIP_VISION_D80
    m = tr.CLIPVisionModelWithProjection(tr.CLIPVisionConfig(**cfg)).eval()
    sd = clip_vision_sd(cfg, seed=5)
    m.load_state_dict(sd, strict=True)
    px, ids = clip_inputs(cfg, TINY_CLIP_TEXT, n=1, seed=4)
    with torch.no_grad():
        assert rel_l2(CO.vision_image_embeds(sd, px, cfg["num_attention_heads"], "quick_gelu"), m(px).image_embeds) < 2e-5
    t = tr.CLIPTextModel(tr.CLIPTextConfig(**TINY_CLIP_TEXT)).eval()
    tsd = clip_text_sd(TINY_CLIP_TEXT, seed=6)
    t.load_state_dict(tsd, strict=False)
    with torch.no_grad():
        assert rel_l2(CO.text_last_hidden_state(tsd, ids, 2), t(ids)[0]) < 2e-5


def test_causal_mask_of_the_text_tower():
    """Token i must not depend on tokens > i."""
    sd = clip_text_sd(TINY_CLIP_TEXT)
    _, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT, n=1)
    ids2 = ids.clone()
    ids2[0, 40:] = (ids2[0, 40:] + 1) % TINY_CLIP_TEXT["vocab_size"]
    a, b = CO.text_last_hidden_state(sd, ids, 2), CO.text_last_hidden_state(sd, ids2, 2)
    assert torch.allclose(a[0, :40], b[0, :40], atol=1e-6) and not torch.allclose(a[0, 40:], b[0, 40:], atol=1e-3)


def test_resize_vs_reference_source_golden():
    g = torch.Generator().manual_seed(3)
    img = torch.rand(1, 3, 96, 160, generator=g) * 2 - 1
    assert torch.allclose(CO.resize_with_antialiasing(img, (56, 56)), GOLD["resize_96x160_to_56"], atol=1e-5)
    assert torch.allclose(CO.resize_with_antialiasing(img[..., :40, :40], (56, 56)), GOLD["resize_40x40_to_56"], atol=1e-5)


def test_assembly_tail_closed_forms():
    g = torch.Generator().manual_seed(0)
    emb, txt = torch.randn(1, 1024, generator=g), torch.randn(1, 77, 1024, generator=g)
    ehs = CO.assemble(emb, txt, do_cfg=True)
    assert ehs.shape == (2, 78, 1024) and float(ehs[0].abs().max()) == 0.0      # zeros FIRST (uncond half)
    slab = ehs[1]
    assert abs(float(slab.mean())) < 1e-6 and abs(float(slab.var(unbiased=False)) - 1.0) < 1e-4  # joint (78, 1024) norm
    ln = torch.nn.LayerNorm((78, 1024))
    assert torch.allclose(slab, ln(torch.cat([txt, emb[:, None]], 1))[0], atol=1e-6)
    assert torch.equal(slab[77] * 0 + 1, torch.ones(1024))                       # image token is the LAST row
    no_text = CO.assemble(emb, None, do_cfg=False)
    assert no_text.shape == (1, 1, 1024) and torch.equal(no_text[0, 0], emb[0])  # use_text=False: no LayerNorm at all
