"""Conditioning builder on the B200 (scope-table next #2): the sm_100a CLIP towers behind svd.clip_towers against the CPU
fp32 oracle (oracle/clip_oracle.py, pinned against `transformers`) and the golden outputs of the library itself, through
the module API the reference's encode_clip calls; plus the two conditioning-only kernels and the causal softmax through
the raw C ABI. Tolerance: rel-L2 vs the fp32 oracle <= torch-eager-bf16 error + 1e-3, cap 3e-2."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import clip_oracle as CO
from tests.common import (TINY_CLIP_TEXT, TINY_CLIP_VISION, TINY_CLIP_VISION_D80, clip_inputs, clip_text_sd,
                          clip_vision_sd, rel_l2)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "clip_golden.pt"
CAP = 3e-2


def _bf(sd):
    return {k: v.to("cuda", torch.bfloat16) for k, v in sd.items()}


def test_conditioning_kernels_vs_torch():
    from this_and_that_vdm_b200 import lib
    lib.init()
    g = torch.Generator().manual_seed(0)
    for kind, fn in ((lib.ACT_GELU, F.gelu), (lib.ACT_QUICK_GELU, lambda v: v * torch.sigmoid(1.702 * v))):
        x = (torch.randn(257, 640, generator=g) * 2).to(torch.bfloat16).cuda()
        want = fn(x.float())
        lib.act_inplace(x, kind)
        assert torch.allclose(x.float(), want, atol=2e-2, rtol=8e-3)
    x = (torch.randn(3, 78 * 1024, generator=g) * 3 + 1).cuda()
    out = torch.empty_like(x)
    lib.layernorm_flat(x, out, rows=3, n=78 * 1024, eps=1e-5)
    assert torch.allclose(out, F.layer_norm(x, (78 * 1024,), None, None, 1e-5), atol=2e-5)
    S, Sp = 77, 128
    sc = (torch.randn(S, S, generator=g) * 3).cuda()
    pr = torch.empty(S, Sp, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(sc, pr, rows=S, cols=S, ldx=S, ldo=Sp, cols_out=Sp, causal=True)
    want = torch.softmax(sc.masked_fill(torch.ones(S, S, dtype=torch.bool, device="cuda").triu(1), float("-inf")), -1)
    assert torch.allclose(pr[:, :S].float(), want, atol=4e-3, rtol=1e-2) and float(pr[:, S:].abs().max()) == 0.0
    assert float(pr[0, 0]) == 1.0 and float(pr[5, 6:].abs().max()) == 0.0
    # blocks of queries (one block per head): row r is query r % period
    sc2 = torch.cat([sc, sc * 0.5])
    pr2 = torch.empty(2 * S, Sp, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(sc2, pr2, rows=2 * S, cols=S, ldx=S, ldo=Sp, cols_out=Sp, causal=S)
    assert torch.equal(pr2[:S], pr) and float(pr2[S, 0]) == 1.0 and float(pr2[S + 5, 6:].abs().max()) == 0.0


@pytest.mark.parametrize("name,cfg", [("vision", TINY_CLIP_VISION), ("vision_d80", TINY_CLIP_VISION_D80)])
def test_vision_tower_vs_oracle_eager_and_transformers_golden(name, cfg):
    from svd.clip_towers import CLIPVisionModelWithProjection
    sd = clip_vision_sd(cfg)
    m = CLIPVisionModelWithProjection(cfg)
    m.load_state_dict(sd)
    m.to("cuda")
    px, _ = clip_inputs(cfg, TINY_CLIP_TEXT, n=2)
    heads, act = cfg["num_attention_heads"], cfg["hidden_act"]
    with torch.no_grad():
        ref = CO.vision_image_embeds(sd, px, heads, act)
        eager = CO.vision_image_embeds(_bf(sd), px.to("cuda", torch.bfloat16), heads, act)
        out = m(px.cuda()).image_embeds
    e, ee = rel_l2(out, ref), rel_l2(eager, ref)
    assert out.shape == ref.shape and e <= ee + 1e-3 and e < CAP, (e, ee)
    assert rel_l2(out, torch.load(GOLD)[name]) < CAP


def test_text_tower_vs_oracle_eager_and_transformers_golden():
    from svd.clip_towers import CLIPTextModel
    sd = clip_text_sd(TINY_CLIP_TEXT)
    m = CLIPTextModel(TINY_CLIP_TEXT)
    m.load_state_dict(sd)
    m.to("cuda")
    _, ids = clip_inputs(TINY_CLIP_VISION, TINY_CLIP_TEXT, n=2)
    with torch.no_grad():
        ref = CO.text_last_hidden_state(sd, ids, 2)
        eager = CO.text_last_hidden_state(_bf(sd), ids.cuda(), 2)
        out = m(ids.cuda())[0]
    e, ee = rel_l2(out, ref), rel_l2(eager, ref)
    assert out.shape == (2, 77, 128) and e <= ee + 1e-3 and e < CAP, (e, ee)
    assert rel_l2(out, torch.load(GOLD)["text"]) < CAP
    ids2 = ids.clone()
    ids2[:, 50:] = (ids2[:, 50:] + 7) % TINY_CLIP_TEXT["vocab_size"]
    with torch.no_grad():
        out2 = m(ids2.cuda())[0]
    assert torch.equal(out[:, :50], out2[:, :50])  # causal: a token never sees later tokens


def test_vit_h_shaped_layers_and_assembly():
    """ViT-H/14 geometry (1280 wide, 16 heads of 80, MLP 5120, 257 tokens, projection 1024) with 2 layers, then the tail
    of encode_clip against the oracle."""
    from svd.clip_towers import CLIPVisionModelWithProjection
    from this_and_that_vdm_b200.clip_engine import assemble_conditioning
    cfg = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=2, num_attention_heads=16, image_size=224,
               patch_size=14, projection_dim=1024, hidden_act="gelu")
    sd = clip_vision_sd(cfg)
    m = CLIPVisionModelWithProjection(cfg)
    m.load_state_dict(sd)
    m.to("cuda")
    g = torch.Generator().manual_seed(1)
    px = torch.randn(1, 3, 224, 224, generator=g)
    txt = torch.randn(1, 77, 1024, generator=g)
    with torch.no_grad():
        ref = CO.vision_image_embeds(sd, px, 16, "gelu")
        out = m(px.cuda()).image_embeds
        assert rel_l2(out, ref) < CAP
        ehs = assemble_conditioning(out, txt.cuda(), True)
    want = CO.assemble(ref, txt, True)
    assert ehs.shape == (2, 78, 1024) and float(ehs[0].abs().max()) == 0.0 and rel_l2(ehs, want) < CAP


def test_cuda_graph_replay_equals_kernel_by_kernel():
    """The drop-in towers replay one CUDA graph per input shape; the result must be bit-identical to launching the same
    kernels one by one, also for a second input of the same shape (static input buffer is refreshed)."""
    from this_and_that_vdm_b200.clip_engine import ClipTowerEngine
    cfg = TINY_CLIP_VISION_D80
    eng = ClipTowerEngine(clip_vision_sd(cfg), cfg, "vision", "cuda:0")
    px, ids = clip_inputs(cfg, TINY_CLIP_TEXT, n=2)
    px2 = px.flip(0) * 0.5
    teng = ClipTowerEngine(clip_text_sd(TINY_CLIP_TEXT), TINY_CLIP_TEXT, "text", "cuda:0")
    with torch.no_grad():
        for x in (px, px2, px):
            assert torch.equal(eng.image_embeds(x.cuda(), use_graph=True), eng.image_embeds(x.cuda()))
        ids2 = (ids + 3) % TINY_CLIP_TEXT["vocab_size"]
        for i in (ids, ids2):
            assert torch.equal(teng.last_hidden_state(i.cuda(), use_graph=True), teng.last_hidden_state(i.cuda()))
    assert len(eng._graphs) == 1 and len(teng._graphs) == 1


@pytest.mark.parametrize("kind", ["vision", "text"])
def test_batched_heads_equal_the_per_head_schedule_on_gpu(kind):
    from this_and_that_vdm_b200.clip_engine import ClipTowerEngine
    cfg = TINY_CLIP_VISION_D80 if kind == "vision" else TINY_CLIP_TEXT
    sd = clip_vision_sd(cfg) if kind == "vision" else clip_text_sd(cfg)
    px, ids = clip_inputs(TINY_CLIP_VISION_D80, TINY_CLIP_TEXT, n=2)
    eng = ClipTowerEngine(sd, cfg, kind, "cuda:0")
    run = (lambda: eng.image_embeds(px.cuda())) if kind == "vision" else (lambda: eng.last_hidden_state(ids.cuda()))
    with torch.no_grad():
        a = run()
        eng.batch_heads = False
        b = run()
    assert rel_l2(a, b) < 2e-3
