"""Self-checks that pin the CPU oracle: the reference ships no golden vectors, so these
invariants + the committed golden fixture are what the oracle is anchored to ("parity unpinned")."""
import json
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import svd_oracle as O
from tests.common import TINY, build_models, make_inputs, oracle_cfg, rel_l2, state
from tools.flop_census import census

GOLD = Path(__file__).parent / "golden"


def test_param_counts_match_published_svd():
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    with torch.device("meta"):
        u = UNetSpatioTemporalConditionModel(num_attention_heads=(5, 10, 20, 20), num_frames=14)
        c = ControlNetModel()
    assert sum(p.numel() for p in u.parameters()) == 1_524_623_082
    assert sum(p.numel() for p in c.parameters()) == 680_946_577
    fu, pu = census(72, 128, 2)
    fc, pc = census(72, 128, 2, controlnet=True)
    assert sum(pu.values()) == 1_524_623_082 and sum(pc.values()) == 680_946_577
    assert abs(sum(fu.values()) / 1e12 - 89.854) < 1e-3 and abs(sum(fc.values()) / 1e12 - 32.953) < 1e-3


def test_state_dict_key_scheme():
    unet, cn = build_models(TINY)
    k = set(unet.state_dict().keys())
    for key in ["conv_in.weight", "time_embedding.linear_1.weight", "add_embedding.linear_2.bias",
                "down_blocks.0.resnets.0.spatial_res_block.norm1.weight",
                "down_blocks.0.resnets.1.temporal_res_block.conv1.weight",
                "down_blocks.0.resnets.0.time_mixer.mix_factor",
                "down_blocks.1.attentions.0.transformer_blocks.0.attn2.to_k.weight",
                "down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_out.0.bias",
                "down_blocks.2.attentions.1.temporal_transformer_blocks.0.ff_in.net.0.proj.weight",
                "down_blocks.2.attentions.1.temporal_transformer_blocks.0.ff.net.2.weight",
                "mid_block.attentions.0.time_pos_embed.linear_1.weight", "mid_block.attentions.0.time_mixer.mix_factor",
                "down_blocks.0.downsamplers.0.conv.weight", "up_blocks.0.upsamplers.0.conv.bias",
                "up_blocks.1.resnets.0.spatial_res_block.conv_shortcut.weight", "conv_norm_out.weight",
                "conv_out.weight"]:
        assert key in k, key
    sd = unet.state_dict()
    assert sd["down_blocks.0.resnets.0.temporal_res_block.conv1.weight"].shape == (64, 64, 3, 1, 1)
    assert sd["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k.weight"].shape == (64, 1024)
    assert sd["down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"].shape == (512, 64)
    assert "down_blocks.3.attentions.0.norm.weight" not in k  # DownBlockSpatioTemporal has no attention
    ck = set(cn.state_dict().keys())
    assert "conv_in_concat.weight" in ck and cn.state_dict()["conv_in_concat.weight"].shape == (64, 12, 3, 3)
    assert all(f"controlnet_down_blocks.{i}.weight" in ck for i in range(12)) and "controlnet_mid_block.bias" in ck
    assert not any(x.startswith("up_blocks") for x in ck)


def test_karras_euler_known_answers():
    sig = O.karras_sigmas(25)
    known = [700, 545.729, 421.569, 322.454, 244.023, 182.547, 134.854, 98.2671, 70.5408, 49.8098, 34.5367, 23.4675,
             15.59, 10.0971, 6.35427, 3.8697, 2.26912, 1.27318, 0.678146, 0.339378, 0.157405, 0.0663991, 0.0248026,
             0.0078825, 0.002]
    assert torch.allclose(sig[:-1], torch.tensor(known), rtol=2e-5)
    assert float(sig[-1]) == 0.0
    ts = O.euler_timesteps(sig)
    assert abs(float(ts[0]) - 1.63777) < 1e-4 and abs(float(ts[-1]) + 1.55365) < 1e-4
    assert abs(O.init_noise_sigma(sig) - 700.000732) < 1e-3
    # Euler v-pred step == training parametrisation c_skip/c_out (train_code/train_csvd.py:902-904)
    x, eps = torch.randn(3, 5), torch.randn(3, 5)
    s, sn = 10.0, 6.0
    x0 = eps * (-s / (s * s + 1) ** 0.5) + x / (s * s + 1)
    assert torch.allclose(O.euler_step(eps, x, s, sn), x + (x - x0) / s * (sn - s))
    # product scheduler agrees with the oracle's table
    from svd.scheduler import EulerDiscreteScheduler
    sch = EulerDiscreteScheduler()
    sch.set_timesteps(25)
    assert torch.allclose(sch.sigmas, sig, rtol=1e-6) and torch.allclose(sch.timesteps, ts, rtol=1e-5, atol=1e-6)
    assert abs(sch.init_noise_sigma - O.init_noise_sigma(sig)) < 1e-4


def test_primitives_against_torch_modules():
    """The oracle's primitives vs torch nn.Modules fed the same parameters (the calls diffusers makes)."""
    torch.manual_seed(0)
    lin1, lin2 = torch.nn.Linear(32, 64), torch.nn.Linear(64, 16)
    sd = {"e.linear_1.weight": lin1.weight, "e.linear_1.bias": lin1.bias, "e.linear_2.weight": lin2.weight,
          "e.linear_2.bias": lin2.bias}
    x = torch.randn(5, 32)
    assert torch.allclose(O.timestep_embedding(sd, "e", x), lin2(F.silu(lin1(x))))
    t = torch.tensor([0.0, 1.5, 200.0])
    emb = O.timesteps_sinusoid(t, 8)
    fr = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(4) / 4)
    assert torch.allclose(emb[:, :4], torch.cos(t[:, None] * fr)) and torch.allclose(emb[:, 4:], torch.sin(t[:, None] * fr))
    # GEGLU feed-forward uses erf GELU on the SECOND half
    p1, p2 = torch.nn.Linear(8, 64), torch.nn.Linear(32, 8)
    sd = {"f.net.0.proj.weight": p1.weight, "f.net.0.proj.bias": p1.bias, "f.net.2.weight": p2.weight,
          "f.net.2.bias": p2.bias}
    x = torch.randn(3, 7, 8)
    hcat = p1(x)
    assert torch.allclose(O.feed_forward(sd, "f", x), p2(hcat[..., :32] * F.gelu(hcat[..., 32:])), atol=1e-6)
    # attention == explicit softmax(QK^T/sqrt(d))V
    heads, C = 2, 128
    ws = {n: torch.randn(C, C) * C ** -0.5 for n in ["q", "k", "v", "o"]}
    sd = {"a.to_q.weight": ws["q"], "a.to_k.weight": ws["k"], "a.to_v.weight": ws["v"], "a.to_out.0.weight": ws["o"],
          "a.to_out.0.bias": torch.randn(C)}
    x = torch.randn(2, 9, C)
    q, k, v = [(x @ ws[n].t()).view(2, 9, heads, 64).transpose(1, 2) for n in "qkv"]
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, -1) @ v
    ref = att.transpose(1, 2).reshape(2, 9, C) @ ws["o"].t() + sd["a.to_out.0.bias"]
    assert torch.allclose(O.attention(sd, "a", x, None, heads), ref, atol=1e-4)


def test_alpha_blender_matches_formula():
    sd = {"m.mix_factor": torch.tensor([0.3])}
    ind = torch.zeros(2, 14)
    a, b = torch.randn(28, 5, 16), torch.randn(28, 5, 16)
    al = torch.sigmoid(torch.tensor(0.3))
    assert torch.allclose(O.alpha_blend(sd, "m", a, b, ind), al * a + (1 - al) * b)
    a5, b5 = torch.randn(2, 8, 14, 3, 3), torch.randn(2, 8, 14, 3, 3)
    assert torch.allclose(O.alpha_blend(sd, "m", a5, b5, ind), al * a5 + (1 - al) * b5)


def test_zero_init_gesturenet_is_exactly_vl():
    """from_unet leaves conv_in_concat and the 13 zero convs at zero => residuals are exactly 0 (§8c item 3)."""
    from svd.temporal_controlnet import ControlNetModel
    unet, _ = build_models(TINY, controlnet=False)
    cn = ControlNetModel(**TINY)
    cfg = oracle_cfg(TINY)
    sample, ehs, ati, cond = make_inputs(2, 14, 8, 16)
    with torch.no_grad():
        d, m = O.controlnet_forward(state(cn), cfg, sample, torch.tensor(0.7), ehs, ati, torch.cat([cond, cond]))
        assert all(float(x.abs().max()) == 0.0 for x in d) and float(m.abs().max()) == 0.0
        y0 = O.unet_forward(state(unet), cfg, sample, torch.tensor(0.7), ehs, ati)
        y1 = O.unet_forward(state(unet), cfg, sample, torch.tensor(0.7), ehs, ati, d, m)
    assert torch.equal(y0, y1)


def test_from_unet_copies_encoder_and_keeps_zero_convs():
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    with torch.device("meta"):
        ControlNetModel()  # defaults construct
    unet = UNetSpatioTemporalConditionModel(num_frames=14, **TINY)
    # from_unet builds from class defaults (SVD size) like the reference; exercise the copy on a same-config model
    cn = ControlNetModel(**TINY)
    cn.down_blocks.load_state_dict(unet.down_blocks.state_dict())
    cn.mid_block.load_state_dict(unet.mid_block.state_dict())
    a = unet.state_dict()["down_blocks.1.resnets.0.spatial_res_block.conv1.weight"]
    assert torch.equal(a, cn.state_dict()["down_blocks.1.resnets.0.spatial_res_block.conv1.weight"])
    assert float(cn.conv_in_concat.weight.abs().max()) == 0.0
    assert all(float(m.weight.abs().max()) == 0.0 for m in cn.controlnet_down_blocks)


def test_time_context_quirk_row_mod_B():
    """§8c item 4: the reference's time_context flatten makes temporal row r = b*S + s read context r mod B."""
    B, Fr, S, L, D = 2, 3, 4, 5, 8
    ehs = torch.arange(B).float()[:, None, None].expand(B, L, D)  # context b is filled with the value b
    ehs_bf = ehs.repeat_interleave(Fr, dim=0)
    first = ehs_bf[None, :].reshape(B, Fr, L, D)[:, 0]
    tc = first[None, :].broadcast_to(S, B, L, D).reshape(S * B, L, D)
    got = tc[:, 0, 0].long().tolist()  # context id seen by temporal rows 0..B*S-1 (row = b*S + s)
    assert got == [r % B for r in range(B * S)] == [0, 1, 0, 1, 0, 1, 0, 1]
    # B = 1 degenerates to the correct mapping
    first1 = ehs_bf[:Fr][None, :].reshape(1, Fr, L, D)[:, 0]
    assert first1[None].broadcast_to(S, 1, L, D).reshape(S, L, D)[:, 0, 0].tolist() == [0.0] * S


def test_quirk_changes_output_vs_correct_indexing():
    """The oracle's transformer really uses the quirky mapping: permuting which context is 'first' changes even rows."""
    unet, _ = build_models(TINY, controlnet=False)
    sd = state(unet)
    B, Fr, h, w = 2, 14, 4, 4
    x = torch.randn(B * Fr, 64, h, w)
    ehs = torch.randn(B, 6, 1024)
    ind = torch.zeros(B, Fr)
    p = "down_blocks.0.attentions.0"
    with torch.no_grad():
        y = O.transformer_spatio_temporal(sd, p, x, ehs.repeat_interleave(Fr, 0), ind, 1)
        # make both contexts identical to context 1 only for the TEMPORAL path by checking sensitivity:
        ehs2 = ehs.clone()
        ehs2[0] = ehs[1]
        y2 = O.transformer_spatio_temporal(sd, p, x, ehs2.repeat_interleave(Fr, 0), ind, 1)
    # batch element 1 never reads context 0 in the spatial block, but its EVEN pixels do in the temporal block
    d = (y - y2)[Fr:].abs().amax(dim=(0, 1)).reshape(-1)  # per pixel s of batch element 1
    rows = torch.arange(h * w) + 1 * h * w
    even = (rows % 2 == 0)
    assert float(d[even].min()) > 0 and float(d[~even].max()) == 0.0


def test_single_token_context_closed_form():
    """§8c item 5: with L = 1 the cross-attention output is to_out(to_v(ctx)), independent of the query."""
    unet, _ = build_models(TINY, controlnet=False)
    sd = state(unet)
    p = "down_blocks.0.attentions.0.transformer_blocks.0.attn2"
    ctx = torch.randn(3, 1, 1024)
    x = torch.randn(3, 10, 64)
    with torch.no_grad():
        y = O.attention(sd, p, x, ctx, 1)
        ref = O.linear(sd, p + ".to_out.0", O.linear(sd, p + ".to_v", ctx)).expand(3, 10, 64)
    assert torch.allclose(y, ref, atol=1e-5)


def test_temporal_layers_reduce_over_frames_only_within_a_video():
    """5-D GroupNorm statistics span all frames of ONE video: changing video 1 must not change video 0 (§8e)."""
    unet, _ = build_models(TINY, controlnet=False)
    sd, cfg = state(unet), oracle_cfg(TINY)
    sample, ehs, ati, _ = make_inputs(2, 14, 8, 8)
    ehs[0] = ehs[1]  # identical contexts so the quirk does not couple the two sequences
    s2 = sample.clone()
    s2[1] += 1.0
    with torch.no_grad():
        a = O.unet_forward(sd, cfg, sample, torch.tensor(0.3), ehs, ati)
        b = O.unet_forward(sd, cfg, s2, torch.tensor(0.3), ehs, ati)
    assert torch.equal(a[0], b[0]) and not torch.equal(a[1], b[1])


def test_golden_fixture():
    """Committed oracle outputs (tests/golden/make_golden.py) — detects any drift of the oracle itself."""
    meta = json.loads((GOLD / "tiny_vgl.json").read_text())
    gold = torch.load(GOLD / "tiny_vgl.pt")
    from tests.golden.make_golden import compute
    out = compute()
    for k in meta["tensors"]:
        assert rel_l2(out[k], gold[k]) < 1e-5, k
