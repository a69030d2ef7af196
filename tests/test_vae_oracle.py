"""Self-checks that pin oracle/vae_oracle.py (PARITY UNPINNED: diffusers is absent and the reference has no VAE tests):
parameter counts and key scheme of the drop-in module, torch-module cross-checks of every primitive, closed forms of
the temporal paths, and the committed golden vectors."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import vae_oracle as VO
from tests.common import SVD_VAE, TINY_VAE, build_vae, rel_l2, state, vae_inputs

GOLD = Path(__file__).parent / "golden"


def test_parameter_counts_and_key_scheme():
    """The encoder half is the Stable Diffusion VAE encoder (published size 34,163,592 parameters) + the 72-parameter
    quant_conv; the temporal decoder adds the Conv3d blocks and the 1 mix factor per SpatioTemporalResBlock."""
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    with torch.device("meta"):
        vae = AutoencoderKLTemporalDecoder(**SVD_VAE)
    n = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert n(vae.encoder) == 34_163_592 and n(vae.quant_conv) == 72
    assert n(vae.decoder) == 63_579_183 and n(vae) == 97_742_847
    keys = set(vae.state_dict().keys())
    for k in ["encoder.conv_in.weight", "encoder.down_blocks.0.resnets.1.conv2.bias",
              "encoder.down_blocks.1.resnets.0.conv_shortcut.weight", "encoder.down_blocks.2.downsamplers.0.conv.weight",
              "encoder.mid_block.attentions.0.group_norm.weight", "encoder.mid_block.attentions.0.to_out.0.bias",
              "encoder.conv_norm_out.weight", "encoder.conv_out.bias", "quant_conv.weight", "decoder.conv_in.weight",
              "decoder.mid_block.resnets.1.temporal_res_block.conv2.weight",
              "decoder.mid_block.resnets.0.time_mixer.mix_factor", "decoder.mid_block.attentions.0.to_q.bias",
              "decoder.up_blocks.0.resnets.2.spatial_res_block.norm2.weight",
              "decoder.up_blocks.2.resnets.0.spatial_res_block.conv_shortcut.weight",
              "decoder.up_blocks.2.upsamplers.0.conv.weight", "decoder.conv_norm_out.bias", "decoder.conv_out.weight",
              "decoder.time_conv_out.weight"]:
        assert k in keys, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in keys
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys
    assert not any("time_emb_proj" in k for k in keys) and "post_quant_conv.weight" not in keys
    sd = vae.state_dict()
    assert tuple(sd["decoder.time_conv_out.weight"].shape) == (3, 3, 3, 1, 1)
    assert tuple(sd["encoder.conv_out.weight"].shape) == (8, 512, 3, 3)
    assert tuple(sd["decoder.mid_block.resnets.0.temporal_res_block.conv1.weight"].shape) == (512, 512, 3, 1, 1)


def test_primitives_against_torch_modules():
    torch.manual_seed(0)
    C = 64
    # ResnetBlock2D(temb=None) with a 1x1 shortcut
    n1, c1, n2, c2 = nn.GroupNorm(32, C, eps=1e-6), nn.Conv2d(C, 128, 3, padding=1), nn.GroupNorm(32, 128, eps=1e-6), \
        nn.Conv2d(128, 128, 3, padding=1)
    sc = nn.Conv2d(C, 128, 1)
    sd = {}
    for name, mod in [("r.norm1", n1), ("r.conv1", c1), ("r.norm2", n2), ("r.conv2", c2), ("r.conv_shortcut", sc)]:
        sd[name + ".weight"], sd[name + ".bias"] = mod.weight.detach(), mod.bias.detach()
    x = torch.randn(2, C, 6, 5)
    with torch.no_grad():
        want = sc(x) + c2(F.silu(n2(c1(F.silu(n1(x))))))
        assert torch.allclose(VO.resnet_block_2d(sd, "r", x), want, atol=1e-5)
    # Attention block (one head) against nn.MultiheadAttention with the same projections
    gn = nn.GroupNorm(32, C, eps=1e-6)
    mha = nn.MultiheadAttention(C, 1, batch_first=True)
    wq, wk, wv = mha.in_proj_weight.detach().chunk(3)
    bq, bk, bv = mha.in_proj_bias.detach().chunk(3)
    sd = {"a.group_norm.weight": gn.weight.detach(), "a.group_norm.bias": gn.bias.detach(),
          "a.to_q.weight": wq, "a.to_q.bias": bq, "a.to_k.weight": wk, "a.to_k.bias": bk, "a.to_v.weight": wv,
          "a.to_v.bias": bv, "a.to_out.0.weight": mha.out_proj.weight.detach(), "a.to_out.0.bias": mha.out_proj.bias.detach()}
    with torch.no_grad():
        t = gn(x).flatten(2).transpose(1, 2)
        want = mha(t, t, t, need_weights=False)[0].transpose(1, 2).reshape(x.shape) + x
        assert torch.allclose(VO.attention_block(sd, "a", x), want, atol=1e-5)


def test_encoder_downsample_pads_bottom_right_only():
    """Downsample2D(padding=0): output pixel (i, j) reads input rows 2i..2i+2 / cols 2j..2j+2, zeros past the edge."""
    vae = build_vae(TINY_VAE)
    sd = state(vae)
    w, b = sd["encoder.down_blocks.0.downsamplers.0.conv.weight"], sd["encoder.down_blocks.0.downsamplers.0.conv.bias"]
    x = torch.randn(1, w.shape[1], 6, 8)
    got = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    assert got.shape[-2:] == (3, 4)
    i, j = 2, 3  # bottom-right output: its last row / column of taps fall on the padding
    patch = torch.zeros(w.shape[1], 3, 3)
    patch[:, :2, :2] = x[0, :, 4:6, 6:8]
    assert torch.allclose(got[0, :, i, j], (w * patch).sum((1, 2, 3)) + b, atol=1e-5)


def test_temporal_closed_forms():
    """(a) num_frames = 1: the 3-tap temporal convolutions and time_conv_out reduce to their centre tap, so decoding
    frame by frame equals a purely 2-D network with those 1x1 taps; (b) frames of different videos never mix;
    (c) mix_factor -> -inf removes the temporal branch (alpha = 1 on the spatial path)."""
    vae = build_vae(TINY_VAE)
    sd = state(vae)
    z, _ = vae_inputs(4, 4, 6)
    with torch.no_grad():
        one = VO.decode(sd, z, 1)
        # (a) a copy whose temporal kernels keep only the centre tap must give the same result for any grouping
        sd_c = dict(sd)
        for k, v in sd.items():
            if v.ndim == 5:
                c = torch.zeros_like(v)
                c[:, :, 1] = v[:, :, 1]
                sd_c[k] = c
        assert torch.allclose(VO.decode(sd_c, z, 1), one, atol=1e-5)
        # with centre-only kernels the only cross-frame path left is the 5-D GroupNorm statistics
        assert not torch.allclose(VO.decode(sd_c, z, 4), one, atol=1e-3)
        # (b) two videos of 2 frames == each video decoded alone
        two = VO.decode(sd, z, 2)
        assert torch.allclose(two[:2], VO.decode(sd, z[:2], 2), atol=1e-5)
        assert torch.allclose(two[2:], VO.decode(sd, z[2:], 2), atol=1e-5)
        # (c) no temporal branch: only time_conv_out still mixes frames
        sd_s = {k: (torch.full_like(v, -1e4) if k.endswith("mix_factor") else v) for k, v in sd.items()}
        w = sd["decoder.time_conv_out.weight"]
        sd_s1 = dict(sd_s)
        c = torch.zeros_like(w)
        c[:, :, 1] = w[:, :, 1]
        sd_s1["decoder.time_conv_out.weight"] = c
        assert torch.allclose(VO.decode(sd_s1, z, 4), VO.decode(sd_s1, z, 1), atol=1e-5)


def test_decode_latents_chunking_and_layout():
    vae = build_vae(TINY_VAE)
    sd = state(vae)
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(1, 6, 4, 4, 6, generator=g) * 0.18215
    with torch.no_grad():
        full = VO.decode_latents(sd, lat, 6, decode_chunk_size=6)
        chunked = VO.decode_latents(sd, lat, 6, decode_chunk_size=4)
        assert full.shape == chunked.shape == (1, 3, 6, 32, 48)
        # every chunk is decoded as its own short video (reference :266-275): frames 0-3 only see frames 0-3
        first = VO.decode(sd, lat[0, :4] / 0.18215, 4)
        assert torch.allclose(chunked[0, :, :4], first.permute(1, 0, 2, 3), atol=1e-5)
        assert not torch.allclose(full, chunked, atol=1e-3)


def test_golden_vectors():
    from tests.golden.make_vae_golden import compute
    gold = torch.load(GOLD / "tiny_vae.pt")
    out = compute()
    for k in gold:
        assert rel_l2(out[k], gold[k]) < 1e-5, k
