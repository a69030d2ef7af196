"""Host logic of this_and_that_vdm_b200/vae_engine.py on CPU (`-m "not gpu"`): packing (quant_conv fold, conv_out
padding, AlphaBlender fold), the per-frame attention schedule with its padded P / V^T buffers, chunked per-image
launches and the drop-in module surface — run through tests/fake_lib.py (torch emulation of the C ABI, validated by
tests/test_host_logic.py on the GPU-proven UNet engine) against oracle/vae_oracle.py. The same engine code runs on
the B200 against the real kernels in tests/test_vae_gpu.py. Tolerance: rel-L2 < 3e-2 vs the fp32 oracle (bf16 storage)."""
import pytest
import torch

from oracle import vae_oracle as VO
from tests import fake_lib
from tests.common import TINY_VAE, build_vae, rel_l2, state, vae_inputs
from this_and_that_vdm_b200 import vae_engine
from this_and_that_vdm_b200.vae_engine import VaeEngine

CAP = 3e-2


@pytest.fixture(scope="module")
def tiny():
    vae = build_vae(TINY_VAE)
    with fake_lib.installed():
        eng = VaeEngine(vae)
    return vae, eng, state(vae)


@pytest.mark.parametrize("n_videos,frames,lh,lw", [(1, 4, 8, 12), (2, 3, 4, 6), (1, 1, 5, 7)])
def test_decode_schedule_vs_oracle(tiny, n_videos, frames, lh, lw):
    vae, eng, sd = tiny
    z, _ = vae_inputs(n_videos * frames, lh, lw)
    with torch.no_grad(), fake_lib.installed():
        out = eng.decode(z, frames)
        ref = VO.decode(sd, z, frames)
    assert out.shape == ref.shape == (n_videos * frames, 3, 8 * lh, 8 * lw)
    assert rel_l2(out, ref) < CAP


@pytest.mark.parametrize("n,lh,lw", [(2, 8, 12), (5, 3, 5)])
def test_encode_schedule_vs_oracle(tiny, n, lh, lw):
    vae, eng, sd = tiny
    _, x = vae_inputs(1, lh, lw, n_images=n)
    with torch.no_grad(), fake_lib.installed():
        mom = eng.encode(x, max_images_per_pass=2)
        ref = VO.encode(sd, x)
    assert mom.shape == (n, 8, lh, lw)
    assert rel_l2(mom[:, :4], ref) < CAP


def test_chunked_launches_give_the_same_result(tiny, monkeypatch):
    """Per-image ops are split so that no launch sees more than _MAX_ELEMS elements; force many groups."""
    vae, eng, sd = tiny
    z, _ = vae_inputs(4, 4, 6)
    with torch.no_grad(), fake_lib.installed():
        whole = eng.decode(z, 4)
        monkeypatch.setattr(vae_engine, "_MAX_ELEMS", 4 * 6 * 64 * 3)
        split = eng.decode(z, 4)
        ref = VO.decode(sd, z, 4)
    # same arithmetic per image; torch's CPU kernels block differently per batch size, so bf16 roundings flip and the
    # two runs differ by bf16 noise — both must sit within the bar of the oracle
    assert rel_l2(split, ref) < CAP and rel_l2(whole, ref) < CAP and rel_l2(split, whole) < CAP


def test_module_surface_matches_diffusers_calls(tiny, monkeypatch):
    """encode(x).latent_dist.mode(), decode(z, num_frames=n).sample, forward(num_frames=) as the pipelines call them
    (svd/pipeline_stable_video_diffusion_controlnet.py:199, :257-283); CPU modules must refuse to run."""
    import inspect
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    vae, eng, sd = tiny
    assert "num_frames" in inspect.signature(vae.forward).parameters
    assert vae.config.scaling_factor == 0.18215 and len(vae.config.block_out_channels) == 4
    z, x = vae_inputs(2, 4, 6)
    with pytest.raises(RuntimeError, match="CUDA"):
        vae.encode(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        vae.decode(z, num_frames=2)
    with pytest.raises(ValueError, match="same number"):
        AutoencoderKLTemporalDecoder(down_block_types=("DownEncoderBlock2D",) * 2, block_out_channels=(64,))
    monkeypatch.setattr(type(vae), "_get_engine", lambda self: eng)
    with torch.no_grad(), fake_lib.installed():
        lat = vae.encode(x).latent_dist.mode()
        dec = vae.decode(z, num_frames=2).sample
        with pytest.raises(ValueError, match="multiple"):
            vae.decode(z[:1].repeat(3, 1, 1, 1), num_frames=2)
        ref_lat, ref_dec = VO.encode(sd, x), VO.decode(sd, z, 2)
    assert rel_l2(lat, ref_lat) < CAP and rel_l2(dec, ref_dec) < CAP


def test_pipeline_decode_latents_through_the_engine(tiny, monkeypatch):
    """decode_latents of the drop-in pipelines: /scaling_factor, chunks of decode_chunk_size frames, [B, C, F, H, W]."""
    from svd.pipeline_common import SVDPipelineBase
    vae, eng, sd = tiny
    monkeypatch.setattr(type(vae), "_get_engine", lambda self: eng)
    pipe = SVDPipelineBase(vae=vae, unet=None)
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(1, 6, 4, 4, 6, generator=g) * 0.18215
    with torch.no_grad(), fake_lib.installed():
        frames = pipe.decode_latents(lat, 6, decode_chunk_size=4)
        ref = VO.decode_latents(sd, lat, 6, decode_chunk_size=4)
    assert frames.shape == (1, 3, 6, 32, 48) and frames.dtype == torch.float32
    assert rel_l2(frames, ref) < CAP


def test_encode_dedupes_identical_images(tiny):
    """12 of the reference's 14 gesture frames are all-zero: identical images are encoded once and scattered back."""
    vae, eng, sd = tiny
    _, x = vae_inputs(1, 4, 6, n_images=2)
    zero = torch.zeros_like(x[:1])
    batch = torch.cat([zero, x[:1], zero, zero, x[1:], zero])
    with torch.no_grad(), fake_lib.installed():
        n0 = fake_lib.launch_count()
        a = eng.encode(batch)
        n_dedup = fake_lib.launch_count() - n0
        b = eng.encode(batch, dedupe=False)
        n_all = fake_lib.launch_count() - n0 - n_dedup
    assert a.shape == b.shape == (6, 8, 4, 6)
    assert torch.equal(a[0], a[2]) and torch.equal(a[0], a[5]) and not torch.equal(a[0], a[1])
    assert rel_l2(a, b) < CAP and rel_l2(a[:, :4], VO.encode(sd, batch)) < CAP
    assert n_dedup * 2 == n_all  # 3 distinct images instead of 6


def test_fp16_module_keeps_dtype_and_accuracy(monkeypatch):
    """The reference runs its VAE in fp16 (test_code/inference.py:364): outputs come back in the caller's dtype."""
    vae = build_vae(TINY_VAE).half()
    sd = {k: v.float() for k, v in vae.state_dict().items()}
    z, x = vae_inputs(4, 4, 6)
    with torch.no_grad(), fake_lib.installed():
        eng = VaeEngine(vae)
        dec = eng.decode(z.half(), 2)
        mom = eng.encode(x.half())
    assert dec.dtype == torch.float16 and mom.dtype == torch.float16
    assert rel_l2(dec, VO.decode(sd, z.half().float(), 2)) < CAP and rel_l2(mom[:, :4], VO.encode(sd, x.half().float())) < CAP


def test_save_and_from_pretrained_round_trip(tmp_path):
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    vae = build_vae(TINY_VAE)
    vae.save_pretrained(tmp_path / "vae")
    again = AutoencoderKLTemporalDecoder.from_pretrained(str(tmp_path), subfolder="vae")
    assert again.config.block_out_channels == (64, 128, 128, 128) and again.config.layers_per_block == 1
    a, b = vae.state_dict(), again.state_dict()
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_decoder_mid_block_attention_is_scheduled_with_two_layers_per_block():
    """layers_per_block = 2 (the SVD value): the decoder's mid block is ResBlock -> Attention -> ResBlock and every up
    block has 3 ResBlocks; with layers_per_block = 1 the mid-block attention is never reached (diffusers zips
    resnets[1:] with the attentions)."""
    cfg = dict(block_out_channels=(64, 64, 128, 128), layers_per_block=2, down_block_types=("DownEncoderBlock2D",) * 4)
    vae = build_vae(cfg, seed=99)
    sd = state(vae)
    z, x = vae_inputs(3, 4, 6, n_images=1)
    with torch.no_grad(), fake_lib.installed():
        eng = VaeEngine(vae)
        assert len(eng.d_mid_res) == 2 and len(eng.d_mid_attn) == 1 and all(len(b["res"]) == 3 for b in eng.d_up)
        dec = eng.decode(z, 3)
        ref = VO.decode(sd, z, 3)
        # knocking out the attention's output projection must change the result: the block is really on the path
        sd2 = dict(sd)
        sd2["decoder.mid_block.attentions.0.to_out.0.weight"] = torch.zeros_like(sd["decoder.mid_block.attentions.0.to_out.0.weight"])
        assert rel_l2(VO.decode(sd2, z, 3), ref) > 1e-3
        mom = eng.encode(x)
    assert rel_l2(dec, ref) < CAP and rel_l2(mom[:, :4], VO.encode(sd, x)) < CAP


def test_flop_census_matches_the_schedule(tiny):
    """tools/flop_census.vae_{de,en}code_flops (the algorithmic figure DESIGN.md quotes) against the FLOPs of the GEMM
    calls the engine really makes (the engine pads conv_in to 64 input channels, conv_out to 4 / 8 outputs and the
    attention's key dimension to a multiple of 64, so it may only be slightly ABOVE the census)."""
    from tools.flop_census import vae_decode_flops, vae_encode_flops
    vae, eng, sd = tiny
    counted = []
    real_gemm = fake_lib.gemm

    def counting_gemm(a, w, out, **kw):
        taps = {0: 1, 1: 9, 2: 3}[kw.get("mode", 0)]
        counted.append(2.0 * kw["M"] * kw["N"] * taps * (kw["k1"] + kw.get("k2", 0)))
        return real_gemm(a, w, out, **kw)

    z, x = vae_inputs(4, 8, 8, n_images=2)
    kw = dict(chans=(64, 128, 128, 128), layers_per_block=1)
    with torch.no_grad(), fake_lib.installed():
        from this_and_that_vdm_b200 import lib
        lib.gemm = counting_gemm
        eng.decode(z, 4)
        dec = sum(counted)
        counted.clear()
        eng.encode(x, dedupe=False)
        enc = sum(counted)
    a_dec, a_enc = vae_decode_flops(8, 8, 4, **kw), vae_encode_flops(64, 64, 2, **kw)
    assert a_dec <= dec <= 1.06 * a_dec, (dec, a_dec)
    pad_in = 2 * 2.0 * 64 * 64 * 9 * (64 - 3) * 64  # the 3 input channels travel as one 64-wide K chunk (2 images)
    assert a_enc <= enc <= 1.02 * a_enc + pad_in, (enc, a_enc, pad_in)
    # the published size: 14 x 576 x 1024
    assert abs(vae_decode_flops(72, 128, 14) / 1e12 - 97.3) < 1.5
