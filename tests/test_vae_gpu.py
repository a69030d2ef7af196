"""VAE path on the B200 (scope-table next #1): the sm_100a engine behind svd.autoencoder_kl_temporal_decoder against the
CPU fp32 oracle (oracle/vae_oracle.py) on the same seeded weights / inputs, through the module API the reference's
pipelines call (vae.encode(x).latent_dist.mode(), vae.decode(z, num_frames=n).sample) and through the raw C-ABI entry
points for the three VAE-only kernels.

Tolerance: rel-L2 over the tensor vs the fp32 oracle, err(engine) <= err(torch-eager bf16 of the same graph) + 1e-3 with
the eager error measured in the same test, plus the absolute cap 3e-2 (same bar as tests/test_parity_gpu.py)."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as VO
from tests.common import SVD_VAE, TINY_VAE, build_vae, rel_l2, state, vae_inputs

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
CAP = 3e-2


def _eager(fn, sd, *tensors):
    sdb = {k: v.to("cuda", torch.bfloat16) for k, v in sd.items()}
    return fn(sdb, *[t.to("cuda", torch.bfloat16) if torch.is_tensor(t) else t for t in tensors])


@pytest.fixture(scope="module")
def tiny():
    vae = build_vae(TINY_VAE)
    sd = state(vae)
    vae.to("cuda")
    return vae, sd


def test_vae_only_kernels_vs_torch():
    from this_and_that_vdm_b200 import lib
    lib.init()
    g = torch.Generator().manual_seed(0)
    # softmax_rows: ragged row length, padded output, row strides larger than the row
    rows, cols, ldx, cols_out, ldo = 37, 100, 104, 128, 136
    x = (torch.randn(rows, ldx, generator=g) * 4).cuda()
    out = torch.full((rows, ldo), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(x, out, rows=rows, cols=cols, ldx=ldx, ldo=ldo, cols_out=cols_out)
    want = torch.softmax(x[:, :cols].float(), -1)
    assert torch.allclose(out[:, :cols].float(), want, atol=4e-3, rtol=1e-2)
    assert float(out[:, cols:cols_out].abs().max()) == 0.0 and float(out[:, cols_out:].min()) == 7.0
    # odd sizes / strides: the scalar path
    x = (torch.randn(5, 41, generator=g) * 4).cuda()
    out = torch.full((5, 67), 7.0, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(x, out, rows=5, cols=37, ldx=41, ldo=67, cols_out=64)
    assert torch.allclose(out[:, :37].float(), torch.softmax(x[:, :37], -1), atol=4e-3, rtol=1e-2)
    assert float(out[:, 37:64].abs().max()) == 0.0 and float(out[:, 64:].min()) == 7.0
    # a long row (re-read variant is only taken beyond 51200 columns; the cached one must hold 9216 = 72x128 tokens)
    x = (torch.randn(3, 9216, generator=g) * 3).cuda()
    out = torch.empty(3, 9216, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(x, out, rows=3, cols=9216, ldx=9216, ldo=9216, cols_out=9216)
    assert torch.allclose(out.float(), torch.softmax(x, -1), atol=1e-4, rtol=1e-2)
    x = (torch.randn(2, 60000, generator=g) * 3).cuda()
    out = torch.empty(2, 60032, dtype=torch.bfloat16, device="cuda")
    lib.softmax_rows(x, out, rows=2, cols=60000, ldx=60000, ldo=60032, cols_out=60032)
    assert torch.allclose(out[:, :60000].float(), torch.softmax(x, -1), atol=1e-4, rtol=1e-2)
    # im2col_s2_pad01 is a pure gather: bit exact against unfold of the bottom/right padded image
    n, H, W, C = 2, 6, 10, 16
    xi = torch.randn(n, H, W, C, generator=g).to(torch.bfloat16).cuda()
    col = torch.empty(n * (H // 2) * (W // 2), 9 * C, dtype=torch.bfloat16, device="cuda")
    lib.im2col_s2_pad01(xi, col, n_img=n, H=H, W=W, C=C)
    u = F.unfold(F.pad(xi.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), 3, stride=2)
    u = u.view(n, C, 9, -1).permute(0, 3, 2, 1).reshape(col.shape)
    assert torch.equal(col.float(), u)
    # time_conv_out: fp32 conv over frames + NCHW scatter
    B, Fr, Hh, Ww = 2, 5, 4, 6
    xr = torch.randn(B * Fr * Hh * Ww, 4, generator=g).cuda()
    w, b = torch.randn(3, 3, 3, generator=g), torch.randn(3, generator=g)
    out = torch.empty(B * Fr, 3, Hh, Ww, device="cuda")
    lib.vae_time_conv_out(xr, w, b, out, B=B, F=Fr, H=Hh, W=Ww, ldx=4)
    x5 = xr[:, :3].view(B, Fr, Hh, Ww, 3).permute(0, 4, 1, 2, 3)
    want = F.conv3d(x5, w.view(3, 3, 3, 1, 1).cuda(), b.cuda(), padding=(1, 0, 0)).permute(0, 2, 1, 3, 4)
    assert torch.allclose(out.view(B, Fr, 3, Hh, Ww), want, atol=1e-5)
    torch.cuda.synchronize()


def test_decode_vs_oracle_eager_and_golden(tiny):
    vae, sd = tiny
    z, _ = vae_inputs(8, 8, 12)
    with torch.no_grad():
        ref = VO.decode(sd, z, 4)
        eager = _eager(VO.decode, sd, z, 4)
        out = vae.decode(z.cuda(), num_frames=4).sample
    e, ee = rel_l2(out, ref), rel_l2(eager, ref)
    assert out.shape == (8, 3, 64, 96) and out.dtype == torch.float32
    assert e <= ee + 1e-3 and e < CAP, (e, ee)
    assert rel_l2(out, torch.load(GOLD / "tiny_vae.pt")["decode"]) < CAP


def test_encode_vs_oracle_eager_and_golden(tiny):
    vae, sd = tiny
    _, x = vae_inputs(8, 8, 12)
    with torch.no_grad():
        ref = VO.encode(sd, x)
        eager = _eager(VO.encode, sd, x)
        dist = vae.encode(x.cuda()).latent_dist
    out = dist.mode()
    e, ee = rel_l2(out, ref), rel_l2(eager, ref)
    assert out.shape == (2, 4, 8, 12) and dist.sample(torch.Generator("cuda").manual_seed(0)).shape == out.shape
    assert e <= ee + 1e-3 and e < CAP, (e, ee)
    assert rel_l2(out, torch.load(GOLD / "tiny_vae.pt")["encode_mean"]) < CAP


@pytest.mark.parametrize("n_videos,frames,lh,lw", [(1, 1, 8, 8), (2, 3, 4, 6), (1, 14, 4, 4)])
def test_decode_shapes_and_video_independence(tiny, n_videos, frames, lh, lw):
    """One frame, several videos per call (frames of different videos never mix) and the reference's 14-frame chunk;
    4x6 latents give 24 tokens per frame: ragged against every tile size of the attention GEMMs."""
    vae, sd = tiny
    z, _ = vae_inputs(n_videos * frames, lh, lw)
    with torch.no_grad():
        ref = VO.decode(sd, z, frames)
        out = vae.decode(z.cuda(), num_frames=frames).sample
        assert rel_l2(out, ref) < CAP
        if n_videos > 1:
            alone = vae.decode(z[:frames].cuda(), num_frames=frames).sample
            assert rel_l2(out[:frames], alone) < 1e-5


def test_svd_config_vae_and_pipeline_decode_latents():
    """The published SVD VAE shape (128/256/512/512, one head of 512 dims in the mid blocks) on a 64x64 px clip, through
    decode_latents of the drop-in pipeline (svd/pipeline_stable_video_diffusion_controlnet.py:257-283): /scaling_factor,
    chunks of decode_chunk_size frames each decoded as its own short video, fp32 [B, 3, F, H, W]."""
    from svd.pipeline_common import SVDPipelineBase
    vae = build_vae(SVD_VAE)
    sd = state(vae)
    vae.to("cuda")
    g = torch.Generator().manual_seed(2)
    lat = torch.randn(1, 3, 4, 8, 8, generator=g) * 0.18215
    x = torch.rand(1, 3, 64, 64, generator=g) * 2 - 1
    pipe = SVDPipelineBase(vae=vae, unet=None)
    with torch.no_grad():
        ref = VO.decode_latents(sd, lat, 3, decode_chunk_size=2)
        frames = pipe.decode_latents(lat.cuda(), 3, decode_chunk_size=2)
        ref_e = VO.encode(sd, x)
        enc = vae.encode(x.cuda()).latent_dist.mode()
    assert frames.shape == (1, 3, 3, 64, 64) and frames.dtype == torch.float32
    assert rel_l2(frames, ref) < CAP and rel_l2(enc, ref_e) < CAP


def test_encode_dedupe_and_errors(tiny):
    vae, sd = tiny
    _, x = vae_inputs(1, 4, 6, n_images=2)
    zero = torch.zeros_like(x[:1])
    batch = torch.cat([zero, x[:1], zero, zero, x[1:], zero]).cuda()
    eng = vae._get_engine()
    with torch.no_grad():
        a = eng.encode(batch)
        b = eng.encode(batch, dedupe=False)
    assert torch.equal(a[0], a[2]) and rel_l2(a, b) < 1e-5
    assert rel_l2(a[:, :4], VO.encode(sd, batch.cpu())) < CAP
    with pytest.raises(ValueError, match="multiples of 8"):
        vae.encode(torch.zeros(1, 3, 20, 24, device="cuda"))
    with pytest.raises(ValueError, match="multiple"):
        vae.decode(torch.zeros(3, 4, 4, 4, device="cuda"), num_frames=2)
