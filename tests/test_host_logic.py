"""Host logic of the engines on CPU (`-m "not gpu"`): weight packing, leading dimensions, buffer reuse, epilogue-fusion
arguments and the kernel schedule of this_and_that_vdm_b200/engine.py + sampler.py, run through tests/fake_lib.py
(a torch emulation of the C ABI, test infrastructure only) and compared with the fp32 oracle on the same seeded
inputs. The same engine code runs on the B200 against the real kernels in tests/test_parity_gpu.py; these tests also
validate the emulation itself (the UNet / GestureNet engine is GPU-proven), which the VAE host-logic tests rely on.

Tolerance: the emulation rounds to bf16 wherever the kernels store bf16, so the bar is the GPU one (rel-L2 < 3e-2
against the fp32 oracle)."""
import pytest
import torch

from oracle import svd_oracle as O
from tests import fake_lib
from tests.common import TINY, build_models, make_inputs, oracle_cfg, rel_l2, state
from this_and_that_vdm_b200.engine import DenoiserEngine
from this_and_that_vdm_b200.sampler import FusedDenoiser

CAP = 3e-2
T0 = torch.tensor(1.63777)


@pytest.fixture(scope="module")
def tiny():
    unet, cn = build_models(TINY)
    with fake_lib.installed():
        eu, ec = DenoiserEngine(unet, "unet"), DenoiserEngine(cn, "controlnet")
    return eu, ec, state(unet), state(cn), oracle_cfg(TINY)


def test_unet_schedule_vs_oracle(tiny):
    eu, ec, usd, csd, cfg = tiny
    sample, ehs, ati, cond = make_inputs(2, 14, 8, 16)
    with torch.no_grad(), fake_lib.installed():
        n0 = fake_lib.launch_count()
        out = eu.unet_forward(sample, T0, ehs, ati)
        launches = fake_lib.launch_count() - n0
        ref = O.unet_forward(usd, cfg, sample, T0, ehs, ati)
    assert out.shape == ref.shape == (2, 14, 4, 8, 16)
    assert rel_l2(out, ref) < CAP
    assert launches > 500  # one call per kernel of the schedule; none skipped


def test_controlnet_and_residual_merge_schedule_vs_oracle(tiny):
    eu, ec, usd, csd, cfg = tiny
    sample, ehs, ati, cond = make_inputs(2, 14, 8, 8)
    cc = torch.cat([cond, cond])
    with torch.no_grad(), fake_lib.installed():
        d, m = ec.controlnet_forward(sample, T0, ehs, ati, cc, 0.8)
        y = eu.unet_forward(sample, T0, ehs, ati, d, m)
        d_ref, m_ref = O.controlnet_forward(csd, cfg, sample, T0, ehs, ati, cc, 0.8)
        y_ref = O.unet_forward(usd, cfg, sample, T0, ehs, ati, d_ref, m_ref)
    assert len(d) == 12 and rel_l2(m, m_ref) < CAP
    for a, b in zip(d, d_ref):
        assert a.shape == b.shape and rel_l2(a, b) < CAP
    assert rel_l2(y, y_ref) < CAP


@pytest.mark.parametrize("b_local,offset", [(2, 0), (1, 0), (1, 1)])
def test_fused_vgl_step_and_split_pairs_vs_oracle(tiny, b_local, offset):
    """FusedDenoiser: hoisted embeddings / context K,V, zero-conv accumulation into the UNet skips, sampler glue; with
    b_local = 1 the rank holds one half of the CFG pair and must index the context quirk by the GLOBAL batch index."""
    eu, ec, usd, csd, cfg = tiny
    F, h, w = 14, 8, 8
    sample, ehs, ati, cond = make_inputs(2, F, h, w)
    g = torch.Generator().manual_seed(3)
    sig = O.karras_sigmas(25)
    ts = O.euler_timesteps(sig)
    i = 12
    lat = torch.randn(F, 4, h, w, generator=g) * float(sig[i])
    img = torch.randn(1, 4, h, w, generator=g)
    img2 = torch.cat([torch.zeros_like(img), img])
    with torch.no_grad(), fake_lib.installed():
        den = FusedDenoiser(eu, ec)
        den.prepare(ehs, img2, ati, sig, ts, torch.linspace(1, 3, F), num_frames=F, height=h, width=w,
                    controlnet_cond=cond, batch_offset=offset, b_local=b_local)
        eps = den.predict(i, lat.clone())
        xin = torch.cat([lat[None]] * 2) / ((float(sig[i]) ** 2 + 1) ** 0.5)
        xin = torch.cat([xin, img2[:, None].repeat(1, F, 1, 1, 1)], dim=2)
        d, m = O.controlnet_forward(csd, cfg, xin, ts[i], ehs, ati, torch.cat([cond, cond]), 1.0)
        ref = O.unet_forward(usd, cfg, xin, ts[i], ehs, ati, d, m)
    got = eps.view(b_local, F, h, w, 4).permute(0, 1, 4, 2, 3)
    assert rel_l2(got, ref[offset:offset + b_local]) < CAP
    if b_local == 2:
        with fake_lib.installed():
            st = lat.clone()
            den.euler_update(i, st, eps[:F * h * w], eps[F * h * w:])
        eu_, ec_ = ref.chunk(2)
        gd = torch.linspace(1, 3, F)[None, :, None, None, None]
        want = O.euler_step(eu_ + gd * (ec_ - eu_), lat[None], float(sig[i]), float(sig[i + 1]))
        assert rel_l2(st, want[0]) < CAP


def test_emulation_is_not_reachable_from_the_product():
    """The wrappers are restored when the context exits: outside it the product still needs the real library."""
    from this_and_that_vdm_b200 import lib
    with fake_lib.installed():
        assert lib.gemm is fake_lib.gemm
    assert lib.gemm is not fake_lib.gemm and lib.gemm.__module__ == "this_and_that_vdm_b200.lib"
    if not torch.cuda.is_available():
        with pytest.raises(lib.TtvdmError):
            lib.init()  # no CUDA device here: the product path fails loudly


def test_emulation_covers_every_kernel_wrapper():
    """A new entry point in lib.py must come with its emulation, or the host-logic tests silently stop covering it."""
    import inspect
    from this_and_that_vdm_b200 import lib
    wrappers = {n for n, f in vars(lib).items()
                if inspect.isfunction(f) and f.__module__ == lib.__name__ and not n.startswith("_")
                and ("call(" in inspect.getsource(f) or "call_raw(" in inspect.getsource(f))
                and n not in ("call", "call_raw")}
    not_emulated = wrappers - set(fake_lib._PATCHED)
    assert not_emulated == {"gesture_raster"}, not_emulated  # the rasteriser has its own oracle / golden tests
    for n in set(fake_lib._PATCHED) & wrappers:
        real = [p for p in inspect.signature(getattr(lib, n)).parameters]
        fake = [p for p in inspect.signature(getattr(fake_lib, n)).parameters]
        assert real == fake, (n, real, fake)


def test_layernorm_fold_schedule_vs_oracle(monkeypatch):
    """TTVDM_FUSE_LN=1: every LayerNorm folded into its consuming GEMM (gamma / beta in the packed weights, row sums from
    the producer's epilogue, the frame positional embedding through rs_addvec / prevec / rowvec) and a two-layer
    transformer stack (transformer_layers_per_block = 2) — against the oracle, on the CPU emulation of the C ABI."""
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from tests import refpin
    from tests.test_reference_pin import _models
    monkeypatch.setenv("TTVDM_FUSE_LN", "1")
    kind, B, F, h, w = refpin.CASES["tiny_2layers"]
    unet = UNetSpatioTemporalConditionModel(num_frames=F, **kind).eval()
    usd, _ = _models(kind, F)
    unet.load_state_dict(usd)
    cfg = dict(O.SVD_CONFIG)
    cfg.update({k: v for k, v in kind.items() if k in cfg})
    sample, ehs, ati, _ = refpin.make_inputs(B, F, h, w)
    with torch.no_grad(), fake_lib.installed():
        eng = DenoiserEngine(unet, "unet")
        assert eng.fuse_layernorm and len(eng.down[0]["tf"][0].layers) == 2
        n0 = fake_lib.launch_count()
        out = eng.unet_forward(sample, torch.tensor(refpin.TIMESTEP), ehs, ati)
        folded_launches = fake_lib.launch_count() - n0
        monkeypatch.setenv("TTVDM_FUSE_LN", "0")
        eng0 = DenoiserEngine(unet, "unet")
        n0 = fake_lib.launch_count()
        out0 = eng0.unet_forward(sample, torch.tensor(refpin.TIMESTEP), ehs, ati)
        plain_launches = fake_lib.launch_count() - n0
        ref = O.unet_forward(usd, cfg, sample, refpin.TIMESTEP, ehs, ati)
    assert rel_l2(out, ref) < CAP and rel_l2(out0, ref) < CAP
    assert rel_l2(out, out0) < 2 * CAP  # two independent bf16 roundings of the same graph
    n_tf = sum(len(b["tf"]) for b in eng.down + [eng.mid] + eng.up)
    # exactly the 7 LayerNorm launches per layer disappear; the folded engine's first forward also runs its one-time
    # positional-embedding projection (one cast per transformer + one GEMM per layer)
    assert plain_launches - folded_launches == 7 * 2 * n_tf - (n_tf + 2 * n_tf)


def test_num_frames_above_kernel_limit_is_rejected_up_front():
    """a 25-frame (SVD-XT) call must fail with a clear message at the boundary, not with a shape error deep
    inside the first temporal transformer block."""
    import pytest
    from this_and_that_vdm_b200 import lib
    from this_and_that_vdm_b200.engine import MAX_FRAMES
    from this_and_that_vdm_b200.sampler import FusedDenoiser

    class _E:  # enough of an engine for prepare() to reach the check
        device = torch.device("cpu")

    den = FusedDenoiser(_E(), None, use_graph=False)
    with pytest.raises(lib.TtvdmError, match="num_frames"):
        den.prepare(torch.zeros(2, 78, 1024), torch.zeros(2, 4, 8, 8), torch.zeros(2, 3), torch.ones(26), torch.ones(25),
                    torch.ones(MAX_FRAMES + 9), num_frames=MAX_FRAMES + 9, height=8, width=8)


def test_head_dim_128_compatibility_path_vs_oracle():
    """The reference UNet's class-default heads (5, 10, 10, 20) give head_dim 128 at level 2
    (svd/unet_spatio_temporal_condition.py:99). Tiny analogue: 256 channels / 2 heads at level 2. The engine takes the
    GEMM + row-softmax path for the spatial self-attention and the head_dim-128 variants of the warp-level cross /
    temporal attention kernels; compared with the oracle (which is pinned to the reference for this head variant,
    tests/test_reference_pin.py)."""
    kind = dict(TINY, num_attention_heads=(1, 2, 2, 4))
    unet, cn = build_models(kind)
    sample, ehs, ati, cond = make_inputs(2, 4, 8, 8)
    with torch.no_grad(), fake_lib.installed():
        eu, ec = DenoiserEngine(unet, "unet"), DenoiserEngine(cn, "controlnet")
        assert [t.hd for blk in eu.down for t in blk["tf"]] == [64, 64, 64, 64, 128, 128]
        d, m = ec.controlnet_forward(sample, T0, ehs, ati, torch.cat([cond[:4], cond[:4]]), 1.0)
        y = eu.unet_forward(sample, T0, ehs, ati, d, m)
        cfg = oracle_cfg(kind)
        d_ref, m_ref = O.controlnet_forward(state(cn), cfg, sample, T0, ehs, ati, torch.cat([cond[:4], cond[:4]]), 1.0)
        y_ref = O.unet_forward(state(unet), cfg, sample, T0, ehs, ati, d_ref, m_ref)
    assert rel_l2(m, m_ref) < CAP and rel_l2(y, y_ref) < CAP, (rel_l2(m, m_ref), rel_l2(y, y_ref))


def test_25_frames_svd_xt_schedule_vs_oracle(tiny):
    """SVD-XT's 25-frame setting (num_frames is a runtime argument of the reference pipelines): frame positional
    embeddings, 5-D GroupNorm instances, temporal conv and the temporal attention's two 16-frame tiles."""
    eu, ec, usd, csd, cfg = tiny
    sample, ehs, ati, cond = make_inputs(2, 25, 8, 8)
    with torch.no_grad(), fake_lib.installed():
        out = eu.unet_forward(sample, T0, ehs, ati)
        ref = O.unet_forward(usd, cfg, sample, T0, ehs, ati)
    assert out.shape == ref.shape == (2, 25, 4, 8, 8)
    assert rel_l2(out, ref) < CAP


def test_upsample_parity_weights_identity_fp32():
    """conv3x3(nearest_x2(x), w, padding=1) == the four 2x2-tap parity convolutions of x with the pre-summed weights of
    engine._pack_upsample_parity, interleaved — checked in plain fp32 torch (no kernels, no emulation): pins the tap
    grouping ({y-1: ky 0, y: ky 1+2} for even output rows, {y: ky 0+1, y+1: ky 2} for odd ones)."""
    import torch.nn.functional as Fn
    g = torch.Generator().manual_seed(5)
    n, C, N, H, W = 2, 6, 5, 5, 7
    x = torch.randn(n, C, H, W, generator=g)
    w = torch.randn(N, C, 3, 3, generator=g)
    ref = Fn.conv2d(Fn.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = torch.zeros(n, N, 2 * H, 2 * W)
    for py in (0, 1):
        for px in (0, 1):
            wp = torch.zeros(N, C, 2, 2)
            for ty in (0, 1):
                for tx in (0, 1):
                    for ky in groups[py][ty]:
                        for kx in groups[px][tx]:
                            wp[:, :, ty, tx] += w[:, :, ky, kx]
            # window whose first tap sits at (py - 1, px - 1) relative to the output pixel: pad accordingly
            xp = Fn.pad(x, (1 - px, px, 1 - py, py))
            out[:, :, py::2, px::2] = Fn.conv2d(xp, wp)
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)
