"""Host-side boundary tests (no GPU): the C-ABI library loads and exports every symbol include/ttvdm.h declares, the
ctypes structs mirror the header, the drop-in svd.* classes keep the reference's constructor / error behaviour, and the
product never imports the oracle or falls back to CPU."""
import ctypes
import re
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _header():
    return (ROOT / "include" / "ttvdm.h").read_text()


def test_library_exports_every_declared_symbol():
    from this_and_that_vdm_b200 import lib
    l = lib.load()
    declared = set(re.findall(r"^\s*(?:int|uint64_t|size_t)\s+(ttvdm_\w+)\s*\(", _header(), flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(l, name), name
    assert l.ttvdm_abi_version() == 4
    assert lib.launch_count() == 0 or lib.launch_count() > 0


def test_ctypes_structs_match_header_sizes(tmp_path):
    """Compile a tiny C program against include/ttvdm.h and compare sizeof() with the ctypes mirrors."""
    from this_and_that_vdm_b200 import lib
    src = tmp_path / "sz.c"
    names = ["gemm", "attn", "xattn", "tattn", "groupnorm", "layernorm", "prepare", "euler", "pack_linear"]
    body = "".join(f'printf("%zu\\n", sizeof(ttvdm_{n}_params));' for n in names)
    src.write_text(f'#include <stdio.h>\n#include "ttvdm.h"\nint main(){{{body}return 0;}}')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [lib.GemmParams, lib.AttnParams, lib.XAttnParams, lib.TAttnParams, lib.GroupNormParams,
               lib.LayerNormParams, lib.PrepareParams, lib.EulerParams, lib.PackLinearParams]
    assert sizes == [ctypes.sizeof(m) for m in mirrors]


def test_no_cpu_fallback_and_no_oracle_import_in_product():
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from tests.common import TINY, make_inputs
    unet = UNetSpatioTemporalConditionModel(num_frames=14, **TINY)
    sample, ehs, ati, _ = make_inputs(1, 14, 8, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        unet(sample, 1.0, ehs, ati)
    with pytest.raises(RuntimeError, match="parameter container"):
        unet.down_blocks[0](sample)
    for path in list((ROOT / "svd").rglob("*.py")) + list((ROOT / "this_and_that_vdm_b200").rglob("*.py")):
        txt = path.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, path
        assert "fake_lib" not in txt and "from tests" not in txt and "import tests" not in txt, path
    if not torch.cuda.is_available():
        from this_and_that_vdm_b200 import lib
        with pytest.raises(lib.TtvdmError):
            lib.init()


def test_constructor_errors_match_reference():
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    with pytest.raises(ValueError, match="same number of `down_block_types` as `up_block_types`"):
        UNetSpatioTemporalConditionModel(up_block_types=("UpBlockSpatioTemporal",))
    with pytest.raises(ValueError, match="`block_out_channels` as `down_block_types`"):
        UNetSpatioTemporalConditionModel(block_out_channels=(64, 128))
    with pytest.raises(ValueError, match="`num_attention_heads` as `down_block_types`"):
        UNetSpatioTemporalConditionModel(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2))
    with pytest.raises(ValueError, match="`layers_per_block` as `down_block_types`"):
        ControlNetModel(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), layers_per_block=(2, 2))
    with pytest.raises(ValueError, match="does not exist"):
        UNetSpatioTemporalConditionModel(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                                         down_block_types=("Nope",) * 4)


def test_config_and_attributes_used_by_callers():
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from tests.common import TINY
    u = UNetSpatioTemporalConditionModel(num_frames=14, sample_size=96, **TINY)
    assert u.config.in_channels == 8 and u.config.addition_time_embed_dim == 256 and u.config.num_frames == 14
    assert u.config.sample_size == 96 and u.add_embedding.linear_1.in_features == 768
    assert u.dtype == torch.float32 and u.device.type == "cpu"
    assert u.config["block_out_channels"] == TINY["block_out_channels"]
    u.set_attn_processor(None)
    u.enable_forward_chunking()
    with pytest.raises(ValueError):
        u.enable_forward_chunking(dim=2)


def test_save_and_from_pretrained_roundtrip(tmp_path):
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from tests.common import TINY
    u = UNetSpatioTemporalConditionModel(num_frames=14, **TINY)
    u.save_pretrained(str(tmp_path / "repo" / "unet"))
    u2 = UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path / "repo"), subfolder="unet",
                                                          low_cpu_mem_usage=True, variant="fp16")
    assert u2.config.block_out_channels == TINY["block_out_channels"]
    for (k, a), (_, b) in zip(u.state_dict().items(), u2.state_dict().items()):
        assert torch.equal(a, b), k
    c = ControlNetModel(**TINY)
    c.save_pretrained(str(tmp_path / "repo" / "gesturenet"))
    c2 = ControlNetModel.from_pretrained(str(tmp_path / "repo"), subfolder="gesturenet")
    assert set(c2.state_dict()) == set(c.state_dict())
    with pytest.raises(EnvironmentError):
        UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path / "nowhere"))


def test_pipeline_argument_errors():
    from svd.pipeline_stable_video_diffusion import StableVideoDiffusionPipeline
    from svd.pipeline_stable_video_diffusion_controlnet import StableVideoDiffusionControlNetPipeline
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from tests.common import TINY
    u = UNetSpatioTemporalConditionModel(num_frames=14, **TINY)
    pipe = StableVideoDiffusionPipeline.from_pretrained("unused", unet=u)
    pipe.set_progress_bar_config(disable=True)
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe.check_inputs(torch.zeros(1, 3, 60, 64), 60, 64)
    with pytest.raises(ValueError, match="`image` has to be of type"):
        pipe.check_inputs("not an image", 64, 64)
    g = [torch.Generator().manual_seed(0)] * 3
    with pytest.raises(ValueError, match="list of generators"):
        pipe.prepare_latents(2, 14, 8, 64, 64, torch.float32, "cpu", g)
    lat = pipe.prepare_latents(1, 14, 8, 64, 96, torch.float32, "cpu", torch.Generator().manual_seed(0))
    assert lat.shape == (1, 14, 4, 8, 12) and abs(float(lat.std()) / 700.0 - 1) < 0.1
    ids = pipe._get_add_time_ids(6, 200, 0.1, torch.float32, 1, 1, True)
    assert ids.shape == (2, 3) and ids[0].tolist() == pytest.approx([6, 200, 0.1])
    with pytest.raises(EnvironmentError):
        StableVideoDiffusionControlNetPipeline.from_pretrained("unused")
    vgl = StableVideoDiffusionControlNetPipeline.from_pretrained("unused", unet=u)
    with pytest.raises(ValueError, match="controlnet"):
        vgl(height=64, width=64, num_frames=14, encoder_hidden_states=torch.zeros(2, 78, 1024))
    import inspect
    params = inspect.signature(vgl.__call__).parameters
    for k in ["image", "condition_img", "controlnet", "prompt", "use_text", "text_encoder", "height", "width",
              "num_frames", "num_inference_steps", "min_guidance_scale", "max_guidance_scale", "fps", "motion_bucket_id",
              "noise_aug_strength", "decode_chunk_size", "num_videos_per_prompt", "generator", "latents", "output_type",
              "callback_on_step_end", "callback_on_step_end_tensor_inputs", "return_dict",
              "controlnet_conditioning_scale", "use_instructpix2pix", "control_guidance_start", "control_guidance_end",
              "inner_conditioning_scale", "guess_mode", "image_guidance_scale"]:
        assert k in params, k
    assert params["height"].default == 576 and params["width"].default == 1024 and params["guess_mode"].default is True
    assert params["motion_bucket_id"].default == 127 and params["noise_aug_strength"].default == 0.02


def test_weight_packing_layouts():
    """Layouts the engine asks the repack entry points for (ttvdm_pack_*; here through the CPU emulation of the C ABI):
    GEGLU interleave with LayerNorm fold, conv tap-major layout with channel padding, fused q|k|v rows."""
    from tests import fake_lib
    from this_and_that_vdm_b200.engine import _fold_ln, _pack_conv
    with fake_lib.installed():
        w = torch.arange(8 * 32, dtype=torch.float32).reshape(8, 32) / 64
        b = torch.arange(8, dtype=torch.float32)
        gamma, beta = torch.full((32,), 2.0), torch.full((32,), 0.5)
        wi, bi, cs = _fold_ln([w], b, gamma, beta, "cpu", geglu=True)
        order = [0, 4, 1, 5, 2, 6, 3, 7]
        assert torch.equal(wi.float(), (2 * w)[order].to(torch.bfloat16).float())
        assert torch.allclose(bi, (b + 0.5 * w.sum(1))[order]) and torch.allclose(cs, wi.float().sum(1))
        q, k, v = torch.randn(4, 32), torch.randn(4, 32), torch.randn(4, 32)
        wqkv, bq, _ = _fold_ln([q, k, v], None, torch.ones(32), torch.zeros(32), "cpu")
        assert torch.equal(wqkv.float(), torch.cat([q, k, v]).to(torch.bfloat16).float()) and float(bq.abs().max()) == 0
        cw = torch.randn(5, 8, 3, 3)
        pk = _pack_conv(cw, "cpu", pad_cin=64).float().reshape(5, 3, 3, 64)
        assert torch.allclose(pk[..., :8], cw.permute(0, 2, 3, 1).to(torch.bfloat16).float())
        assert float(pk[..., 8:].abs().max()) == 0
        tw = torch.randn(4, 4, 3, 1, 1)
        tp = _pack_conv(tw, "cpu").float().reshape(4, 3, 4)
        assert torch.allclose(tp[:, 1, :], tw[:, :, 1, 0, 0].to(torch.bfloat16).float())
