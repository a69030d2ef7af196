"""Shared end-to-end case: the VGL pipeline called the way test_code/inference.py:246-270 calls it — PIL first frame,
token ids, numpy gesture condition, CLIP towers, VAE, UNet, GestureNet — and the same computation composed from the
three oracles. Used on CPU through tests/fake_lib.py (tests/test_pipeline_host_logic.py) and on the B200
(tests/test_pipeline_gpu.py)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import clip_oracle as CO
from oracle import svd_oracle as O
from oracle import vae_oracle as VO
from tests.common import (TINY, TINY_VAE, build_models, build_vae, clip_text_sd, clip_vision_sd, oracle_cfg, state)

VISION = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2, image_size=224,
              patch_size=32, projection_dim=1024, hidden_act="gelu")
TEXT = dict(vocab_size=300, hidden_size=1024, intermediate_size=256, num_hidden_layers=1, num_attention_heads=16,
            max_position_embeddings=77, hidden_act="gelu")
H, W, FRAMES, STEPS = 64, 128, 14, 2
MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)


def build(device="cpu"):
    from svd.clip_towers import CLIPTextModel, CLIPVisionModelWithProjection
    unet, cn = build_models(TINY)
    vae = build_vae(TINY_VAE)
    vis, txt = CLIPVisionModelWithProjection(VISION), CLIPTextModel(TEXT)
    vsd, tsd = clip_vision_sd(VISION), clip_text_sd(TEXT)
    vis.load_state_dict(vsd)
    txt.load_state_dict(tsd)
    sds = dict(unet=state(unet), cn=state(cn), vae=state(vae), vis=vsd, txt=tsd)
    mods = dict(unet=unet, cn=cn, vae=vae, vis=vis, txt=txt)
    for m in mods.values():
        m.to(device)
    return mods, sds


def inputs():
    import PIL.Image
    g = torch.Generator().manual_seed(21)
    arr = (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8).numpy()
    image = PIL.Image.fromarray(arr)
    ids = torch.randint(0, TEXT["vocab_size"], (1, 77), generator=g)
    cond = np.zeros((FRAMES, 3, H, W), dtype=np.float32)  # the rasteriser's output: 2 frames with a point, 12 zero frames
    cond[0] = torch.rand(3, H, W, generator=g).numpy()
    cond[-1] = torch.rand(3, H, W, generator=g).numpy()
    return image, ids, cond


def run_pipeline(mods, device, output_type="pt", seed=5):
    from svd.pipeline_stable_video_diffusion_controlnet import StableVideoDiffusionControlNetPipeline
    image, ids, cond = inputs()
    pipe = StableVideoDiffusionControlNetPipeline.from_pretrained(None, vae=mods["vae"], image_encoder=mods["vis"],
                                                                  unet=mods["unet"])
    pipe.to(device)
    pipe.set_progress_bar_config(disable=True)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = pipe(image, cond, controlnet=mods["cn"], prompt=ids.to(device), use_text=True, text_encoder=mods["txt"],
               height=H, width=W, num_frames=FRAMES, num_inference_steps=STEPS, decode_chunk_size=8, fps=7,
               motion_bucket_id=200, noise_aug_strength=0.02, generator=gen, guess_mode=False, output_type=output_type)
    return out.frames


def run_oracle(sds, seed=5, latent_dtype=torch.float32):
    """Same computation from oracle/: CLIP towers + assembly, VAE encode (first frame + gesture frames), 2 Euler steps of
    UNet + GestureNet with CFG, chunked VAE decode. Consumes the generator in the pipeline's order."""
    from svd.pipeline_common import randn_tensor
    image, ids, cond = inputs()
    cfg = oracle_cfg(TINY)
    px = torch.from_numpy(np.array(image).astype(np.float32) / 255.0).permute(2, 0, 1)[None]
    clip_in = (CO.resize_with_antialiasing(px * 2 - 1, (224, 224)) + 1) / 2
    emb = CO.vision_image_embeds(sds["vis"], (clip_in - MEAN) / STD, VISION["num_attention_heads"], "gelu")
    text = CO.text_last_hidden_state(sds["txt"], ids, TEXT["num_attention_heads"], "gelu")
    ehs = CO.assemble(emb, text, do_cfg=True)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    img = px * 2 - 1  # the PIL image already has the target size: VaeImageProcessor's resize is the identity
    img = img + 0.02 * randn_tensor(img.shape, generator=gen, dtype=img.dtype)
    lat_img = VO.encode(sds["vae"], img)
    img_lat = torch.cat([torch.zeros_like(lat_img), lat_img])[:, None].repeat(1, FRAMES, 1, 1, 1)
    cond_lat = VO.encode(sds["vae"], torch.from_numpy(cond).to(torch.float16).float())
    ati = torch.tensor([[6.0, 200.0, 0.02]] * 2)
    sig = O.karras_sigmas(STEPS)
    # the pipelines draw the initial noise in the conditioning's dtype (fp16 when the modules are fp16)
    lat = randn_tensor((1, FRAMES, 4, H // 8, W // 8), generator=gen, dtype=latent_dtype).float() * O.init_noise_sigma(sig)
    lat = O.denoise_loop(sds["unet"], cfg, lat, img_lat, ehs, ati, STEPS, 1.0, 3.0, sds["cn"], cfg, cond_lat, 1.0)
    return VO.decode_latents(sds["vae"], lat, FRAMES, decode_chunk_size=8), lat


def run_vl_pipeline(mods, device, output_type="pt", seed=6):
    """VL (UNet only) the way test_code/inference.py:108-145 calls it with use_text=False: the context is the single
    CLIP image token (L = 1, no LayerNorm)."""
    from svd.pipeline_stable_video_diffusion import StableVideoDiffusionPipeline
    image, _, _ = inputs()
    pipe = StableVideoDiffusionPipeline.from_pretrained(None, vae=mods["vae"], image_encoder=mods["vis"],
                                                        unet=mods["unet"])
    pipe.to(device)
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = pipe(image, prompt=None, use_text=False, text_encoder=None, height=H, width=W, num_frames=FRAMES,
               num_inference_steps=STEPS, decode_chunk_size=14, fps=7, motion_bucket_id=127, noise_aug_strength=0.02,
               generator=gen, output_type=output_type)
    return out.frames


def run_vl_oracle(sds, seed=6):
    from svd.pipeline_common import randn_tensor
    image, _, _ = inputs()
    cfg = oracle_cfg(TINY)
    px = torch.from_numpy(np.array(image).astype(np.float32) / 255.0).permute(2, 0, 1)[None]
    clip_in = (CO.resize_with_antialiasing(px * 2 - 1, (224, 224)) + 1) / 2
    emb = CO.vision_image_embeds(sds["vis"], (clip_in - MEAN) / STD, VISION["num_attention_heads"], "gelu")
    ehs = CO.assemble(emb, None, do_cfg=True)  # [2, 1, 1024]
    gen = torch.Generator(device="cpu").manual_seed(seed)
    img = px * 2 - 1
    img = img + 0.02 * randn_tensor(img.shape, generator=gen, dtype=img.dtype)
    lat_img = VO.encode(sds["vae"], img)
    img_lat = torch.cat([torch.zeros_like(lat_img), lat_img])[:, None].repeat(1, FRAMES, 1, 1, 1)
    ati = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    sig = O.karras_sigmas(STEPS)
    lat = randn_tensor((1, FRAMES, 4, H // 8, W // 8), generator=gen, dtype=torch.float32) * O.init_noise_sigma(sig)
    lat = O.denoise_loop(sds["unet"], cfg, lat, img_lat, ehs, ati, STEPS, 1.0, 3.0)
    return VO.decode_latents(sds["vae"], lat, FRAMES, decode_chunk_size=14)
