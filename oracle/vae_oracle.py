"""CPU fp32 ORACLE of the VAE either side of the denoising loop.  *** TEST INFRASTRUCTURE ONLY ***

PARITY UNPINNED: the reference calls `self.vae.encode(image).latent_dist.mode()`
(svd/pipeline_stable_video_diffusion_controlnet.py:199, :652) and `self.vae.decode(latents[i:i+chunk],
num_frames=...)` (:257-283) on diffusers' `AutoencoderKLTemporalDecoder` (imported at test_code/inference.py:22,
loaded at :328-330). The class lives in the un-vendored dependency diffusers==0.25.1 (requirements.txt:23), which is
not installable here, and the reference has no tests or golden vectors for it. This file restates the published
0.25.1 layer semantics:

  encoder  (diffusers `Encoder`, double_z): conv_in 3->128; 4 x DownEncoderBlock2D = 2 x ResnetBlock2D(temb=None,
           eps 1e-6, 1x1 conv_shortcut when Cin != Cout) + [Downsample2D(padding=0): F.pad(x,(0,1,0,1)) -> Conv2d(3, s2)]
           (all but the last); UNetMidBlock2D = ResnetBlock2D -> Attention(1 head of 512, GroupNorm(32, 1e-6), bias,
           residual) -> ResnetBlock2D; GroupNorm(32, 1e-6) -> SiLU -> conv_out 512->8; quant_conv 1x1;
           DiagonalGaussianDistribution.mode() = first 4 channels.
  decoder  (`TemporalDecoder`): conv_in 4->512; MidBlockTemporalDecoder = SpatioTemporalResBlock -> Attention ->
           SpatioTemporalResBlock; 4 x UpBlockTemporalDecoder = 3 x SpatioTemporalResBlock + [nearest x2 -> Conv2d 3x3];
           GroupNorm(32, 1e-6) -> SiLU -> conv_out 128->3 -> time_conv_out Conv3d(3, 3, (3,1,1), pad (1,0,0)).
           SpatioTemporalResBlock here: temb_channels=None, eps 1e-6 (spatial) / temporal_eps 1e-5, AlphaBlender
           merge_strategy="learned", switch_spatial_to_temporal_mix=True: a = 1 - sigmoid(mix_factor),
           out = a * spatial + (1 - a) * temporal.

It is pinned only by the self-checks of tests/test_vae_oracle.py (torch module cross-checks, parameter counts, key
scheme, single-frame / time-mixer closed forms). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module; the product never does.

Functional style over the diffusers-format state dict (keys `encoder.*`, `quant_conv.*`, `decoder.*`).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

SVD_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                      layers_per_block=2, scaling_factor=0.18215)


def _conv(sd: SD, p: str, x: torch.Tensor, **kw) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], **kw)


def _gn(sd: SD, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def resnet_block_2d(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """ResnetBlock2D with temb_channels=None (no time_emb_proj), output_scale_factor 1."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, eps)), padding=1)
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, eps)), padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def temporal_resnet_block(sd: SD, p: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """TemporalResnetBlock with temb_channels=None on x [B, C, F, h, w] (GroupNorm statistics span all frames)."""
    h = F.group_norm(x, 32, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = F.conv3d(F.silu(h), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=(1, 0, 0))
    h = F.group_norm(h, 32, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.conv3d(F.silu(h), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=(1, 0, 0))
    return x + h


def spatio_temporal_res_block(sd: SD, p: str, x: torch.Tensor, num_frames: int) -> torch.Tensor:
    h = resnet_block_2d(sd, p + ".spatial_res_block", x, 1e-6)
    BF, C, hh, ww = h.shape
    B = BF // num_frames
    h5 = h[None, :].reshape(B, num_frames, C, hh, ww).permute(0, 2, 1, 3, 4)
    t5 = temporal_resnet_block(sd, p + ".temporal_res_block", h5, 1e-5)
    alpha = 1.0 - torch.sigmoid(sd[p + ".time_mixer.mix_factor"]).to(h.dtype)  # "learned" + switch_spatial_to_temporal_mix
    out = alpha * h5 + (1.0 - alpha) * t5
    return out.permute(0, 2, 1, 3, 4).reshape(BF, C, hh, ww)


def attention_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers Attention on a 4-D input: GroupNorm(32, 1e-6) over tokens, one head of C dims, biased q/k/v/out,
    residual_connection=True, rescale_output_factor 1."""
    N, C, hh, ww = x.shape
    t = x.view(N, C, hh * ww)
    y = F.group_norm(t, 32, sd[p + ".group_norm.weight"], sd[p + ".group_norm.bias"], 1e-6).transpose(1, 2)
    q = F.linear(y, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(y, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(y, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(N, C, hh, ww) + x


def _count(sd: SD, fmt: str) -> int:
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def encode(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """vae.encode(x).latent_dist.mode(): x [N, 3, H, W] in [-1, 1] -> [N, 4, H/8, W/8] (NOT multiplied by the scaling
    factor — the reference pipelines do not, svd/pipeline_stable_video_diffusion_controlnet.py:199)."""
    h = _conv(sd, "encoder.conv_in", x, padding=1)
    n_blocks = _count(sd, "encoder.down_blocks.{}.resnets.0.norm1.weight")
    for i in range(n_blocks):
        p = f"encoder.down_blocks.{i}"
        for j in range(_count(sd, p + ".resnets.{}.norm1.weight")):
            h = resnet_block_2d(sd, f"{p}.resnets.{j}", h)
        if (p + ".downsamplers.0.conv.weight") in sd:
            h = _conv(sd, p + ".downsamplers.0.conv", F.pad(h, (0, 1, 0, 1)), stride=2)
    h = resnet_block_2d(sd, "encoder.mid_block.resnets.0", h)
    h = attention_block(sd, "encoder.mid_block.attentions.0", h)
    h = resnet_block_2d(sd, "encoder.mid_block.resnets.1", h)
    h = _conv(sd, "encoder.conv_out", F.silu(_gn(sd, "encoder.conv_norm_out", h, 1e-6)), padding=1)
    moments = _conv(sd, "quant_conv", h)
    return moments[:, : moments.shape[1] // 2]


def decode(sd: SD, z: torch.Tensor, num_frames: int) -> torch.Tensor:
    """vae.decode(z, num_frames).sample: z [B*num_frames, 4, h, w] (already divided by the scaling factor by the
    caller, svd/pipeline_stable_video_diffusion_controlnet.py:261) -> [B*num_frames, 3, 8h, 8w]."""
    h = _conv(sd, "decoder.conv_in", z, padding=1)
    h = spatio_temporal_res_block(sd, "decoder.mid_block.resnets.0", h, num_frames)
    n_res = _count(sd, "decoder.mid_block.resnets.{}.spatial_res_block.norm1.weight")
    for j in range(1, n_res):
        h = attention_block(sd, f"decoder.mid_block.attentions.{j - 1}", h)
        h = spatio_temporal_res_block(sd, f"decoder.mid_block.resnets.{j}", h, num_frames)
    for i in range(_count(sd, "decoder.up_blocks.{}.resnets.0.spatial_res_block.norm1.weight")):
        p = f"decoder.up_blocks.{i}"
        for j in range(_count(sd, p + ".resnets.{}.spatial_res_block.norm1.weight")):
            h = spatio_temporal_res_block(sd, f"{p}.resnets.{j}", h, num_frames)
        if (p + ".upsamplers.0.conv.weight") in sd:
            h = _conv(sd, p + ".upsamplers.0.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"), padding=1)
    h = _conv(sd, "decoder.conv_out", F.silu(_gn(sd, "decoder.conv_norm_out", h, 1e-6)), padding=1)
    BF, C, hh, ww = h.shape
    B = BF // num_frames
    h5 = h[None, :].reshape(B, num_frames, C, hh, ww).permute(0, 2, 1, 3, 4)
    h5 = F.conv3d(h5, sd["decoder.time_conv_out.weight"], sd["decoder.time_conv_out.bias"], padding=(1, 0, 0))
    return h5.permute(0, 2, 1, 3, 4).reshape(BF, C, hh, ww)


def decode_latents(sd: SD, latents: torch.Tensor, num_frames: int, decode_chunk_size: int = 14,
                   scaling_factor: float = 0.18215) -> torch.Tensor:
    """decode_latents of the pipelines (svd/pipeline_stable_video_diffusion_controlnet.py:257-283): latents
    [B, F, 4, h, w] -> fp32 [B, 3, F, H, W]; every chunk is decoded as ONE video of `chunk` frames."""
    lat = latents.flatten(0, 1) / scaling_factor
    frames = []
    for i in range(0, lat.shape[0], decode_chunk_size):
        chunk = lat[i:i + decode_chunk_size]
        frames.append(decode(sd, chunk, chunk.shape[0]))
    frames = torch.cat(frames, 0)
    return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
