"""CPU fp32 ORACLE of the This&That / SVD denoising hot path.  *** TEST INFRASTRUCTURE ONLY ***

PARITY UNPINNED: the reference ships no tests / golden vectors for this path and its arithmetic lives in the
un-vendored dependency diffusers==0.25.1 (requirements.txt:23), which is not installable here. This file restates
(a) the reference's own data flow — svd/unet_spatio_temporal_condition.py:363-536, svd/temporal_controlnet.py:455-641,
svd/diffusion_arch/unet_3d_blocks.py:1870-2396, svd/diffusion_arch/transformer_temporal.py:276-381, the Euler loops
at svd/pipeline_stable_video_diffusion_controlnet.py:582-720 and svd/pipeline_stable_video_diffusion.py:495-562 —
and (b) the diffusers 0.25.1 layer semantics listed in SURVEY.md Appendix A. It is pinned only by the self-checks
in tests/test_oracle.py (exact parameter counts, torch built-in cross-checks, zero-init == no ControlNet, the
time_context quirk, L=1 closed form, Karras/Euler known answers).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (svd/, this_and_that_vdm_b200/) must never import it.

Functional style: every function takes the diffusers-format state dict `sd` (keys as in SURVEY.md Appendix C) and a
key prefix, so the oracle shares no code with the product's module tree.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

SVD_CONFIG = dict(
    in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20),
    layers_per_block=2, cross_attention_dim=1024, addition_time_embed_dim=256, num_frames=14,
    down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
    up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3,
)


# ------------------------------------------------------------------------------------------------ A.1 / A.2
def timesteps_sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): fp32, cos first."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    arg = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1)


def linear(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def timestep_embedding(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return linear(sd, p + ".linear_2", F.silu(linear(sd, p + ".linear_1", x)))


# ------------------------------------------------------------------------------------------------ A.3 - A.7
def resnet_block_2d(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, eps: float) -> torch.Tensor:
    h = F.group_norm(x, 32, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = F.conv2d(F.silu(h), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = h + linear(sd, p + ".time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.group_norm(h, 32, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.conv2d(F.silu(h), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def temporal_resnet_block(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor, eps: float) -> torch.Tensor:
    """x [B, C, F, h, w]; temb [B, F, 1280]. GroupNorm over the 5-D tensor (stats span all frames)."""
    h = F.group_norm(x, 32, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = F.conv3d(F.silu(h), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=(1, 0, 0))
    t = linear(sd, p + ".time_emb_proj", F.silu(temb))  # [B, F, C]
    h = h + t.permute(0, 2, 1)[:, :, :, None, None]
    h = F.group_norm(h, 32, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.conv3d(F.silu(h), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=(1, 0, 0))
    return x + h


def alpha_blend(sd: SD, p: str, x_spatial: torch.Tensor, x_temporal: torch.Tensor,
                image_only_indicator: torch.Tensor) -> torch.Tensor:
    """AlphaBlender 'learned_with_images': alpha = where(indicator, 1, sigmoid(mix_factor))."""
    alpha = torch.where(image_only_indicator.bool(), torch.ones(1, 1, device=x_spatial.device),
                        torch.sigmoid(sd[p + ".mix_factor"])[..., None])
    if x_spatial.ndim == 5:
        alpha = alpha[:, None, :, None, None]
    elif x_spatial.ndim == 3:
        alpha = alpha.reshape(-1)[:, None, None]
    else:
        raise ValueError(x_spatial.ndim)
    alpha = alpha.to(x_spatial.dtype)
    return alpha * x_spatial + (1.0 - alpha) * x_temporal


def spatio_temporal_res_block(sd: SD, p: str, x: torch.Tensor, temb: torch.Tensor,
                              image_only_indicator: torch.Tensor, eps: float) -> torch.Tensor:
    B, Fr = image_only_indicator.shape
    h = resnet_block_2d(sd, p + ".spatial_res_block", x, temb, eps)
    BF, C, hh, ww = h.shape
    h5 = h[None, :].reshape(B, Fr, C, hh, ww).permute(0, 2, 1, 3, 4)
    h_mix = h5
    temb5 = temb.reshape(B, Fr, -1)
    h5 = temporal_resnet_block(sd, p + ".temporal_res_block", h5, temb5, eps)
    h5 = alpha_blend(sd, p + ".time_mixer", h_mix, h5, image_only_indicator)
    return h5.permute(0, 2, 1, 3, 4).reshape(BF, C, hh, ww)


# ------------------------------------------------------------------------------------------------ A.8
def attention(sd: SD, p: str, x: torch.Tensor, ctx: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """diffusers Attention + AttnProcessor2_0 (no mask, scale head_dim**-0.5)."""
    ctx = x if ctx is None else ctx
    q, k, v = linear(sd, p + ".to_q", x), linear(sd, p + ".to_k", ctx), linear(sd, p + ".to_v", ctx)
    Bn, Sq, inner = q.shape
    d = inner // heads
    q = q.view(Bn, Sq, heads, d).transpose(1, 2)
    k = k.view(Bn, -1, heads, d).transpose(1, 2)
    v = v.view(Bn, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    o = o.transpose(1, 2).reshape(Bn, Sq, inner)
    return linear(sd, p + ".to_out.0", o)


def feed_forward(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward(activation_fn='geglu'): proj -> chunk(hidden, gate) -> hidden * gelu_erf(gate) -> Linear."""
    hidden, gate = linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, p + ".net.2", hidden * F.gelu(gate))


def layer_norm(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def basic_transformer_block(sd: SD, p: str, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    x = x + attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), None, heads)
    x = x + attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), ctx, heads)
    x = x + feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x))
    return x


def temporal_basic_transformer_block(sd: SD, p: str, x: torch.Tensor, num_frames: int, ctx: torch.Tensor,
                                     heads: int) -> torch.Tensor:
    BF, S, C = x.shape
    B = BF // num_frames
    x = x[None, :].reshape(B, num_frames, S, C).permute(0, 2, 1, 3).reshape(B * S, num_frames, C)
    residual = x
    x = feed_forward(sd, p + ".ff_in", layer_norm(sd, p + ".norm_in", x)) + residual  # is_res (dim == inner dim)
    x = x + attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), None, heads)
    x = x + attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), ctx, heads)
    x = feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    x = x[None, :].reshape(B, S, num_frames, C).permute(0, 2, 1, 3).reshape(B * num_frames, S, C)
    return x


def transformer_spatio_temporal(sd: SD, p: str, x: torch.Tensor, ehs: torch.Tensor,
                                image_only_indicator: torch.Tensor, heads: int) -> torch.Tensor:
    """svd/diffusion_arch/transformer_temporal.py:276-381, including the time_context construction :309-319
    reproduced LITERALLY (broadcast to (S, B, L, D) then flattened S-major) — this is what makes temporal row
    r = b*S + s attend context r mod B."""
    BF, C, hh, ww = x.shape
    Fr = image_only_indicator.shape[-1]
    B = BF // Fr
    time_context = ehs
    first = time_context[None, :].reshape(B, Fr, -1, time_context.shape[-1])[:, 0]
    L = first.shape[1]
    time_context = first[None, :].broadcast_to(hh * ww, B, L, first.shape[-1])
    time_context = time_context.reshape(hh * ww * B, L, first.shape[-1])

    residual = x
    h = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = h.permute(0, 2, 3, 1).reshape(BF, hh * ww, C)
    h = linear(sd, p + ".proj_in", h)

    frames = torch.arange(Fr, device=x.device).repeat(B, 1).reshape(-1)
    t_emb = timesteps_sinusoid(frames, C).to(h.dtype)
    emb = timestep_embedding(sd, p + ".time_pos_embed", t_emb)[:, None, :]

    n_layers = 0
    while f"{p}.transformer_blocks.{n_layers}.norm1.weight" in sd:
        n_layers += 1
    for i in range(n_layers):
        h = basic_transformer_block(sd, f"{p}.transformer_blocks.{i}", h, ehs, heads)
        h_mix = h + emb
        h_mix = temporal_basic_transformer_block(sd, f"{p}.temporal_transformer_blocks.{i}", h_mix, Fr, time_context,
                                                 heads)
        h = alpha_blend(sd, p + ".time_mixer", h, h_mix, image_only_indicator)
    h = linear(sd, p + ".proj_out", h)
    h = h.reshape(BF, hh, ww, C).permute(0, 3, 1, 2).contiguous()
    return h + residual


# ------------------------------------------------------------------------------------------------ blocks
def _count(sd: SD, fmt: str) -> int:
    n = 0
    while fmt.format(n) in sd:
        n += 1
    return n


def down_block(sd: SD, p: str, x, temb, ehs, ind, heads: int, cross: bool):
    """CrossAttnDownBlockSpatioTemporal (eps 1e-6) / DownBlockSpatioTemporal (eps 1e-5)."""
    outs = []
    n = _count(sd, p + ".resnets.{}.spatial_res_block.norm1.weight")
    for j in range(n):
        x = spatio_temporal_res_block(sd, f"{p}.resnets.{j}", x, temb, ind, 1e-6 if cross else 1e-5)
        if cross:
            x = transformer_spatio_temporal(sd, f"{p}.attentions.{j}", x, ehs, ind, heads)
        outs.append(x)
    if (p + ".downsamplers.0.conv.weight") in sd:
        x = F.conv2d(x, sd[p + ".downsamplers.0.conv.weight"], sd[p + ".downsamplers.0.conv.bias"], stride=2,
                     padding=1)
        outs.append(x)
    return x, outs


def mid_block(sd: SD, p: str, x, temb, ehs, ind, heads: int):
    x = spatio_temporal_res_block(sd, p + ".resnets.0", x, temb, ind, 1e-5)
    n = _count(sd, p + ".attentions.{}.norm.weight")
    for j in range(n):
        x = transformer_spatio_temporal(sd, f"{p}.attentions.{j}", x, ehs, ind, heads)
        x = spatio_temporal_res_block(sd, f"{p}.resnets.{j + 1}", x, temb, ind, 1e-5)
    return x


def up_block(sd: SD, p: str, x, skips: List[torch.Tensor], temb, ehs, ind, heads: int, cross: bool):
    """UpBlockSpatioTemporal / CrossAttnUpBlockSpatioTemporal (both eps 1e-6): pops skips from the END."""
    n = _count(sd, p + ".resnets.{}.spatial_res_block.norm1.weight")
    for j in range(n):
        x = torch.cat([x, skips.pop()], dim=1)
        x = spatio_temporal_res_block(sd, f"{p}.resnets.{j}", x, temb, ind, 1e-6)
        if cross:
            x = transformer_spatio_temporal(sd, f"{p}.attentions.{j}", x, ehs, ind, heads)
    if (p + ".upsamplers.0.conv.weight") in sd:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        x = F.conv2d(x, sd[p + ".upsamplers.0.conv.weight"], sd[p + ".upsamplers.0.conv.bias"], padding=1)
    return x


def _embed(sd: SD, cfg: dict, timestep, added_time_ids: torch.Tensor, B: int, dtype, device) -> torch.Tensor:
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.float64 if isinstance(t, float) else torch.int64, device=device)
    elif t.ndim == 0:
        t = t[None].to(device)
    t = t.expand(B)
    C0 = cfg["block_out_channels"][0]
    emb = timestep_embedding(sd, "time_embedding", timesteps_sinusoid(t, C0).to(dtype))
    te = timesteps_sinusoid(added_time_ids.flatten(), cfg["addition_time_embed_dim"]).reshape(B, -1).to(emb.dtype)
    return emb + timestep_embedding(sd, "add_embedding", te)


def unet_forward(sd: SD, cfg: dict, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                 added_time_ids: torch.Tensor, down_block_additional_residuals: Optional[Sequence[torch.Tensor]] = None,
                 mid_block_additional_residual: Optional[torch.Tensor] = None,
                 return_intermediates: bool = False):
    """svd/unet_spatio_temporal_condition.py:363-536. sample [B, F, 8, h, w] -> [B, F, 4, h, w]."""
    B, Fr = sample.shape[:2]
    heads = cfg["num_attention_heads"]
    emb = _embed(sd, cfg, timestep, added_time_ids, B, sample.dtype, sample.device)
    x = sample.flatten(0, 1)
    emb = emb.repeat_interleave(Fr, dim=0)
    ehs = encoder_hidden_states.repeat_interleave(Fr, dim=0)
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    ind = torch.zeros(B, Fr, dtype=x.dtype, device=x.device)
    inter = {"conv_in": x}

    skips = [x]
    for i, btype in enumerate(cfg["down_block_types"]):
        x, outs = down_block(sd, f"down_blocks.{i}", x, emb, ehs, ind, heads[i], btype.startswith("CrossAttn"))
        skips += outs
        inter[f"down{i}"] = x
    if mid_block_additional_residual is not None and down_block_additional_residuals is not None:
        skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
    x = mid_block(sd, "mid_block", x, emb, ehs, ind, heads[-1])
    if mid_block_additional_residual is not None and down_block_additional_residuals is not None:
        x = x + mid_block_additional_residual
    inter["mid"] = x
    rheads = list(reversed(heads))
    for i, btype in enumerate(cfg["up_block_types"]):
        x = up_block(sd, f"up_blocks.{i}", x, skips, emb, ehs, ind, rheads[i], btype.startswith("CrossAttn"))
        inter[f"up{i}"] = x
    x = F.group_norm(x, 32, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5)
    x = F.conv2d(F.silu(x), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    x = x.reshape(B, Fr, *x.shape[1:])
    return (x, inter) if return_intermediates else x


def controlnet_forward(sd: SD, cfg: dict, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                       added_time_ids: torch.Tensor, controlnet_cond: torch.Tensor, conditioning_scale: float = 1.0,
                       guess_mode: bool = False) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """svd/temporal_controlnet.py:455-641. controlnet_cond [B*F, 4, h, w] (already VAE-encoded)."""
    B, Fr = sample.shape[:2]
    heads = cfg["num_attention_heads"]
    emb = _embed(sd, cfg, timestep, added_time_ids, B, sample.dtype, sample.device)
    x = sample.flatten(0, 1)
    emb = emb.repeat_interleave(Fr, dim=0)
    ehs = encoder_hidden_states.repeat_interleave(Fr, dim=0)
    ind = torch.zeros(B, Fr, dtype=x.dtype, device=x.device)
    x = torch.cat([x, controlnet_cond], dim=1)
    x = F.conv2d(x, sd["conv_in_concat.weight"], sd["conv_in_concat.bias"], padding=1)
    skips = [x]
    for i, btype in enumerate(cfg["down_block_types"]):
        x, outs = down_block(sd, f"down_blocks.{i}", x, emb, ehs, ind, heads[i], btype.startswith("CrossAttn"))
        skips += outs
    x = mid_block(sd, "mid_block", x, emb, ehs, ind, heads[-1])
    down = [F.conv2d(s, sd[f"controlnet_down_blocks.{i}.weight"], sd[f"controlnet_down_blocks.{i}.bias"])
            for i, s in enumerate(skips)]
    mid = F.conv2d(x, sd["controlnet_mid_block.weight"], sd["controlnet_mid_block.bias"])
    if guess_mode:
        scales = torch.logspace(-1, 0, len(down) + 1, device=x.device) * conditioning_scale
        down = [d * s for d, s in zip(down, scales)]
        mid = mid * scales[-1]
    else:
        down = [d * conditioning_scale for d in down]
        mid = mid * conditioning_scale
    return down, mid


# ------------------------------------------------------------------------------------------------ A.9 sampler
def karras_sigmas(n: int, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0) -> torch.Tensor:
    """EulerDiscreteScheduler.set_timesteps with use_karras_sigmas (SVD scheduler_config): n sigmas + final 0."""
    ramp = torch.linspace(0, 1, n, dtype=torch.float64)
    min_inv, max_inv = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (max_inv + ramp * (min_inv - max_inv)) ** rho
    return torch.cat([sig, torch.zeros(1, dtype=torch.float64)]).to(torch.float32)


def euler_timesteps(sigmas: torch.Tensor) -> torch.Tensor:
    """timestep_type 'continuous': t = 0.25 * ln(sigma)."""
    return 0.25 * torch.log(sigmas[:-1])


def init_noise_sigma(sigmas: torch.Tensor) -> float:
    return float((sigmas.max() ** 2 + 1) ** 0.5)


def euler_step(model_output: torch.Tensor, sample: torch.Tensor, sigma: float, sigma_next: float) -> torch.Tensor:
    """v-prediction Euler step in fp32 (s_churn = 0)."""
    sample = sample.float()
    x0 = model_output.float() * (-sigma / (sigma ** 2 + 1) ** 0.5) + sample / (sigma ** 2 + 1)
    d = (sample - x0) / sigma
    return sample + d * (sigma_next - sigma)


def denoise_loop(unet_sd: SD, cfg: dict, latents: torch.Tensor, image_latents: torch.Tensor,
                 encoder_hidden_states: torch.Tensor, added_time_ids: torch.Tensor, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                 controlnet_sd: Optional[SD] = None, controlnet_cfg: Optional[dict] = None,
                 controlnet_cond: Optional[torch.Tensor] = None, conditioning_scale: float = 1.0,
                 max_steps: Optional[int] = None) -> torch.Tensor:
    """The 25-step loop of both pipelines for ONE video with CFG (B = 2):
    latents [1, F, 4, h, w] (already multiplied by init_noise_sigma), image_latents [2, F, 4, h, w] (uncond row 0
    is zeros), encoder_hidden_states [2, L, 1024], added_time_ids [2, 3], controlnet_cond [F, 4, h, w]."""
    Fr = latents.shape[1]
    sigmas = karras_sigmas(num_inference_steps)
    ts = euler_timesteps(sigmas)
    guidance = torch.linspace(min_guidance_scale, max_guidance_scale, Fr)[None, :, None, None, None]
    n = num_inference_steps if max_steps is None else min(max_steps, num_inference_steps)
    for i in range(n):
        sigma, sigma_next = float(sigmas[i]), float(sigmas[i + 1])
        x = torch.cat([latents] * 2) / ((sigma ** 2 + 1) ** 0.5)
        x = torch.cat([x, image_latents], dim=2)
        down = mid = None
        if controlnet_sd is not None:
            cc = torch.cat([controlnet_cond, controlnet_cond])
            down, mid = controlnet_forward(controlnet_sd, controlnet_cfg, x, ts[i], encoder_hidden_states,
                                           added_time_ids, cc, conditioning_scale, False)
        eps = unet_forward(unet_sd, cfg, x, ts[i], encoder_hidden_states, added_time_ids, down, mid)
        eu, ec = eps.chunk(2)
        eps = eu + guidance * (ec - eu)
        latents = euler_step(eps, latents, sigma, sigma_next)
    return latents
