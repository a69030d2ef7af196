"""CPU fp32 ORACLE of the conditioning builder (encode_clip).  *** TEST INFRASTRUCTURE ONLY ***

PARITY PINNED (towers) / restated (assembly): `encode_clip` of the reference
(svd/pipeline_stable_video_diffusion_controlnet.py:130-188; same code in svd/pipeline_stable_video_diffusion.py) calls
two `transformers` modules — `self.image_encoder(image).image_embeds` (CLIPVisionModelWithProjection, loaded at
test_code/inference.py:325-327) and `text_encoder(prompt)[0]` (CLIPTextModel, :347-348) — then concatenates
[text(77) | image(1)] tokens, applies a freshly built `nn.LayerNorm((78, 1024))` (:172-173) and stacks zeros for
classifier-free guidance (:176-186). `transformers` IS installed in this image (the reference pins 4.x, the image has
5.5; the CLIP arithmetic is unchanged), so the tower restatements below are pinned against the library itself:
tests/golden/make_clip_golden.py runs transformers' own modules on seeded tiny configs and commits their outputs, and
tests/test_clip_oracle.py checks this file against both the live library (when importable) and the golden file.

Functional style over the HF-format state dict (`vision_model.*`, `visual_projection.weight`, `text_model.*`).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _act(x: torch.Tensor, kind: str) -> torch.Tensor:
    if kind == "gelu":
        return F.gelu(x)
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(f"unsupported hidden_act {kind!r}")


def _ln(sd: SD, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def encoder_layer(sd: SD, p: str, x: torch.Tensor, heads: int, act: str, eps: float, causal: bool) -> torch.Tensor:
    """transformers CLIPEncoderLayer: x + attn(LN1(x)); x + fc2(act(fc1(LN2(x)))). Attention scale = head_dim**-0.5."""
    B, S, C = x.shape
    d = C // heads
    y = _ln(sd, p + ".layer_norm1", x, eps)
    q = _lin(sd, p + ".self_attn.q_proj", y).view(B, S, heads, d).transpose(1, 2)
    k = _lin(sd, p + ".self_attn.k_proj", y).view(B, S, heads, d).transpose(1, 2)
    v = _lin(sd, p + ".self_attn.v_proj", y).view(B, S, heads, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * d ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(S, S, dtype=torch.bool, device=x.device).triu(1), float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, C)
    x = x + _lin(sd, p + ".self_attn.out_proj", o)
    y = _ln(sd, p + ".layer_norm2", x, eps)
    return x + _lin(sd, p + ".mlp.fc2", _act(_lin(sd, p + ".mlp.fc1", y), act))


def _n_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.encoder.layers.{n}.layer_norm1.weight" in sd:
        n += 1
    return n


def vision_image_embeds(sd: SD, pixel_values: torch.Tensor, heads: int, act: str = "gelu", eps: float = 1e-5) -> torch.Tensor:
    """CLIPVisionModelWithProjection(pixel_values).image_embeds: [N, 3, H, W] -> [N, projection_dim]."""
    p = "vision_model"
    w = sd[p + ".embeddings.patch_embedding.weight"]
    x = F.conv2d(pixel_values, w, stride=w.shape[-1]).flatten(2).transpose(1, 2)  # [N, P, C]
    cls = sd[p + ".embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + sd[p + ".embeddings.position_embedding.weight"][None]
    x = _ln(sd, p + ".pre_layrnorm", x, eps)
    for i in range(_n_layers(sd, p)):
        x = encoder_layer(sd, f"{p}.encoder.layers.{i}", x, heads, act, eps, causal=False)
    pooled = _ln(sd, p + ".post_layernorm", x[:, 0], eps)
    return F.linear(pooled, sd["visual_projection.weight"])


def text_last_hidden_state(sd: SD, input_ids: torch.Tensor, heads: int, act: str = "gelu", eps: float = 1e-5) -> torch.Tensor:
    """CLIPTextModel(input_ids)[0]: [B, L] token ids -> [B, L, hidden] (causal self-attention, final_layer_norm)."""
    p = "text_model"
    L = input_ids.shape[1]
    x = sd[p + ".embeddings.token_embedding.weight"][input_ids] + sd[p + ".embeddings.position_embedding.weight"][:L][None]
    for i in range(_n_layers(sd, p)):
        x = encoder_layer(sd, f"{p}.encoder.layers.{i}", x, heads, act, eps, causal=True)
    return _ln(sd, p + ".final_layer_norm", x, eps)


def assemble(image_embeds: torch.Tensor, text_states: Optional[torch.Tensor], do_cfg: bool,
             num_videos_per_prompt: int = 1) -> torch.Tensor:
    """The tail of encode_clip (:156-186): unsqueeze + repeat, [text | image] concat, fresh LayerNorm over the whole
    (tokens, dim) slab (weight 1, bias 0, eps 1e-5), zeros stacked in FRONT for classifier-free guidance."""
    ehs = image_embeds.unsqueeze(1)
    bs, seq, _ = ehs.shape
    ehs = ehs.repeat(1, num_videos_per_prompt, 1).view(bs * num_videos_per_prompt, seq, -1)
    if text_states is not None:
        ehs = torch.cat((text_states, ehs), dim=1)
        ehs = F.layer_norm(ehs, tuple(ehs.shape[1:]), None, None, 1e-5)
    if do_cfg:
        ehs = torch.cat([torch.zeros_like(ehs), ehs])
    return ehs


def resize_with_antialiasing(image: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """_resize_with_antialiasing of the reference (svd/pipeline_stable_video_diffusion_controlnet.py:741-845):
    Gaussian pre-blur with sigma = max((factor - 1) / 2, 0.001), kernel 2*2*sigma (min 3, made odd), reflect padding,
    then bicubic interpolation with align_corners=True."""
    h, w = image.shape[-2:]
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = [int(max(2.0 * 2.0 * s, 3)) for s in sigmas]
    ks = [k + 1 if k % 2 == 0 else k for k in ks]

    def k1d(n, sigma):
        x = torch.arange(n, dtype=image.dtype) - n // 2
        g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
        return g / g.sum()

    ky, kx = k1d(ks[0], sigmas[0]), k1d(ks[1], sigmas[1])
    c = image.shape[1]
    x = F.pad(image, [ks[1] // 2, ks[1] // 2, ks[0] // 2, ks[0] // 2], mode="reflect")
    x = F.conv2d(x, kx.view(1, 1, 1, -1).expand(c, 1, 1, -1), groups=c)
    x = F.conv2d(x, ky.view(1, 1, -1, 1).expand(c, 1, -1, 1), groups=c)
    return F.interpolate(x, size=size, mode="bicubic", align_corners=True)
