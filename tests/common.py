"""Shared builders for the parity tests: seeded tiny / SVD-config models and synthetic inputs."""
from __future__ import annotations

import torch

from oracle import svd_oracle as O
from svd.temporal_controlnet import ControlNetModel
from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel

TINY = dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4))
SVD = dict(block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20))


def oracle_cfg(kind: dict) -> dict:
    cfg = dict(O.SVD_CONFIG)
    cfg.update(kind)
    return cfg


def randomize_special(model, seed: int) -> None:
    """Zero-inits (GestureNet) get non-zero values and mix factors move off 0.5 so every path carries signal."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("mix_factor"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif name.startswith("controlnet_") or name.startswith("conv_in_concat"):
                fan_in = p[0].numel() if p.ndim > 1 else p.numel()
                p.copy_(torch.randn(p.shape, generator=g) * (fan_in ** -0.5 if p.ndim > 1 else 0.1))
            elif ".norm" in name or name.startswith("conv_norm_out"):
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))


def build_models(kind: dict, seed: int = 1234, controlnet: bool = True):
    torch.manual_seed(seed)
    unet = UNetSpatioTemporalConditionModel(num_frames=14, **kind).eval()
    randomize_special(unet, seed + 1)
    cn = None
    if controlnet:
        cn = ControlNetModel(**kind).eval()
        randomize_special(cn, seed + 2)
    return unet, cn


def state(model):
    return {k: v.detach().float() for k, v in model.state_dict().items()}


def make_inputs(B: int, F: int, h: int, w: int, L: int = 78, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn(B, F, 8, h, w, generator=g)
    ehs = torch.randn(B, L, 1024, generator=g)
    ehs = torch.nn.functional.layer_norm(ehs, (L, 1024))
    if B > 1:
        ehs[0] = 0.0  # CFG: unconditional half is zeros
    ati = torch.tensor([[6.0, 200.0, 0.1]] * B)
    cond = torch.randn(F, 4, h, w, generator=g)
    return sample, ehs, ati, cond


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


# ---- VAE (AutoencoderKLTemporalDecoder) builders
TINY_VAE = dict(block_out_channels=(64, 128, 128, 128), layers_per_block=1,
                down_block_types=("DownEncoderBlock2D",) * 4)
SVD_VAE = dict(block_out_channels=(128, 256, 512, 512), layers_per_block=2,
               down_block_types=("DownEncoderBlock2D",) * 4)


def build_vae(kind: dict, seed: int = 4321):
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    torch.manual_seed(seed)
    vae = AutoencoderKLTemporalDecoder(**kind).eval()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in vae.named_parameters():
            if name.endswith("mix_factor"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
            elif "norm" in name:
                if name.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
                else:
                    p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return vae


def vae_inputs(n_frames: int, lh: int, lw: int, n_images: int = 2, seed: int = 5):
    """Latents [n_frames, 4, lh, lw] (unit scale, i.e. already / scaling_factor) and images [n_images, 3, 8lh, 8lw]."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n_frames, 4, lh, lw, generator=g)
    x = torch.rand(n_images, 3, 8 * lh, 8 * lw, generator=g) * 2.0 - 1.0
    return z, x


# ---- CLIP towers (HF key names); weights are generated here from a seed so that no test needs `transformers`
TINY_CLIP_VISION = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                        image_size=56, patch_size=14, projection_dim=64, hidden_act="gelu")
TINY_CLIP_VISION_D80 = dict(hidden_size=320, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4,
                            image_size=42, patch_size=14, projection_dim=64, hidden_act="quick_gelu")  # head_dim 80 (ViT-H)
TINY_CLIP_TEXT = dict(vocab_size=300, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                      num_attention_heads=2, max_position_embeddings=77, hidden_act="gelu")


def _clip_layers(sd, prefix, cfg, g):
    C, I = cfg["hidden_size"], cfg["intermediate_size"]
    rn = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale  # noqa: E731
    for i in range(cfg["num_hidden_layers"]):
        p = f"{prefix}.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[f"{p}.self_attn.{n}.weight"] = rn(C, C, scale=C ** -0.5)
            sd[f"{p}.self_attn.{n}.bias"] = rn(C, scale=0.1)
        for n in ("layer_norm1", "layer_norm2"):
            sd[f"{p}.{n}.weight"] = 1.0 + rn(C, scale=0.1)
            sd[f"{p}.{n}.bias"] = rn(C, scale=0.1)
        sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"] = rn(I, C, scale=C ** -0.5), rn(I, scale=0.1)
        sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"] = rn(C, I, scale=I ** -0.5), rn(C, scale=0.1)


def clip_vision_sd(cfg: dict, seed: int = 77):
    g = torch.Generator().manual_seed(seed)
    C, ps = cfg["hidden_size"], cfg["patch_size"]
    n_pos = (cfg["image_size"] // ps) ** 2 + 1
    rn = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale  # noqa: E731
    sd = {"vision_model.embeddings.class_embedding": rn(C),
          "vision_model.embeddings.patch_embedding.weight": rn(C, 3, ps, ps, scale=(3 * ps * ps) ** -0.5),
          "vision_model.embeddings.position_embedding.weight": rn(n_pos, C, scale=0.3)}
    for n in ("pre_layrnorm", "post_layernorm"):
        sd[f"vision_model.{n}.weight"], sd[f"vision_model.{n}.bias"] = 1.0 + rn(C, scale=0.1), rn(C, scale=0.1)
    _clip_layers(sd, "vision_model", cfg, g)
    sd["visual_projection.weight"] = rn(cfg["projection_dim"], C, scale=C ** -0.5)
    return sd


def clip_text_sd(cfg: dict, seed: int = 78):
    g = torch.Generator().manual_seed(seed)
    C = cfg["hidden_size"]
    rn = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale  # noqa: E731
    sd = {"text_model.embeddings.token_embedding.weight": rn(cfg["vocab_size"], C, scale=0.5),
          "text_model.embeddings.position_embedding.weight": rn(cfg["max_position_embeddings"], C, scale=0.3),
          "text_model.final_layer_norm.weight": 1.0 + rn(C, scale=0.1), "text_model.final_layer_norm.bias": rn(C, scale=0.1)}
    _clip_layers(sd, "text_model", cfg, g)
    return sd


def clip_inputs(vcfg: dict, tcfg: dict, n: int = 1, seed: int = 9):
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(n, 3, vcfg["image_size"], vcfg["image_size"], generator=g)
    ids = torch.randint(0, tcfg["vocab_size"], (n, tcfg["max_position_embeddings"]), generator=g)
    return px, ids
