"""Seeded synthetic inputs of the benchmark — shared by bench.py's arms; imports nothing from
oracle/ or the product so both see identical bits."""
import torch

FRAMES = 14
INIT_NOISE_SIGMA = (700.0 ** 2 + 1.0) ** 0.5


def make(h: int, w: int, n_videos: int, seed: int = 0, L: int = 78):
    g = torch.Generator(device="cpu").manual_seed(seed)
    N = n_videos
    ehs = torch.nn.functional.layer_norm(torch.randn(N, L, 1024, generator=g), (L, 1024))
    img = torch.randn(N, 4, h, w, generator=g)
    return {
        "encoder_hidden_states": torch.cat([torch.zeros_like(ehs), ehs]),      # CFG: uncond rows first
        "image_latents": torch.cat([torch.zeros_like(img), img]),
        "added_time_ids": torch.tensor([[6.0, 200.0, 0.1]] * (2 * N)),
        "controlnet_cond": torch.randn(N, FRAMES, 4, h, w, generator=g),
        "latents": torch.randn(N, FRAMES, 4, h, w, generator=g) * INIT_NOISE_SIGMA,
    }
