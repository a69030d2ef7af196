"""Deterministic weights / inputs shared by the reference-pinning golden generator and the tests that consume it.

TEST INFRASTRUCTURE. Stand-alone on purpose (imports only torch): `tests/golden/make_reference_golden.py` runs in a
process whose `svd` package is the REFERENCE's (/root/reference/svd), so it must not pull in this repo's `svd` or
`oracle`. Weights are a pure function of (parameter name, shape, seed) — no file holds them, every consumer regenerates
the same bits from the key list of whichever implementation it drives (reference modules, oracle, CUDA engine).
"""
from __future__ import annotations

import zlib
from typing import Dict, Iterable, Tuple

import torch

CASES = {
    # name: (config kwargs common to UNet / ControlNet, B, F, h, w)
    "tiny": (dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4)), 2, 14, 16, 24),
    "tiny_b1": (dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4)), 1, 14, 16, 24),
    "tiny_2layers": (dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                          transformer_layers_per_block=2), 2, 6, 8, 16),
    "svd": (dict(block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20)), 2, 14, 16, 24),
    "svd_hd128": (dict(block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 10, 20)), 2, 4, 8, 8),
}


def _gen(name: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def fill_value(name: str, shape: Tuple[int, ...], seed: int) -> torch.Tensor:
    """fp32 tensor for parameter `name`: weights ~ N(0, 1/fan_in), biases ~ 0.1 N(0,1), norm scales 1 + 0.1 N(0,1),
    mix factors 0.5 N(0,1). ControlNet zero-inits get the same treatment (non-zero, so every path carries signal)."""
    g = _gen(name, seed)
    r = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if name.endswith("mix_factor"):
        return 0.5 * r
    is_norm = ".norm" in name or name.startswith("conv_norm_out") or name.endswith("norm.weight") or \
        name.endswith("norm.bias")
    if name.endswith(".weight") and len(shape) > 1:
        fan_in = 1
        for d in shape[1:]:
            fan_in *= d
        return r * fan_in ** -0.5
    if name.endswith(".weight"):
        return 1.0 + 0.1 * r if is_norm else r
    return 0.1 * r


def fill_state_dict(named_shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int) -> Dict[str, torch.Tensor]:
    return {k: fill_value(k, tuple(s), seed) for k, s in named_shapes}


def make_inputs(B: int, F: int, h: int, w: int, L: int = 78, seed: int = 0):
    """sample [B,F,8,h,w], ehs [B,L,1024] (row 0 zeros when B > 1: the CFG uncond half), added_time_ids [B,3],
    controlnet_cond [F,4,h,w] (VAE-encoded gesture frames of one video)."""
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn(B, F, 8, h, w, generator=g)
    ehs = torch.randn(B, L, 1024, generator=g)
    ehs = torch.nn.functional.layer_norm(ehs, (L, 1024))
    if B > 1:
        ehs[0] = 0.0
    ati = torch.tensor([[6.0, 200.0, 0.1]] * B)
    cond = torch.randn(F, 4, h, w, generator=g)
    return sample, ehs, ati, cond


TIMESTEP = 1.0977  # 0.25 * ln(sigma) for a mid-schedule sigma (~80)


def fingerprint(t: torch.Tensor) -> torch.Tensor:
    """Small summary of a big residual tensor: [sum, abs-sum, sum of squares] in fp64 + a 256-value strided sample."""
    t64 = t.detach().double().reshape(-1)
    stride = max(1, t64.numel() // 256)
    return torch.cat([torch.stack([t64.sum(), t64.abs().sum(), (t64 * t64).sum()]), t64[::stride][:256]])
