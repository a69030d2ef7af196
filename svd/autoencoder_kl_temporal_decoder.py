"""AutoencoderKLTemporalDecoder — drop-in for the diffusers class the reference loads as its VAE
(`from diffusers import AutoencoderKLTemporalDecoder`, test_code/inference.py:22, loaded at :328-330 and handed to the
pipelines as `vae=`, :171-180) and calls either side of the denoising loop:

  * `vae.encode(image).latent_dist.mode()`          svd/pipeline_stable_video_diffusion_controlnet.py:199 (first frame)
                                                    and :652 (gesture condition frames)
  * `vae.decode(latents[i:i+chunk], num_frames=n).sample`   :257-283 (decode_latents, 8-frame chunks by default)
  * `.config.scaling_factor`, `.config.block_out_channels`, `.dtype`, `.to()`, `from_pretrained(path, subfolder="vae")`

Constructor kwargs, `.config`, state-dict keys (`encoder.*`, `quant_conv.*`, `decoder.*` — the names of the published
SVD `vae/diffusion_pytorch_model.safetensors`) and the method signatures follow diffusers 0.25.1. The module tree
below only OWNS the parameters; the arithmetic runs on the sm_100a kernels of libttvdm_sm100.so through
this_and_that_vdm_b200.vae_engine (bf16 storage, fp32 accumulation). There is no eager / CPU implementation: calling
encode() / decode() on a module that is not on a CUDA sm_100 device raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import nn

from .diffusion_arch.primitives import AlphaBlender, Downsample2D, Upsample2D, _NoForward
from .modeling_utils import ModelBase, register_to_config


# ---- parameter containers (diffusers key names; no forward)
class _ResnetBlock2D(_NoForward):
    """ResnetBlock2D(temb_channels=None): norm1, conv1, norm2, conv2 (+ 1x1 conv_shortcut when Cin != Cout)."""

    def __init__(self, in_channels: int, out_channels: int, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class _TemporalResnetBlock(_NoForward):
    def __init__(self, channels: int, eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv1 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))
        self.norm2 = nn.GroupNorm(32, channels, eps=eps, affine=True)
        self.conv2 = nn.Conv3d(channels, channels, (3, 1, 1), padding=(1, 0, 0))


class _SpatioTemporalResBlock(_NoForward):
    """SpatioTemporalResBlock(temb_channels=None, eps=1e-6, temporal_eps=1e-5, merge_factor=0.0,
    merge_strategy="learned", switch_spatial_to_temporal_mix=True) — the decoder's flavour."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.spatial_res_block = _ResnetBlock2D(in_channels, out_channels, 1e-6)
        self.temporal_res_block = _TemporalResnetBlock(out_channels, 1e-5)
        self.time_mixer = AlphaBlender(alpha=0.0, merge_strategy="learned")


class _Attention(_NoForward):
    """Attention(query_dim=C, heads=1, dim_head=C, eps=1e-6, norm_num_groups=32, bias=True, residual_connection=True)."""

    def __init__(self, channels: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, channels, eps=1e-6, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])


class _DownEncoderBlock2D(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList(
            [_ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=0)]) if add_downsample else None


class _UNetMidBlock2D(_NoForward):
    def __init__(self, channels: int):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(channels)])
        self.resnets = nn.ModuleList([_ResnetBlock2D(channels, channels), _ResnetBlock2D(channels, channels)])


class Encoder(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, block_out_channels: Tuple[int, ...], layers_per_block: int,
                 double_z: bool = True):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        blocks, c = [], block_out_channels[0]
        for i, co in enumerate(block_out_channels):
            blocks.append(_DownEncoderBlock2D(c, co, layers_per_block, i != len(block_out_channels) - 1))
            c = co
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = _UNetMidBlock2D(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * out_channels if double_z else out_channels, 3, padding=1)


class _MidBlockTemporalDecoder(_NoForward):
    def __init__(self, channels: int, num_layers: int):
        super().__init__()
        self.attentions = nn.ModuleList([_Attention(channels)])
        self.resnets = nn.ModuleList([_SpatioTemporalResBlock(channels, channels) for _ in range(num_layers)])


class _UpBlockTemporalDecoder(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, num_layers: int, add_upsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList(
            [_SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class TemporalDecoder(_NoForward):
    def __init__(self, in_channels: int = 4, out_channels: int = 3,
                 block_out_channels: Tuple[int, ...] = (128, 256, 512, 512), layers_per_block: int = 2):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[-1], 3, padding=1)
        self.mid_block = _MidBlockTemporalDecoder(block_out_channels[-1], layers_per_block)
        rev = list(reversed(block_out_channels))
        blocks, c = [], rev[0]
        for i, co in enumerate(rev):
            blocks.append(_UpBlockTemporalDecoder(c, co, layers_per_block + 1, i != len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, block_out_channels[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))


# ---- outputs (diffusers names)
class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


@dataclass
class AutoencoderKLOutput:
    latent_dist: DiagonalGaussianDistribution = None


@dataclass
class DecoderOutput:
    sample: torch.FloatTensor = None


class AutoencoderKLTemporalDecoder(ModelBase):
    _supports_gradient_checkpointing = True
    # the pipelines skip diffusers' fp16 -> fp32 "force_upcast" dance for this class: the engine stores bf16 and
    # accumulates in fp32 whatever the module dtype is, and every .to() would only trigger a weight re-pack
    _ttvdm_native = True

    @register_to_config
    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        down_block_types: Tuple[str] = ("DownEncoderBlock2D",),
        block_out_channels: Tuple[int] = (64,),
        layers_per_block: int = 1,
        latent_channels: int = 4,
        sample_size: int = 32,
        scaling_factor: float = 0.18215,
        force_upcast: float = True,
    ):
        super().__init__()
        if len(down_block_types) != len(block_out_channels):
            raise ValueError(
                f"Must provide the same number of `down_block_types` as `block_out_channels`. `down_block_types`: "
                f"{down_block_types}. `block_out_channels`: {block_out_channels}.")
        for t in down_block_types:
            if t != "DownEncoderBlock2D":
                raise ValueError(f"{t} does not exist.")
        self.encoder = Encoder(in_channels, latent_channels, tuple(block_out_channels), layers_per_block, double_z=True)
        self.decoder = TemporalDecoder(latent_channels, out_channels, tuple(block_out_channels), layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self._engine = None

    def _get_engine(self):
        from this_and_that_vdm_b200.vae_engine import VaeEngine
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                "AutoencoderKLTemporalDecoder runs only on a CUDA sm_100 device (hand-written kernels in "
                "libttvdm_sm100.so); move the model with .to('cuda') — there is no CPU / eager fallback")
        if self._engine is None or self._engine.device != dev:
            self._engine = VaeEngine(self)
        return self._engine

    def refresh_engine(self) -> None:
        """Call after changing weights in place so the packed bf16 kernel-layout copies are rebuilt."""
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    @torch.no_grad()
    def encode(self, x: torch.FloatTensor, return_dict: bool = True):
        moments = self._get_engine().encode(x)
        posterior = DiagonalGaussianDistribution(moments)
        if not return_dict:
            return (posterior,)
        return AutoencoderKLOutput(latent_dist=posterior)

    @torch.no_grad()
    def decode(self, z: torch.FloatTensor, num_frames: int, return_dict: bool = True):
        if z.shape[0] % num_frames != 0:
            raise ValueError(f"z has {z.shape[0]} frames, not a multiple of num_frames={num_frames}")
        decoded = self._get_engine().decode(z, num_frames)
        if not return_dict:
            return (decoded,)
        return DecoderOutput(sample=decoded)

    def forward(self, sample: torch.FloatTensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None, num_frames: int = 1):
        posterior = self.encode(sample).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        dec = self.decode(z, num_frames=num_frames).sample
        if not return_dict:
            return (dec,)
        return DecoderOutput(sample=dec)
