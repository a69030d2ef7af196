"""ControlNetModel ("GestureNet") — drop-in for svd/temporal_controlnet.py of the reference.

Encoder half of the SVD UNet with a zero-initialised 12-channel `conv_in_concat` (:203-205), 12 + 1 zero 1x1
convolutions over the skip tensors (:253-297) and `conditioning_scale` (:616-633). Same constructor, `from_unet`
(:311-339), state-dict keys and `forward` signature (:455-469). forward() executes on sm_100a kernels via
this_and_that_vdm_b200.engine; no eager fallback.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
from torch import nn

from .diffusion_arch.primitives import TimestepEmbedding, Timesteps
from .diffusion_arch.unet_3d_blocks import UNetMidBlockSpatioTemporal, get_down_block
from .modeling_utils import ModelBase, register_to_config
from .unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel


@dataclass
class ControlNetOutput:
    down_block_res_samples: Tuple[torch.Tensor] = None
    mid_block_res_sample: torch.Tensor = None


def zero_module(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class ControlNetModel(ModelBase):
    _supports_gradient_checkpointing = True

    @register_to_config
    def __init__(
        self,
        in_channels: int = 8,
        conditioning_channels: int = 3,
        flip_sin_to_cos: bool = True,
        freq_shift: int = 0,
        down_block_types: Tuple[str, ...] = ("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                                             "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
        mid_block_type: Optional[str] = "UNetMidBlockSpatioTemporal",
        block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280),
        addition_time_embed_dim: int = 256,
        layers_per_block: int = 2,
        act_fn: str = "silu",
        cross_attention_dim: int = 1024,
        projection_class_embeddings_input_dim: Optional[int] = 768,
        conditioning_embedding_out_channels: Optional[Tuple[int, ...]] = (16, 32, 96, 256),
        transformer_layers_per_block: Union[int, Tuple[int], Tuple[Tuple]] = 1,
        num_attention_heads: Union[int, Tuple[int]] = (5, 10, 20, 20),
        encoder_hid_dim: Optional[int] = None,
        encoder_hid_dim_type: Optional[str] = None,
        controlnet_conditioning_channel_order="rgb",
    ):
        super().__init__()
        self.controlnet_conditioning_channel_order = controlnet_conditioning_channel_order
        n = len(down_block_types)
        if len(block_out_channels) != n:
            raise ValueError(
                f"Must provide the same number of `block_out_channels` as `down_block_types`. `block_out_channels`: "
                f"{block_out_channels}. `down_block_types`: {down_block_types}.")
        if not isinstance(num_attention_heads, int) and len(num_attention_heads) != n:
            raise ValueError(
                f"Must provide the same number of `num_attention_heads` as `down_block_types`. "
                f"`num_attention_heads`: {num_attention_heads}. `down_block_types`: {down_block_types}.")
        if isinstance(cross_attention_dim, list) and len(cross_attention_dim) != n:
            raise ValueError(
                f"Must provide the same number of `cross_attention_dim` as `down_block_types`. "
                f"`cross_attention_dim`: {cross_attention_dim}. `down_block_types`: {down_block_types}.")
        if not isinstance(layers_per_block, int) and len(layers_per_block) != n:
            raise ValueError(
                f"Must provide the same number of `layers_per_block` as `down_block_types`. `layers_per_block`: "
                f"{layers_per_block}. `down_block_types`: {down_block_types}.")
        if encoder_hid_dim is None and encoder_hid_dim_type is not None:
            raise ValueError(
                f"`encoder_hid_dim` has to be defined when `encoder_hid_dim_type` is set to {encoder_hid_dim_type}.")

        # 8 latent + 4 VAE-encoded gesture channels, zero-initialised
        self.conv_in_concat = zero_module(nn.Conv2d(12, block_out_channels[0], kernel_size=3, padding=1))
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim, act_fn=act_fn)
        self.add_time_proj = Timesteps(addition_time_embed_dim, True, downscale_freq_shift=0)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, time_embed_dim)

        if isinstance(num_attention_heads, int):
            num_attention_heads = (num_attention_heads,) * n
        if isinstance(cross_attention_dim, int):
            cross_attention_dim = (cross_attention_dim,) * n
        if isinstance(layers_per_block, int):
            layers_per_block = [layers_per_block] * n
        if isinstance(transformer_layers_per_block, int):
            transformer_layers_per_block = [transformer_layers_per_block] * n

        self.down_blocks = nn.ModuleList([])
        self.controlnet_down_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        self.controlnet_down_blocks.append(zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))
        for i, down_block_type in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            is_final = i == n - 1
            self.down_blocks.append(get_down_block(
                down_block_type, num_layers=layers_per_block[i],
                transformer_layers_per_block=transformer_layers_per_block[i], in_channels=input_channel,
                out_channels=output_channel, temb_channels=time_embed_dim, add_downsample=not is_final,
                resnet_eps=1e-5, cross_attention_dim=cross_attention_dim[i],
                num_attention_heads=num_attention_heads[i], resnet_act_fn="silu"))
            for _ in range(layers_per_block[0]):
                self.controlnet_down_blocks.append(
                    zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))
            if not is_final:
                self.controlnet_down_blocks.append(
                    zero_module(nn.Conv2d(output_channel, output_channel, kernel_size=1)))

        mid_ch = block_out_channels[-1]
        self.controlnet_mid_block = zero_module(nn.Conv2d(mid_ch, mid_ch, kernel_size=1))
        if mid_block_type == "UNetMidBlockSpatioTemporal":
            self.mid_block = UNetMidBlockSpatioTemporal(
                mid_ch, temb_channels=time_embed_dim, transformer_layers_per_block=transformer_layers_per_block[-1],
                cross_attention_dim=cross_attention_dim[-1], num_attention_heads=num_attention_heads[-1])
        else:
            raise ValueError(f"unknown mid_block_type : {mid_block_type}")
        self._engine = None

    @classmethod
    def from_unet(cls, unet: UNetSpatioTemporalConditionModel, conditioning_channels: int = 3,
                  load_weights_from_unet: bool = True):
        """Built from class DEFAULTS (not unet.config), like the reference (:329); copies time embeddings, down and
        mid blocks — conv_in_concat and the zero convs stay zero."""
        controlnet = cls(conditioning_channels=conditioning_channels)
        if load_weights_from_unet:
            controlnet.time_proj.load_state_dict(unet.time_proj.state_dict())
            controlnet.time_embedding.load_state_dict(unet.time_embedding.state_dict())
            controlnet.down_blocks.load_state_dict(unet.down_blocks.state_dict())
            controlnet.mid_block.load_state_dict(unet.mid_block.state_dict())
        return controlnet

    def _get_engine(self):
        from this_and_that_vdm_b200.engine import DenoiserEngine
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                "ControlNetModel.forward runs only on a CUDA sm_100 device (hand-written kernels in "
                "libttvdm_sm100.so); move the model with .to('cuda') — there is no CPU / eager fallback")
        if self._engine is None or self._engine.device != dev:
            self._engine = DenoiserEngine(self, kind="controlnet")
        return self._engine

    def refresh_engine(self) -> None:
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    @torch.no_grad()
    def forward(
        self,
        sample: torch.FloatTensor,
        timestep: Union[torch.Tensor, float, int],
        encoder_hidden_states: torch.Tensor,
        added_time_ids: torch.Tensor,
        added_positions: torch.Tensor = None,
        controlnet_cond: torch.FloatTensor = None,
        conditioning_scale: float = 1.0,
        inner_conditioning_scale: float = 1.0,
        timestep_cond: Optional[torch.Tensor] = None,
        attention_mask: Optional[torch.Tensor] = None,
        guess_mode: bool = False,
        return_dict: bool = True,
    ) -> Union[ControlNetOutput, Tuple]:
        """`added_positions`, `inner_conditioning_scale`, `timestep_cond` and `attention_mask` are accepted and
        ignored exactly as in the reference (:461-466, :522-524). Returns 12 residuals + mid residual in the
        reference's NCHW layout ([B*F, C, h_l, w_l])."""
        down, mid = self._get_engine().controlnet_forward(
            sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond, conditioning_scale, guess_mode)
        if not return_dict:
            return (down, mid)
        return ControlNetOutput(down_block_res_samples=down, mid_block_res_sample=mid)
