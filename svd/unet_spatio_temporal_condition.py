"""UNetSpatioTemporalConditionModel — drop-in for svd/unet_spatio_temporal_condition.py of the reference.

Constructor kwargs, `.config`, state-dict keys and the `forward` signature (:363-373) match the reference; the
arithmetic of forward() runs on hand-written sm_100a kernels through this_and_that_vdm_b200.engine (bf16
storage, fp32 accumulation). There is no eager / CPU implementation in the product: calling forward() on a
module that is not on a CUDA sm_100 device raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import torch
from torch import nn

from .diffusion_arch.primitives import TimestepEmbedding, Timesteps
from .diffusion_arch.unet_3d_blocks import UNetMidBlockSpatioTemporal, get_down_block, get_up_block
from .modeling_utils import ModelBase, register_to_config


@dataclass
class UNetSpatioTemporalConditionOutput:
    sample: torch.FloatTensor = None


class UNetSpatioTemporalConditionModel(ModelBase):
    _supports_gradient_checkpointing = True

    @register_to_config
    def __init__(
        self,
        sample_size: Optional[int] = None,
        in_channels: int = 8,
        out_channels: int = 4,
        down_block_types: Tuple[str] = ("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                                        "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
        up_block_types: Tuple[str] = ("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                                      "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
        block_out_channels: Tuple[int] = (320, 640, 1280, 1280),
        addition_time_embed_dim: int = 256,
        projection_class_embeddings_input_dim: int = 768,
        layers_per_block: Union[int, Tuple[int]] = 2,
        cross_attention_dim: Union[int, Tuple[int]] = 1024,
        transformer_layers_per_block: Union[int, Tuple[int], Tuple[Tuple]] = 1,
        num_attention_heads: Union[int, Tuple[int]] = (5, 10, 10, 20),
        num_frames: int = 25,
    ):
        super().__init__()
        self.sample_size = sample_size
        n = len(down_block_types)
        if n != len(up_block_types):
            raise ValueError(
                f"Must provide the same number of `down_block_types` as `up_block_types`. `down_block_types`: "
                f"{down_block_types}. `up_block_types`: {up_block_types}.")
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

        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], kernel_size=3, padding=1)
        time_embed_dim = block_out_channels[0] * 4
        self.time_proj = Timesteps(block_out_channels[0], True, downscale_freq_shift=0)
        self.time_embedding = TimestepEmbedding(block_out_channels[0], time_embed_dim)
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
        self.up_blocks = nn.ModuleList([])
        output_channel = block_out_channels[0]
        for i, down_block_type in enumerate(down_block_types):
            input_channel, output_channel = output_channel, block_out_channels[i]
            self.down_blocks.append(get_down_block(
                down_block_type, num_layers=layers_per_block[i],
                transformer_layers_per_block=transformer_layers_per_block[i], in_channels=input_channel,
                out_channels=output_channel, temb_channels=time_embed_dim, add_downsample=i != n - 1,
                resnet_eps=1e-5, cross_attention_dim=cross_attention_dim[i],
                num_attention_heads=num_attention_heads[i], resnet_act_fn="silu"))

        self.mid_block = UNetMidBlockSpatioTemporal(
            block_out_channels[-1], temb_channels=time_embed_dim,
            transformer_layers_per_block=transformer_layers_per_block[-1],
            cross_attention_dim=cross_attention_dim[-1], num_attention_heads=num_attention_heads[-1])

        self.num_upsamplers = 0
        rev_ch = list(reversed(block_out_channels))
        rev_heads = list(reversed(num_attention_heads))
        rev_layers = list(reversed(layers_per_block))
        rev_xdim = list(reversed(cross_attention_dim))
        rev_tl = list(reversed(transformer_layers_per_block))
        output_channel = rev_ch[0]
        for i, up_block_type in enumerate(up_block_types):
            is_final = i == n - 1
            prev_output_channel, output_channel = output_channel, rev_ch[i]
            input_channel = rev_ch[min(i + 1, n - 1)]
            if not is_final:
                self.num_upsamplers += 1
            self.up_blocks.append(get_up_block(
                up_block_type, num_layers=rev_layers[i] + 1, transformer_layers_per_block=rev_tl[i],
                in_channels=input_channel, out_channels=output_channel, prev_output_channel=prev_output_channel,
                temb_channels=time_embed_dim, add_upsample=not is_final, resnet_eps=1e-5, resolution_idx=i,
                cross_attention_dim=rev_xdim[i], num_attention_heads=rev_heads[i], resnet_act_fn="silu"))

        self.conv_norm_out = nn.GroupNorm(num_channels=block_out_channels[0], num_groups=32, eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, kernel_size=3, padding=1)
        self._engine = None

    # ------------------------------------------------------------------------------------------ engine plumbing
    def _get_engine(self):
        from this_and_that_vdm_b200.engine import DenoiserEngine
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError(
                "UNetSpatioTemporalConditionModel.forward runs only on a CUDA sm_100 device (hand-written kernels in "
                "libttvdm_sm100.so); move the model with .to('cuda') — there is no CPU / eager fallback")
        if self._engine is None or self._engine.device != dev:
            self._engine = DenoiserEngine(self, kind="unet")
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
    def forward(
        self,
        sample: torch.FloatTensor,
        timestep: Union[torch.Tensor, float, int],
        encoder_hidden_states: torch.Tensor,
        added_time_ids: torch.Tensor,
        added_positions: torch.Tensor = None,
        down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
        mid_block_additional_residual: Optional[torch.Tensor] = None,
        return_dict: bool = True,
    ) -> Union[UNetSpatioTemporalConditionOutput, Tuple]:
        """sample [B, F, C_in, h, w]; timestep python number / 0-dim / 1-dim tensor; encoder_hidden_states
        [B, L, D]; added_time_ids [B, 3]; 12 + 1 optional ControlNet residuals in the reference's NCHW layout
        ([B*F, C, h_l, w_l]). Returns the reference's `[B, F, 4, h, w]` tensor (same dtype as `sample`)."""
        out = self._get_engine().unet_forward(
            sample, timestep, encoder_hidden_states, added_time_ids,
            down_block_additional_residuals, mid_block_additional_residual)
        if not return_dict:
            return (out,)
        return UNetSpatioTemporalConditionOutput(sample=out)
