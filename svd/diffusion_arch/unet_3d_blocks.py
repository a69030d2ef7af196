"""SpatioTemporal block containers — drop-in for the live part (L1870-2396) of the reference's
svd/diffusion_arch/unet_3d_blocks.py plus its get_down_block / get_up_block factories (:39-164, :167-303).

Only the four SVD block types are provided; the 3D / Motion / TemporalDecoder blocks of that file are dead code
in every reference config. GroupNorm eps values follow the reference exactly:
CrossAttnDown 1e-6 (:2098), Down 1e-5 (:1999), Mid 1e-5 (:1895), Up / CrossAttnUp 1e-6 (:2201, :2291; the
factories do not forward `resnet_eps`).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

from torch import nn

from .primitives import Downsample2D, SpatioTemporalResBlock, Upsample2D, _NoForward
from .transformer_temporal import TransformerSpatioTemporalModel


def _per_layer(v: Union[int, Tuple[int, ...]], n: int):
    return [v] * n if isinstance(v, int) else list(v)


class UNetMidBlockSpatioTemporal(_NoForward):
    def __init__(self, in_channels: int, temb_channels: int, num_layers: int = 1,
                 transformer_layers_per_block: Union[int, Tuple[int]] = 1, num_attention_heads: int = 1,
                 cross_attention_dim: int = 1280):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        tl = _per_layer(transformer_layers_per_block, num_layers)
        resnets = [SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5)]
        attentions = []
        for i in range(num_layers):
            attentions.append(TransformerSpatioTemporalModel(
                num_attention_heads, in_channels // num_attention_heads, in_channels=in_channels,
                num_layers=tl[i], cross_attention_dim=cross_attention_dim))
            resnets.append(SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.gradient_checkpointing = False


class DownBlockSpatioTemporal(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, num_layers: int = 1,
                 add_downsample: bool = True):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=1e-5)
            for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         name="op")]) if add_downsample else None)
        self.gradient_checkpointing = False


class CrossAttnDownBlockSpatioTemporal(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, num_layers: int = 1,
                 transformer_layers_per_block: Union[int, Tuple[int]] = 1, num_attention_heads: int = 1,
                 cross_attention_dim: int = 1280, add_downsample: bool = True):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        tl = _per_layer(transformer_layers_per_block, num_layers)
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=1e-6)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads,
                                           in_channels=out_channels, num_layers=tl[i],
                                           cross_attention_dim=cross_attention_dim) for i in range(num_layers)])
        self.downsamplers = (nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                         padding=1, name="op")]) if add_downsample else None)
        self.gradient_checkpointing = False


class UpBlockSpatioTemporal(_NoForward):
    def __init__(self, in_channels: int, prev_output_channel: int, out_channels: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, num_layers: int = 1, resnet_eps: float = 1e-6,
                 add_upsample: bool = True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(cin + skip, out_channels, temb_channels, eps=resnet_eps))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False
        self.resolution_idx = resolution_idx


class CrossAttnUpBlockSpatioTemporal(_NoForward):
    def __init__(self, in_channels: int, out_channels: int, prev_output_channel: int, temb_channels: int,
                 resolution_idx: Optional[int] = None, num_layers: int = 1,
                 transformer_layers_per_block: Union[int, Tuple[int]] = 1, resnet_eps: float = 1e-6,
                 num_attention_heads: int = 1, cross_attention_dim: int = 1280, add_upsample: bool = True):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        tl = _per_layer(transformer_layers_per_block, num_layers)
        resnets, attentions = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            cin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(cin + skip, out_channels, temb_channels, eps=resnet_eps))
            attentions.append(TransformerSpatioTemporalModel(
                num_attention_heads, out_channels // num_attention_heads, in_channels=out_channels,
                num_layers=tl[i], cross_attention_dim=cross_attention_dim))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attentions)
        self.upsamplers = (nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)])
                           if add_upsample else None)
        self.gradient_checkpointing = False
        self.resolution_idx = resolution_idx


def get_down_block(down_block_type: str, num_layers: int, in_channels: int, out_channels: int, temb_channels: int,
                   add_downsample: bool, num_attention_heads: int, resnet_eps=None, resnet_act_fn=None,
                   transformer_layers_per_block: int = 1, cross_attention_dim: Optional[int] = None, **_unused):
    if down_block_type == "DownBlockSpatioTemporal":
        return DownBlockSpatioTemporal(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                       temb_channels=temb_channels, add_downsample=add_downsample)
    if down_block_type == "CrossAttnDownBlockSpatioTemporal":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnDownBlockSpatioTemporal")
        return CrossAttnDownBlockSpatioTemporal(
            in_channels=in_channels, out_channels=out_channels, temb_channels=temb_channels, num_layers=num_layers,
            transformer_layers_per_block=transformer_layers_per_block, add_downsample=add_downsample,
            cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads)
    raise ValueError(f"{down_block_type} does not exist.")


def get_up_block(up_block_type: str, num_layers: int, in_channels: int, out_channels: int, prev_output_channel: int,
                 temb_channels: int, add_upsample: bool, num_attention_heads: int, resnet_eps=None,
                 resnet_act_fn=None, resolution_idx: Optional[int] = None, transformer_layers_per_block: int = 1,
                 cross_attention_dim: Optional[int] = None, **_unused):
    if up_block_type == "UpBlockSpatioTemporal":
        return UpBlockSpatioTemporal(num_layers=num_layers, in_channels=in_channels, out_channels=out_channels,
                                     prev_output_channel=prev_output_channel, temb_channels=temb_channels,
                                     resolution_idx=resolution_idx, add_upsample=add_upsample)
    if up_block_type == "CrossAttnUpBlockSpatioTemporal":
        if cross_attention_dim is None:
            raise ValueError("cross_attention_dim must be specified for CrossAttnUpBlockSpatioTemporal")
        return CrossAttnUpBlockSpatioTemporal(
            in_channels=in_channels, out_channels=out_channels, prev_output_channel=prev_output_channel,
            temb_channels=temb_channels, num_layers=num_layers,
            transformer_layers_per_block=transformer_layers_per_block, add_upsample=add_upsample,
            cross_attention_dim=cross_attention_dim, num_attention_heads=num_attention_heads,
            resolution_idx=resolution_idx)
    raise ValueError(f"{up_block_type} does not exist.")
