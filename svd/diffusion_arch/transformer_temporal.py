"""TransformerSpatioTemporalModel — drop-in for svd/diffusion_arch/transformer_temporal.py:201-381 of the reference.

Same constructor signature and state-dict keys. The forward data flow (GroupNorm -> proj_in -> spatial block ->
+frame positional embedding -> temporal block -> alpha blend -> proj_out -> +residual, including the reference's
`time_context` row-selection quirk at :309-319) is executed by this_and_that_vdm_b200.engine on sm_100a kernels.
"""
from __future__ import annotations

from typing import Optional

from torch import nn

from .primitives import (AlphaBlender, BasicTransformerBlock, TemporalBasicTransformerBlock, TimestepEmbedding,
                         Timesteps, _NoForward)


class TransformerSpatioTemporalModel(_NoForward):
    def __init__(self, num_attention_heads: int = 16, attention_head_dim: int = 88, in_channels: int = 320,
                 out_channels: Optional[int] = None, num_layers: int = 1, cross_attention_dim: Optional[int] = None):
        super().__init__()
        self.num_attention_heads = num_attention_heads
        self.attention_head_dim = attention_head_dim
        inner_dim = num_attention_heads * attention_head_dim
        self.inner_dim = inner_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner_dim, num_attention_heads, attention_head_dim,
                                  cross_attention_dim=cross_attention_dim) for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList([
            TemporalBasicTransformerBlock(inner_dim, inner_dim, num_attention_heads, attention_head_dim,
                                          cross_attention_dim=cross_attention_dim) for _ in range(num_layers)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_proj = Timesteps(in_channels, True, 0)
        self.time_mixer = AlphaBlender(alpha=0.5, merge_strategy="learned_with_images")
        self.out_channels = in_channels if out_channels is None else out_channels
        self.proj_out = nn.Linear(inner_dim, in_channels)
        self.gradient_checkpointing = False
