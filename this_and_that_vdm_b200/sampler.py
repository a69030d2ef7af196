"""Fused denoising step for the VL / VGL pipelines (the 25x hot loop of
svd/pipeline_stable_video_diffusion_controlnet.py:623-720 and svd/pipeline_stable_video_diffusion.py:527-562).

Everything that is loop invariant in the reference is hoisted here: the timestep-embedding MLPs and all 44+20
time_emb_proj layers for ALL steps (one batched GEMM chain, K13), the cross-attention K/V of the constant context
(K6/K8), the frame positional embeddings, and — by construction — the VAE-encoded gesture latents (the reference
re-runs vae.encode(condition_img) every step, :652).

Schedule of one VGL step on one GPU holding B_local (1 or 2) of the CFG pair's sequences:
   prepare (CFG duplicate, /sqrt(sigma^2+1), concat image latents [+ gesture latents]) -> UNet encoder + mid ->
   GestureNet encoder + mid -> 13 zero-conv GEMMs whose epilogues accumulate into the UNet skips / mid ->
   UNet decoder -> CFG combine + Euler step (fp32 state).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import lib
from .engine import PAD_IN, DenoiserEngine


class FusedDenoiser:
    def __init__(self, unet_engine: DenoiserEngine, controlnet_engine: Optional[DenoiserEngine] = None):
        self.unet = unet_engine
        self.cn = controlnet_engine
        self.device = unet_engine.device
        self._prepared = False

    def prepare(self, encoder_hidden_states: torch.Tensor, image_latents: torch.Tensor,
                added_time_ids: torch.Tensor, sigmas: torch.Tensor, timesteps: torch.Tensor,
                guidance: torch.Tensor, *, num_frames: int, height: int, width: int,
                controlnet_cond: Optional[torch.Tensor] = None, conditioning_scale: float = 1.0,
                batch_offset: int = 0, b_local: Optional[int] = None) -> None:
        """encoder_hidden_states [B, L, D] (all contexts of the CFG pair — needed for the temporal context quirk even
        when this rank only runs one half), image_latents fp32 [B, 4, h, w] (row 0 zeros under CFG),
        added_time_ids [B, 3], sigmas [n+1], timesteps [n], guidance [F], controlnet_cond fp32 [F, 4, h, w]."""
        dev = self.device
        B = encoder_hidden_states.shape[0]
        self.B = B
        self.b_local = B if b_local is None else b_local
        self.batch_offset = batch_offset
        self.F, self.h, self.w = num_frames, height, width
        if height % 8 or width % 8:
            raise ValueError(f"latent height/width must be multiples of 8, got {height}x{width}")
        self.sigmas = [float(s) for s in sigmas]
        n = len(self.sigmas) - 1
        self.n_steps = n
        self.image_latents = image_latents.to(dev, torch.float32).contiguous()
        self.cond = None if controlnet_cond is None else controlnet_cond.to(dev, torch.float32).contiguous()
        self.guidance = guidance.to(dev, torch.float32).contiguous()
        self.cond_scale = float(conditioning_scale)
        ids = added_time_ids.to(dev, torch.float32)
        sl = slice(batch_offset, batch_offset + self.b_local)
        # all steps' embeddings at once: rows ordered (step, local batch element)
        t_rows = timesteps.to(dev, torch.float32).reshape(n, 1).expand(n, self.b_local).reshape(-1).contiguous()
        id_rows = ids[sl].unsqueeze(0).expand(n, self.b_local, ids.shape[1]).reshape(n * self.b_local, -1).contiguous()
        self.unet._ensure_pos_emb(num_frames)
        self.temb_u = self.unet.time_embeddings(t_rows, id_rows)
        self.kv_u = self.unet.context_kv(encoder_hidden_states.to(dev))
        if self.cn is not None:
            if self.cond is None:
                raise ValueError("controlnet_cond is required when a ControlNet is given")
            self.cn._ensure_pos_emb(num_frames)
            self.temb_c = self.cn.time_embeddings(t_rows, id_rows)
            self.kv_c = self.cn.context_kv(encoder_hidden_states.to(dev))
        rows = self.b_local * num_frames * height * width
        self.x_in = torch.empty(rows, PAD_IN, dtype=torch.bfloat16, device=dev)
        self._prepared = True

    def predict(self, i: int, latents: torch.Tensor) -> torch.Tensor:
        """Noise prediction of step i for the local sequences: fp32 [b_local*F*h*w, 4] channels-last.
        latents: fp32 [F, 4, h, w] (the reference's state, one video)."""
        assert self._prepared
        F, h, w, bl = self.F, self.h, self.w, self.b_local
        lib.sampler_prepare(latents, self.image_latents, self.cond, self.x_in, c_pad=PAD_IN, B_local=bl,
                            batch_offset=self.batch_offset, F=F, h=h, w=w, sigma=self.sigmas[i])
        kw = dict(B=bl, F=F, n_ctx=self.B, batch_offset=self.batch_offset)
        self.unet.begin_step()
        if self.cn is not None:
            self.cn.begin_step()
        temb_u = self.temb_u[i * bl:(i + 1) * bl]
        x, skips, dims, ti = self.unet.encode(self.x_in, temb_u, self.kv_u, H=h, W=w, **kw)
        hl, wl = dims[-1]
        x, ti = self.unet.middle(x, temb_u, self.kv_u, ti, H=hl, W=wl, **kw)
        if self.cn is not None:
            temb_c = self.temb_c[i * bl:(i + 1) * bl]
            cx, cskips, _, cti = self.cn.encode(self.x_in, temb_c, self.kv_c, H=h, W=w, **kw)
            cx, _ = self.cn.middle(cx, temb_c, self.kv_c, cti, H=hl, W=wl, **kw)
            # the down path is finished, so the residuals can be accumulated into the skip tensors in place
            self.cn.zero_convs(cskips, cx, [self.cond_scale] * (len(cskips) + 1), into=skips, mid_into=x, n_img=bl * F)
        return self.unet.decode(x, skips, temb_u, self.kv_u, ti, H=hl, W=wl, **kw)

    def euler_update(self, i: int, latents: torch.Tensor, eps_u: torch.Tensor, eps_c: torch.Tensor) -> None:
        lib.sampler_euler_step(latents, eps_u, eps_c, self.guidance, ld_eps=eps_u.shape[1] if eps_u.dim() == 2 else 4,
                               F=self.F, h=self.h, w=self.w, sigma=self.sigmas[i], sigma_next=self.sigmas[i + 1])

    def step(self, i: int, latents: torch.Tensor) -> None:
        """One full CFG step when this rank holds both halves of the pair (in-place on latents)."""
        if self.b_local != 2 or self.B != 2:
            raise lib.TtvdmError("step() needs both CFG halves on this rank; use predict() + exchange + euler_update()")
        eps = self.predict(i, latents)
        n = self.F * self.h * self.w
        self.euler_update(i, latents, eps[:n], eps[n:])
