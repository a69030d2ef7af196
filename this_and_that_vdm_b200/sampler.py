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

import os
from typing import Optional

import torch

from . import lib
from .engine import MAX_FRAMES, PAD_IN, DenoiserEngine


class FusedDenoiser:
    """use_graph (default: env TTVDM_STEP_GRAPH, on): the network part of a step (UNet encoder + mid, GestureNet, zero
    convs, UNet decoder: ~1000 launches) is captured ONCE into a CUDA graph whose private memory pool is the static
    activation arena, and replayed for every Euler step of every video of the same shape: per step the host then issues a
    small copy (the step's timestep-embedding rows into a static buffer), the prepare kernel, ONE graph launch and the
    Euler kernel — no allocation, no per-kernel launch cost. Everything the graph reads sits at fixed addresses: model
    input, timestep rows, context K/V (overwritten in place per video), frame positional embeddings, weights."""

    def __init__(self, unet_engine: DenoiserEngine, controlnet_engine: Optional[DenoiserEngine] = None,
                 use_graph: Optional[bool] = None):
        self.unet = unet_engine
        self.cn = controlnet_engine
        self.device = unet_engine.device
        self._prepared = False
        if use_graph is None:
            use_graph = os.environ.get("TTVDM_STEP_GRAPH", "1") != "0"
        self.use_graph = bool(use_graph)
        self._graphs = {}      # key -> (graph, eps output)
        self._static = {}      # key -> dict of static buffers (x_in, temb_u, temb_c, kv_u, kv_c)
        self._warm = {}        # key -> eager steps run so far (pool sizes are learned before capture)

    def _key(self):
        return (self.B, self.b_local, self.batch_offset, self.F, self.h, self.w, self.cn is not None)

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
        if num_frames > MAX_FRAMES:
            raise lib.TtvdmError(f"num_frames = {num_frames}: the temporal-attention kernel holds at most {MAX_FRAMES} "
                                 f"frames per sequence (the reference runs 14, SVD-XT 25)")
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
        if self.cn is not None and self.cond is None:
            raise ValueError("controlnet_cond is required when a ControlNet is given")
        st = self._static.setdefault(self._key(), {})
        L = encoder_hidden_states.shape[1]
        if st.get("L") != L:  # context length changed (use_text on / off): the K/V buffers and the graph are stale
            st.clear()
            self._graphs.pop(self._key(), None)
            self._warm.pop(self._key(), None)
            st["L"] = L
        self.unet._ensure_pos_emb(num_frames)
        self.temb_u = self.unet.time_embeddings(t_rows, id_rows)
        st["kv_u"] = self.kv_u = self.unet.context_kv(encoder_hidden_states.to(dev), out=st.get("kv_u"))
        if self.cn is not None:
            self.cn._ensure_pos_emb(num_frames)
            self.temb_c = self.cn.time_embeddings(t_rows, id_rows)
            st["kv_c"] = self.kv_c = self.cn.context_kv(encoder_hidden_states.to(dev), out=st.get("kv_c"))
        rows = self.b_local * num_frames * height * width
        if "x_in" not in st:
            st["x_in"] = torch.empty(rows, PAD_IN, dtype=torch.bfloat16, device=dev)
            st["temb_u"] = torch.empty(self.b_local, self.temb_u.shape[1], dtype=torch.float32, device=dev)
            if self.cn is not None:
                st["temb_c"] = torch.empty(self.b_local, self.temb_c.shape[1], dtype=torch.float32, device=dev)
        self.x_in = st["x_in"]
        self._prepared = True

    # ---- the network part of a step: reads only static buffers (graph capturable)
    def _network(self, temb_u, temb_c, cond_scale: float) -> torch.Tensor:
        F, h, w, bl = self.F, self.h, self.w, self.b_local
        kw = dict(B=bl, F=F, n_ctx=self.B, batch_offset=self.batch_offset)
        self.unet.begin_step()
        if self.cn is not None:
            self.cn.begin_step()
        x, skips, dims, ti = self.unet.encode(self.x_in, temb_u, self.kv_u, H=h, W=w, **kw)
        hl, wl = dims[-1]
        x, ti = self.unet.middle(x, temb_u, self.kv_u, ti, H=hl, W=wl, **kw)
        if self.cn is not None:
            cx, cskips, _, cti = self.cn.encode(self.x_in, temb_c, self.kv_c, H=h, W=w, **kw)
            cx, _ = self.cn.middle(cx, temb_c, self.kv_c, cti, H=hl, W=wl, **kw)
            # the down path is finished, so the residuals can be accumulated into the skip tensors in place
            self.cn.zero_convs(cskips, cx, [cond_scale] * (len(cskips) + 1), into=skips, mid_into=x, n_img=bl * F)
        return self.unet.decode(x, skips, temb_u, self.kv_u, ti, H=hl, W=wl, **kw)

    def predict(self, i: int, latents: torch.Tensor) -> torch.Tensor:
        """Noise prediction of step i for the local sequences: fp32 [b_local*F*h*w, 4] channels-last.
        latents: fp32 [F, 4, h, w] (the reference's state, one video)."""
        assert self._prepared
        F, h, w, bl = self.F, self.h, self.w, self.b_local
        lib.sampler_prepare(latents, self.image_latents, self.cond, self.x_in, c_pad=PAD_IN, B_local=bl,
                            batch_offset=self.batch_offset, F=F, h=h, w=w, sigma=self.sigmas[i])
        temb_u = self.temb_u[i * bl:(i + 1) * bl]
        temb_c = self.temb_c[i * bl:(i + 1) * bl] if self.cn is not None else None
        if not self.use_graph or self.device.type != "cuda" or lib.profiling():
            return self._network(temb_u, temb_c, self.cond_scale)
        key = self._key()
        st = self._static[key]
        st["temb_u"].copy_(temb_u)
        if self.cn is not None:
            st["temb_c"].copy_(temb_c)
        gkey = (key, self.cond_scale)
        ent = self._graphs.get(key)
        if ent is not None and ent[2] != self.cond_scale:
            # the conditioning scale is baked into the captured zero-conv launches; a per-step schedule of scales
            # (control_guidance_start / end) runs eagerly instead of re-capturing every step
            return self._network(st["temb_u"], st.get("temb_c"), self.cond_scale)
        if ent is None:
            warm = self._warm.get(gkey, 0)
            if warm < 2:
                # two eager steps first: the statistics pools learn their size, every lazy initialisation has happened
                self._warm[gkey] = warm + 1
                return self._network(st["temb_u"], st.get("temb_c"), self.cond_scale)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            n0 = lib.launch_count()
            with torch.cuda.graph(graph):
                eps = self._network(st["temb_u"], st.get("temb_c"), self.cond_scale)
            ent = (graph, eps, self.cond_scale, lib.launch_count() - n0)  # kernel nodes of this library in the graph
            lib.add_graph_launches(-ent[3])  # the capture itself launched nothing
            self._graphs[key] = ent
        ent[0].replay()
        lib.add_graph_launches(ent[3])
        return ent[1]

    @staticmethod
    def cached(unet_engine: DenoiserEngine, controlnet_engine: Optional[DenoiserEngine] = None) -> "FusedDenoiser":
        """One FusedDenoiser per (UNet engine, GestureNet engine) pair, so the captured step graphs and their static
        buffers survive across pipeline calls."""
        cache = unet_engine.__dict__.setdefault("_fused_denoisers", {})
        key = id(controlnet_engine) if controlnet_engine is not None else 0
        den = cache.get(key)
        if den is None or den.cn is not controlnet_engine:
            den = cache[key] = FusedDenoiser(unet_engine, controlnet_engine)
        return den

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
