"""B200 execution engine of the This&That / SVD denoiser: weight packing + kernel schedule.

Host-side (Python) orchestration of the sm_100a kernels in libttvdm_sm100.so for
  * UNetSpatioTemporalConditionModel.forward   (reference svd/unet_spatio_temporal_condition.py:363-536)
  * ControlNetModel.forward ("GestureNet")     (reference svd/temporal_controlnet.py:455-641)
  * the fused VGL/VL step used by the pipelines (this_and_that_vdm_b200/sampler.py)

Data layout in HBM: every activation is a bf16 token matrix [B*F*S, C] (== channels-last NHWC per frame), rows
ordered (b, f, s) — the same order as the reference's `[batch*frames, h*w, C]` tokens, so the conv <-> transformer
boundary needs no permute and the temporal layers walk frames with a row stride of S. Weights are packed once
(bf16, K-major, 3x3 taps tap-major; GEGLU hidden/gate rows interleaved; q|k|v fused) and stay resident.

torch is used only to own device memory and for boundary layout conversion (NCHW fp32/fp16 <-> NHWC bf16) in
the stand-alone forward() calls; the fused sampler path has its own glue kernels.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib

BF16 = torch.bfloat16
PAD_IN = 64  # conv_in channels padded to one 64-wide K chunk


def _bf(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=BF16).contiguous()


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _pack_conv3x3(w: torch.Tensor, dev, pad_cin: int = 0) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin'] tap-major then channel (Cin' = Cin zero-padded to pad_cin)."""
    cout, cin = w.shape[:2]
    w = w.detach().permute(0, 2, 3, 1)  # [Cout, 3, 3, Cin]
    if pad_cin and pad_cin > cin:
        w = torch.nn.functional.pad(w, (0, pad_cin - cin))
    return _bf(w.reshape(cout, -1), dev)


def _pack_tconv(w: torch.Tensor, dev) -> torch.Tensor:
    """Conv3d (3,1,1) weight [C, C, 3, 1, 1] -> [C, 3*C] tap-major."""
    cout, cin = w.shape[:2]
    return _bf(w.detach().reshape(cout, cin, 3).permute(0, 2, 1).reshape(cout, 3 * cin), dev)


def _pack_geglu(w: torch.Tensor, b: torch.Tensor, dev) -> Tuple[torch.Tensor, torch.Tensor]:
    """GEGLU proj [8C, C]: rows [0,4C) hidden, [4C,8C) gate -> interleaved (hidden_j, gate_j)."""
    n2 = w.shape[0] // 2
    wi = torch.stack([w[:n2], w[n2:]], dim=1).reshape(w.shape[0], w.shape[1])
    bi = torch.stack([b[:n2], b[n2:]], dim=1).reshape(-1)
    return _bf(wi, dev), _f32(bi, dev)


@dataclass
class ResW:
    cin: int
    cout: int
    eps: float
    n1_g: torch.Tensor = None
    n1_b: torch.Tensor = None
    w1: torch.Tensor = None
    b1: torch.Tensor = None
    n2_g: torch.Tensor = None
    n2_b: torch.Tensor = None
    w2: torch.Tensor = None
    b2: torch.Tensor = None
    wsc: Optional[torch.Tensor] = None
    bsc: Optional[torch.Tensor] = None
    tn1_g: torch.Tensor = None
    tn1_b: torch.Tensor = None
    tw1: torch.Tensor = None
    tb1: torch.Tensor = None
    tn2_g: torch.Tensor = None
    tn2_b: torch.Tensor = None
    tw2: torch.Tensor = None
    tb2: torch.Tensor = None
    alpha: float = 0.5
    temb_off_s: int = 0  # column offsets into the fused time_emb_proj output
    temb_off_t: int = 0


@dataclass
class AttnW:
    wqkv: Optional[torch.Tensor] = None  # self-attn fused [3C, C]
    wq: Optional[torch.Tensor] = None    # cross-attn query [C, C]
    wk: Optional[torch.Tensor] = None    # cross-attn [C, 1024]
    wv: Optional[torch.Tensor] = None
    wo: torch.Tensor = None
    bo: torch.Tensor = None


@dataclass
class FFW:
    w1: torch.Tensor = None
    b1: torch.Tensor = None
    w2: torch.Tensor = None
    b2: torch.Tensor = None


@dataclass
class TfW:
    C: int
    heads: int
    gn_g: torch.Tensor = None
    gn_b: torch.Tensor = None
    w_in: torch.Tensor = None
    b_in: torch.Tensor = None
    ln: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = field(default_factory=dict)
    s_attn1: AttnW = None
    s_attn2: AttnW = None
    s_ff: FFW = None
    t_ff_in: FFW = None
    t_attn1: AttnW = None
    t_attn2: AttnW = None
    t_ff: FFW = None
    pos_emb: torch.Tensor = None  # fp32 [F, C], input independent
    alpha: float = 0.5
    w_out: torch.Tensor = None
    b_out: torch.Tensor = None
    pos_w1: torch.Tensor = None
    pos_b1: torch.Tensor = None
    pos_w2: torch.Tensor = None
    pos_b2: torch.Tensor = None


class DenoiserEngine:
    """Packed weights + kernel schedule for one network (kind = 'unet' | 'controlnet')."""

    def __init__(self, model, kind: str):
        lib.init()
        self.kind = kind
        self.device = model.device
        cfg = model.config
        self.cfg = cfg
        self.chans = tuple(cfg.block_out_channels)
        self.heads = tuple(cfg.num_attention_heads) if not isinstance(cfg.num_attention_heads, int) else \
            (cfg.num_attention_heads,) * len(self.chans)
        for c, h in zip(self.chans, self.heads):
            if c % 64 != 0 or c // h != 64:
                raise lib.TtvdmError(
                    f"sm_100a engine supports head_dim 64 and channels %% 64 == 0 (got C={c}, heads={h})")
        self.temb_dim = self.chans[0] * 4
        self.add_dim = cfg.addition_time_embed_dim
        self._pos_cache: Dict[int, bool] = {}
        self._pack(model)

    # ============================================================================================ packing
    def _pack(self, m) -> None:
        dev = self.device
        sd = {k: v for k, v in m.state_dict().items()}
        g = lambda k: sd[k]  # noqa: E731
        self.temb_cols = 0
        temb_w, temb_b = [], []

        def res(prefix: str, eps: float) -> ResW:
            sp, tp = prefix + ".spatial_res_block", prefix + ".temporal_res_block"
            cout, cin = g(sp + ".conv1.weight").shape[:2]
            r = ResW(cin=cin, cout=cout, eps=eps)
            r.n1_g, r.n1_b = _f32(g(sp + ".norm1.weight"), dev), _f32(g(sp + ".norm1.bias"), dev)
            r.w1, r.b1 = _pack_conv3x3(g(sp + ".conv1.weight"), dev), _f32(g(sp + ".conv1.bias"), dev)
            r.n2_g, r.n2_b = _f32(g(sp + ".norm2.weight"), dev), _f32(g(sp + ".norm2.bias"), dev)
            r.w2, r.b2 = _pack_conv3x3(g(sp + ".conv2.weight"), dev), _f32(g(sp + ".conv2.bias"), dev)
            if (sp + ".conv_shortcut.weight") in sd:
                r.wsc = _bf(g(sp + ".conv_shortcut.weight").reshape(cout, cin), dev)
                r.bsc = _f32(g(sp + ".conv_shortcut.bias"), dev)
            r.tn1_g, r.tn1_b = _f32(g(tp + ".norm1.weight"), dev), _f32(g(tp + ".norm1.bias"), dev)
            r.tw1, r.tb1 = _pack_tconv(g(tp + ".conv1.weight"), dev), _f32(g(tp + ".conv1.bias"), dev)
            r.tn2_g, r.tn2_b = _f32(g(tp + ".norm2.weight"), dev), _f32(g(tp + ".norm2.bias"), dev)
            r.tw2, r.tb2 = _pack_tconv(g(tp + ".conv2.weight"), dev), _f32(g(tp + ".conv2.bias"), dev)
            r.alpha = float(torch.sigmoid(g(prefix + ".time_mixer.mix_factor").float()).item())
            # all time_emb_proj layers are fused into ONE [sum C, 1280] GEMM per forward (K13 hoist)
            r.temb_off_s = self.temb_cols
            temb_w.append(g(sp + ".time_emb_proj.weight")); temb_b.append(g(sp + ".time_emb_proj.bias"))
            self.temb_cols += cout
            r.temb_off_t = self.temb_cols
            temb_w.append(g(tp + ".time_emb_proj.weight")); temb_b.append(g(tp + ".time_emb_proj.bias"))
            self.temb_cols += cout
            return r

        def attn(prefix: str, cross: bool) -> AttnW:
            a = AttnW()
            if cross:
                a.wq = _bf(g(prefix + ".to_q.weight"), dev)
                a.wk = _bf(g(prefix + ".to_k.weight"), dev)
                a.wv = _bf(g(prefix + ".to_v.weight"), dev)
            else:
                a.wqkv = _bf(torch.cat([g(prefix + ".to_q.weight"), g(prefix + ".to_k.weight"),
                                        g(prefix + ".to_v.weight")], 0), dev)
            a.wo, a.bo = _bf(g(prefix + ".to_out.0.weight"), dev), _f32(g(prefix + ".to_out.0.bias"), dev)
            return a

        def ff(prefix: str) -> FFW:
            f = FFW()
            f.w1, f.b1 = _pack_geglu(g(prefix + ".net.0.proj.weight"), g(prefix + ".net.0.proj.bias"), dev)
            f.w2, f.b2 = _bf(g(prefix + ".net.2.weight"), dev), _f32(g(prefix + ".net.2.bias"), dev)
            return f

        def tf(prefix: str, heads: int) -> TfW:
            C = g(prefix + ".proj_in.weight").shape[0]
            t = TfW(C=C, heads=heads)
            if (prefix + ".transformer_blocks.1.norm1.weight") in sd:
                raise lib.TtvdmError("transformer_layers_per_block > 1 is not supported by the sm_100a engine")
            t.gn_g, t.gn_b = _f32(g(prefix + ".norm.weight"), dev), _f32(g(prefix + ".norm.bias"), dev)
            t.w_in, t.b_in = _bf(g(prefix + ".proj_in.weight"), dev), _f32(g(prefix + ".proj_in.bias"), dev)
            sb, tb = prefix + ".transformer_blocks.0", prefix + ".temporal_transformer_blocks.0"
            for name, p in [("s1", sb + ".norm1"), ("s2", sb + ".norm2"), ("s3", sb + ".norm3"),
                            ("tin", tb + ".norm_in"), ("t1", tb + ".norm1"), ("t2", tb + ".norm2"),
                            ("t3", tb + ".norm3")]:
                t.ln[name] = (_f32(g(p + ".weight"), dev), _f32(g(p + ".bias"), dev))
            t.s_attn1, t.s_attn2, t.s_ff = attn(sb + ".attn1", False), attn(sb + ".attn2", True), ff(sb + ".ff")
            t.t_ff_in, t.t_attn1 = ff(tb + ".ff_in"), attn(tb + ".attn1", False)
            t.t_attn2, t.t_ff = attn(tb + ".attn2", True), ff(tb + ".ff")
            t.alpha = float(torch.sigmoid(g(prefix + ".time_mixer.mix_factor").float()).item())
            t.w_out, t.b_out = _bf(g(prefix + ".proj_out.weight"), dev), _f32(g(prefix + ".proj_out.bias"), dev)
            pe = prefix + ".time_pos_embed"
            t.pos_w1, t.pos_b1 = _bf(g(pe + ".linear_1.weight"), dev), _f32(g(pe + ".linear_1.bias"), dev)
            t.pos_w2, t.pos_b2 = _bf(g(pe + ".linear_2.weight"), dev), _f32(g(pe + ".linear_2.bias"), dev)
            return t

        n_levels = len(self.chans)
        cross_down = [t.startswith("CrossAttn") for t in self.cfg.down_block_types]
        # ---- embeddings
        self.te_w1, self.te_b1 = _bf(g("time_embedding.linear_1.weight"), dev), _f32(g("time_embedding.linear_1.bias"), dev)
        self.te_w2, self.te_b2 = _bf(g("time_embedding.linear_2.weight"), dev), _f32(g("time_embedding.linear_2.bias"), dev)
        self.ae_w1, self.ae_b1 = _bf(g("add_embedding.linear_1.weight"), dev), _f32(g("add_embedding.linear_1.bias"), dev)
        self.ae_w2, self.ae_b2 = _bf(g("add_embedding.linear_2.weight"), dev), _f32(g("add_embedding.linear_2.bias"), dev)
        # ---- conv_in
        cin_name = "conv_in" if self.kind == "unet" else "conv_in_concat"
        self.conv_in_w = _pack_conv3x3(g(cin_name + ".weight"), dev, pad_cin=PAD_IN)
        self.conv_in_b = _f32(g(cin_name + ".bias"), dev)
        # ---- down / mid
        self.down: List[dict] = []
        for i in range(n_levels):
            p = f"down_blocks.{i}"
            blk = {"res": [], "tf": [], "down_w": None, "down_b": None}
            j = 0
            while f"{p}.resnets.{j}.spatial_res_block.norm1.weight" in sd:
                blk["res"].append(res(f"{p}.resnets.{j}", 1e-6 if cross_down[i] else 1e-5))
                if cross_down[i]:
                    blk["tf"].append(tf(f"{p}.attentions.{j}", self.heads[i]))
                j += 1
            if f"{p}.downsamplers.0.conv.weight" in sd:
                blk["down_w"] = _pack_conv3x3(g(f"{p}.downsamplers.0.conv.weight"), dev)
                blk["down_b"] = _f32(g(f"{p}.downsamplers.0.conv.bias"), dev)
            self.down.append(blk)
        self.mid = {"res": [res("mid_block.resnets.0", 1e-5), res("mid_block.resnets.1", 1e-5)],
                    "tf": [tf("mid_block.attentions.0", self.heads[-1])]}
        # ---- up (UNet) or zero convs (ControlNet)
        self.up: List[dict] = []
        if self.kind == "unet":
            cross_up = [t.startswith("CrossAttn") for t in self.cfg.up_block_types]
            rheads = list(reversed(self.heads))
            for i in range(n_levels):
                p = f"up_blocks.{i}"
                blk = {"res": [], "tf": [], "up_w": None, "up_b": None}
                j = 0
                while f"{p}.resnets.{j}.spatial_res_block.norm1.weight" in sd:
                    blk["res"].append(res(f"{p}.resnets.{j}", 1e-6))
                    if cross_up[i]:
                        blk["tf"].append(tf(f"{p}.attentions.{j}", rheads[i]))
                    j += 1
                if f"{p}.upsamplers.0.conv.weight" in sd:
                    blk["up_w"] = _pack_conv3x3(g(f"{p}.upsamplers.0.conv.weight"), dev)
                    blk["up_b"] = _f32(g(f"{p}.upsamplers.0.conv.bias"), dev)
                self.up.append(blk)
            self.out_g, self.out_b = _f32(g("conv_norm_out.weight"), dev), _f32(g("conv_norm_out.bias"), dev)
            self.conv_out_w = _pack_conv3x3(g("conv_out.weight"), dev)
            self.conv_out_b = _f32(g("conv_out.bias"), dev)
            self.out_channels = g("conv_out.weight").shape[0]
        else:
            self.zero_w, self.zero_b = [], []
            i = 0
            while f"controlnet_down_blocks.{i}.weight" in sd:
                w = g(f"controlnet_down_blocks.{i}.weight")
                self.zero_w.append(_bf(w.reshape(w.shape[0], w.shape[1]), dev))
                self.zero_b.append(_f32(g(f"controlnet_down_blocks.{i}.bias"), dev))
                i += 1
            w = g("controlnet_mid_block.weight")
            self.zero_mid_w = _bf(w.reshape(w.shape[0], w.shape[1]), dev)
            self.zero_mid_b = _f32(g("controlnet_mid_block.bias"), dev)
        # ---- fused time_emb_proj
        self.temb_w = _bf(torch.cat(temb_w, 0), dev)
        self.temb_b = _f32(torch.cat(temb_b, 0), dev)

    def all_transformers(self) -> List[TfW]:
        out = []
        for blk in self.down:
            out += blk["tf"]
        out += self.mid["tf"]
        for blk in self.up:
            out += blk["tf"]
        return out

    # ============================================================================================ small helpers
    def _empty(self, *shape, dtype=BF16) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def _linear(self, a, w, *, M, bias=None, out=None, res1=None, s1=1.0, res2=None, s2=1.0, s0=1.0, geglu=False,
                a2=None, k2=0, act=0, out_fp32=False, lda=None):
        N, K = w.shape[0], w.shape[1] - k2
        if out is None:
            out = self._empty(M, N // 2 if geglu else N, dtype=torch.float32 if out_fp32 else BF16)
        lib.gemm(a, w, out, M=M, N=N, k1=K, a2=a2, k2=k2, bias=bias, res1=res1, s1=s1, res2=res2, s2=s2, s0=s0,
                 geglu=geglu, act=act, out_fp32=out_fp32, lda=lda)
        return out

    def _conv3(self, x, w, bias, *, n_img, H, W, cin, out=None, rowvec=None, rows_per_vec=0, ldrv=0, res1=None,
               out_fp32=False):
        N = w.shape[0]
        M = n_img * H * W
        if out is None:
            out = self._empty(M, N, dtype=torch.float32 if out_fp32 else BF16)
        lib.gemm(x, w, out, M=M, N=N, k1=cin, mode=lib.A_CONV3X3, n_img=n_img, H=H, W=W, bias=bias, rowvec=rowvec,
                 rows_per_vec=rows_per_vec, ldrv=ldrv, res1=res1, out_fp32=out_fp32)
        return out

    def _gn(self, x, gamma, beta, *, rows, rows_per_inst, eps, silu, x2=None):
        c1 = x.shape[1]
        c2 = x2.shape[1] if x2 is not None else 0
        out = self._empty(rows, c1 + c2)
        stats = self._empty((rows // rows_per_inst) * 64, dtype=torch.float64)
        lib.groupnorm(x, out, stats, gamma, beta, c1=c1, rows=rows, rows_per_inst=rows_per_inst, eps=eps, silu=silu,
                      x2=x2, c2=c2)
        return out

    def _ln(self, x, gb, *, rows, C, addvec=None, F=0, S=0, sum_out=None):
        out = self._empty(rows, C)
        lib.layernorm(x, out, gb[0], gb[1], rows=rows, C=C, addvec=addvec, F=F, S=S, sum_out=sum_out)
        return out

    # ============================================================================================ embeddings
    def time_embeddings(self, timesteps: torch.Tensor, added_time_ids: torch.Tensor) -> torch.Tensor:
        """U2 (svd/unet_spatio_temporal_condition.py:399-432) + every ResBlock's time_emb_proj(silu(emb)) in one
        GEMM. timesteps fp32 [R] (one row per (step, batch element)), added_time_ids fp32 [R, 3].
        Returns fp32 [R, sum C] — the per-layer additive shifts."""
        R = timesteps.shape[0]
        C0 = self.chans[0]
        ts = self._empty(R, C0)
        lib.sinusoid(timesteps.contiguous(), ts, n=R, dim=C0)
        h = self._linear(ts, self.te_w1, M=R, bias=self.te_b1, act=1)
        emb = self._linear(h, self.te_w2, M=R, bias=self.te_b2)
        n_ids = added_time_ids.shape[1]
        ids = self._empty(R * n_ids, self.add_dim)
        lib.sinusoid(added_time_ids.reshape(-1).contiguous(), ids, n=R * n_ids, dim=self.add_dim)
        h2 = self._linear(ids.view(R, n_ids * self.add_dim), self.ae_w1, M=R, bias=self.ae_b1, act=1)
        # silu(emb + aug_emb): the only consumer of `emb` is time_emb_proj(silu(emb))
        semb = self._linear(h2, self.ae_w2, M=R, bias=self.ae_b2, res1=emb, act=1)
        return self._linear(semb, self.temb_w, M=R, bias=self.temb_b, out_fp32=True)

    def context_kv(self, ehs: torch.Tensor) -> List[Tuple[torch.Tensor, ...]]:
        """Cross-attention K/V of the (constant) context for every transformer: computed once per video instead of
        per frame / per pixel (reference: svd/unet_spatio_temporal_condition.py:452 repeat_interleave,
        svd/diffusion_arch/transformer_temporal.py:316-319 broadcast). ehs [B, L, D] -> per transformer
        (k_s, v_s, k_t, v_t), each bf16 [B, L, C]."""
        B, L, D = ehs.shape
        x = _bf(ehs.reshape(B * L, D), self.device)
        out = []
        for t in self.all_transformers():
            ks = self._linear(x, t.s_attn2.wk, M=B * L)
            vs = self._linear(x, t.s_attn2.wv, M=B * L)
            kt = self._linear(x, t.t_attn2.wk, M=B * L)
            vt = self._linear(x, t.t_attn2.wv, M=B * L)
            out.append((ks, vs, kt, vt))
        return out

    def _ensure_pos_emb(self, F: int) -> None:
        """time_pos_embed(time_proj(arange(F))) is input independent (transformer_temporal.py:328-339): once."""
        if F in self._pos_cache:
            return
        frames = torch.arange(F, device=self.device, dtype=torch.float32)
        for t in self.all_transformers():
            ts = self._empty(F, t.C)
            lib.sinusoid(frames, ts, n=F, dim=t.C)
            h = self._linear(ts, t.pos_w1, M=F, bias=t.pos_b1, act=1)
            t.pos_emb = self._linear(h, t.pos_w2, M=F, bias=t.pos_b2, out_fp32=True)
        self._pos_cache[F] = True

    # ============================================================================================ blocks
    def _resblock(self, r: ResW, x, skip, *, B, F, H, W, temb, out=None, out_res2=None, out_s2=0.0):
        """SpatioTemporalResBlock (diffusers; A.3-A.6). x [rows, c], skip optional second source (channel concat).
        temb fp32 [B, sum C]. Returns [rows, cout]."""
        S = H * W
        rows = B * F * S
        n_img = B * F
        ldrv = self.temb_cols
        y = self._gn(x, r.n1_g, r.n1_b, rows=rows, rows_per_inst=S, eps=r.eps, silu=True, x2=skip)
        h = self._conv3(y, r.w1, r.b1, n_img=n_img, H=H, W=W, cin=r.cin, rowvec=temb[:, r.temb_off_s:],
                        rows_per_vec=F * S, ldrv=ldrv)
        y = self._gn(h, r.n2_g, r.n2_b, rows=rows, rows_per_inst=S, eps=r.eps, silu=True)
        if r.wsc is not None:
            if skip is not None:
                sc = self._linear(x, r.wsc, M=rows, bias=r.bsc, a2=skip, k2=skip.shape[1])
            else:
                sc = self._linear(x, r.wsc, M=rows, bias=r.bsc)
        else:
            sc = x
        hs = self._conv3(y, r.w2, r.b2, n_img=n_img, H=H, W=W, cin=r.cout, res1=sc, out=h)
        # temporal branch: 5-D GroupNorm (stats over all frames of a video) + 3-tap conv over frames
        y = self._gn(hs, r.tn1_g, r.tn1_b, rows=rows, rows_per_inst=F * S, eps=r.eps, silu=True)
        t1 = self._empty(rows, r.cout)
        lib.gemm(y, r.tw1, t1, M=rows, N=r.cout, k1=r.cout, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=r.tb1,
                 rowvec=temb[:, r.temb_off_t:], rows_per_vec=F * S, ldrv=ldrv)
        y = self._gn(t1, r.tn2_g, r.tn2_b, rows=rows, rows_per_inst=F * S, eps=r.eps, silu=True)
        # AlphaBlender: a*hs + (1-a)*(hs + conv) = hs + (1-a)*conv ; optional fused extra residual (ControlNet)
        if out is None:
            out = t1
        lib.gemm(y, r.tw2, out, M=rows, N=r.cout, k1=r.cout, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=r.tb2,
                 s0=1.0 - r.alpha, res1=hs, s1=1.0, res2=out_res2, s2=out_s2)
        return out

    def _transformer(self, t: TfW, x, kv, *, B, F, H, W, n_ctx, batch_offset):
        """TransformerSpatioTemporalModel.forward (svd/diffusion_arch/transformer_temporal.py:276-381)."""
        S = H * W
        rows = B * F * S
        C = t.C
        ks, vs, kt, vt = kv
        L = ks.shape[0] // n_ctx
        scale = 0.125
        y = self._gn(x, t.gn_g, t.gn_b, rows=rows, rows_per_inst=S, eps=1e-6, silu=False)
        h = self._linear(y, t.w_in, M=rows, bias=t.b_in)
        # ---- spatial BasicTransformerBlock
        y = self._ln(h, t.ln["s1"], rows=rows, C=C)
        qkv = self._linear(y, t.s_attn1.wqkv, M=rows)
        o = y  # reuse
        lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], o, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=B * F,
                         heads=t.heads, seq=S, scale=scale)
        self._linear(o, t.s_attn1.wo, M=rows, bias=t.s_attn1.bo, res1=h, out=h)
        y = self._ln(h, t.ln["s2"], rows=rows, C=C)
        q = self._linear(y, t.s_attn2.wq, M=rows)
        lib.attn_cross(q, ks, vs, y, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                       temporal=False, batch_offset=batch_offset, scale=scale)
        self._linear(y, t.s_attn2.wo, M=rows, bias=t.s_attn2.bo, res1=h, out=h)
        y = self._ln(h, t.ln["s3"], rows=rows, C=C)
        gg = self._linear(y, t.s_ff.w1, M=rows, bias=t.s_ff.b1, geglu=True)
        self._linear(gg, t.s_ff.w2, M=rows, bias=t.s_ff.b2, res1=h, out=h)  # h == x_spatial
        # ---- TemporalBasicTransformerBlock on (h + frame positional embedding); rows stay (b, f, s)
        hm = self._empty(rows, C)
        y = self._ln(h, t.ln["tin"], rows=rows, C=C, addvec=t.pos_emb, F=F, S=S, sum_out=hm)
        gg = self._linear(y, t.t_ff_in.w1, M=rows, bias=t.t_ff_in.b1, geglu=True, out=gg)
        self._linear(gg, t.t_ff_in.w2, M=rows, bias=t.t_ff_in.b2, res1=hm, out=hm)
        y = self._ln(hm, t.ln["t1"], rows=rows, C=C)
        qkv = self._linear(y, t.t_attn1.wqkv, M=rows, out=qkv)
        lib.attn_temporal(qkv, qkv[:, C:], qkv[:, 2 * C:], y, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, B=B, F=F, S=S,
                          heads=t.heads, scale=scale)
        self._linear(y, t.t_attn1.wo, M=rows, bias=t.t_attn1.bo, res1=hm, out=hm)
        y = self._ln(hm, t.ln["t2"], rows=rows, C=C)
        q = self._linear(y, t.t_attn2.wq, M=rows, out=q)
        lib.attn_cross(q, kt, vt, y, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                       temporal=True, batch_offset=batch_offset, scale=scale)
        self._linear(y, t.t_attn2.wo, M=rows, bias=t.t_attn2.bo, res1=hm, out=hm)
        y = self._ln(hm, t.ln["t3"], rows=rows, C=C)
        gg = self._linear(y, t.t_ff.w1, M=rows, bias=t.t_ff.b1, geglu=True, out=gg)
        # AlphaBlender fused: a*h + (1-a)*(ff + hm)
        a = t.alpha
        self._linear(gg, t.t_ff.w2, M=rows, bias=t.t_ff.b2, s0=1.0 - a, res1=hm, s1=1.0 - a, res2=h, s2=a, out=h)
        # ---- proj_out + input residual
        return self._linear(h, t.w_out, M=rows, bias=t.b_out, res1=x, out=hm)

    # ============================================================================================ network halves
    def encode(self, x_in, temb, kvs, *, B, F, H, W, n_ctx, batch_offset):
        """conv_in + down blocks. Returns (x, skips[12], dims[12], ti) — shared by UNet and ControlNet."""
        n_img = B * F
        x = self._conv3(x_in, self.conv_in_w, self.conv_in_b, n_img=n_img, H=H, W=W, cin=PAD_IN)
        skips, dims = [x], [(H, W)]
        ti = 0
        for blk in self.down:
            for j, r in enumerate(blk["res"]):
                x = self._resblock(r, x, None, B=B, F=F, H=H, W=W, temb=temb)
                if blk["tf"]:
                    x = self._transformer(blk["tf"][j], x, kvs[ti], B=B, F=F, H=H, W=W, n_ctx=n_ctx,
                                          batch_offset=batch_offset)
                    ti += 1
                skips.append(x)
                dims.append((H, W))
            if blk["down_w"] is not None:
                C = x.shape[1]
                col = self._empty(n_img * (H // 2) * (W // 2), 9 * C)
                lib.im2col_s2(x, col, n_img=n_img, H=H, W=W, C=C)
                H, W = H // 2, W // 2
                x = self._linear(col, blk["down_w"], M=n_img * H * W, bias=blk["down_b"])
                skips.append(x)
                dims.append((H, W))
        return x, skips, dims, ti

    def middle(self, x, temb, kvs, ti, *, B, F, H, W, n_ctx, batch_offset):
        x = self._resblock(self.mid["res"][0], x, None, B=B, F=F, H=H, W=W, temb=temb)
        x = self._transformer(self.mid["tf"][0], x, kvs[ti], B=B, F=F, H=H, W=W, n_ctx=n_ctx,
                              batch_offset=batch_offset)
        x = self._resblock(self.mid["res"][1], x, None, B=B, F=F, H=H, W=W, temb=temb)
        return x, ti + 1

    def decode(self, x, skips, temb, kvs, ti, *, B, F, H, W, n_ctx, batch_offset):
        """up blocks + conv_norm_out/SiLU/conv_out. Returns fp32 [rows, out_channels]."""
        n_img = B * F
        skips = list(skips)
        for blk in self.up:
            for j, r in enumerate(blk["res"]):
                x = self._resblock(r, x, skips.pop(), B=B, F=F, H=H, W=W, temb=temb)
                if blk["tf"]:
                    x = self._transformer(blk["tf"][j], x, kvs[ti], B=B, F=F, H=H, W=W, n_ctx=n_ctx,
                                          batch_offset=batch_offset)
                    ti += 1
            if blk["up_w"] is not None:
                C = x.shape[1]
                up = self._empty(n_img * 4 * H * W, C)
                lib.upsample2x(x, up, n_img=n_img, H=H, W=W, C=C)
                H, W = 2 * H, 2 * W
                x = self._conv3(up, blk["up_w"], blk["up_b"], n_img=n_img, H=H, W=W, cin=C)
        rows = n_img * H * W
        y = self._gn(x, self.out_g, self.out_b, rows=rows, rows_per_inst=H * W, eps=1e-5, silu=True)
        return self._conv3(y, self.conv_out_w, self.conv_out_b, n_img=n_img, H=H, W=W, cin=x.shape[1], out_fp32=True)

    def zero_convs(self, skips, mid, scales: Sequence[float], into: Optional[Sequence[torch.Tensor]] = None,
                   mid_into: Optional[torch.Tensor] = None):
        """ControlNet 1x1 'zero' convs x conditioning scale (svd/temporal_controlnet.py:616-633). With `into`, the
        residual is accumulated straight into the UNet's skip tensors in the GEMM epilogue (U4 / K15)."""
        outs = []
        for i, s in enumerate(skips):
            rows = s.shape[0]
            if into is not None:
                outs.append(self._linear(s, self.zero_w[i], M=rows, bias=self.zero_b[i], s0=scales[i], res1=into[i],
                                         s1=1.0, out=into[i]))
            else:
                outs.append(self._linear(s, self.zero_w[i], M=rows, bias=self.zero_b[i], s0=scales[i]))
        rows = mid.shape[0]
        if mid_into is not None:
            m = self._linear(mid, self.zero_mid_w, M=rows, bias=self.zero_mid_b, s0=scales[-1], res1=mid_into, s1=1.0,
                             out=mid_into)
        else:
            m = self._linear(mid, self.zero_mid_w, M=rows, bias=self.zero_mid_b, s0=scales[-1])
        return outs, m

    # ============================================================================================ boundary API
    def _prep_inputs(self, sample, timestep, ehs, added_time_ids, extra_channels=None):
        if sample.device != self.device:
            raise lib.TtvdmError(f"sample on {sample.device}, model on {self.device}")
        B, F, Cin, H, W = sample.shape
        if H % 8 != 0 or W % 8 != 0:
            raise ValueError(f"latent height/width must be multiples of 8 (3 stride-2 levels), got {H}x{W}")
        if ehs.shape[0] != B:
            raise ValueError(f"encoder_hidden_states batch {ehs.shape[0]} != sample batch {B}")
        x = sample.reshape(B * F, Cin, H, W)
        if extra_channels is not None:
            x = torch.cat([x, extra_channels.to(x.dtype)], dim=1)
        x_in = torch.zeros(B * F, H, W, PAD_IN, dtype=BF16, device=self.device)
        x_in[..., : x.shape[1]] = x.permute(0, 2, 3, 1)
        if not torch.is_tensor(timestep):
            t = torch.tensor([float(timestep)], dtype=torch.float32, device=self.device)
        else:
            t = timestep.to(device=self.device, dtype=torch.float32).reshape(-1)
        t = t.expand(B).contiguous()
        ids = added_time_ids.to(device=self.device, dtype=torch.float32)
        if ids.shape[0] != B or ids.shape[1] * self.add_dim != self.ae_w1.shape[1]:
            raise ValueError(
                f"Model expects an added time embedding vector of length {self.ae_w1.shape[1]}, but a vector of "
                f"{ids.shape[1] * self.add_dim} was created. The model has an incorrect config.")
        self._ensure_pos_emb(F)
        temb = self.time_embeddings(t, ids)
        kvs = self.context_kv(ehs)
        return x_in.view(B * F * H * W, PAD_IN), temb, kvs, (B, F, H, W)

    @staticmethod
    def _to_nchw(x, n_img, H, W, dtype):
        return x.view(n_img, H, W, -1).permute(0, 3, 1, 2).to(dtype).contiguous()

    def _from_nchw(self, t, n_img, H, W):
        return t.to(device=self.device, dtype=BF16).permute(0, 2, 3, 1).reshape(n_img * H * W, -1).contiguous()

    def unet_forward(self, sample, timestep, ehs, added_time_ids, down_res=None, mid_res=None):
        assert self.kind == "unet"
        x_in, temb, kvs, (B, F, H, W) = self._prep_inputs(sample, timestep, ehs, added_time_ids)
        kw = dict(B=B, F=F, n_ctx=B, batch_offset=0)
        x, skips, dims, ti = self.encode(x_in, temb, kvs, H=H, W=W, **kw)
        hl, wl = dims[-1]
        x, ti = self.middle(x, temb, kvs, ti, H=hl, W=wl, **kw)
        if down_res is not None and mid_res is not None:
            # U4: skip_i += residual_i ; mid += residual (svd/unet_spatio_temporal_condition.py:481-502)
            new_skips = []
            for s, r, (hh, ww) in zip(skips, down_res, dims):
                rr = self._from_nchw(r, B * F, hh, ww)
                o = self._empty(*s.shape)
                lib.axpy(s, rr, o, 1.0, s.numel())
                new_skips.append(o)
            skips = new_skips
            rr = self._from_nchw(mid_res, B * F, hl, wl)
            lib.axpy(x, rr, x, 1.0, x.numel())
        eps = self.decode(x, skips, temb, kvs, ti, H=hl, W=wl, **kw)
        out = eps.view(B, F, H, W, self.out_channels).permute(0, 1, 4, 2, 3).to(sample.dtype).contiguous()
        return out

    def controlnet_forward(self, sample, timestep, ehs, added_time_ids, controlnet_cond, conditioning_scale=1.0,
                           guess_mode=False):
        assert self.kind == "controlnet"
        if controlnet_cond is None:
            raise ValueError("controlnet_cond is required")
        x_in, temb, kvs, (B, F, H, W) = self._prep_inputs(sample, timestep, ehs, added_time_ids,
                                                          extra_channels=controlnet_cond)
        kw = dict(B=B, F=F, n_ctx=B, batch_offset=0)
        x, skips, dims, ti = self.encode(x_in, temb, kvs, H=H, W=W, **kw)
        hl, wl = dims[-1]
        x, ti = self.middle(x, temb, kvs, ti, H=hl, W=wl, **kw)
        n = len(skips) + 1
        if guess_mode:
            scales = [float(v) * conditioning_scale for v in torch.logspace(-1, 0, n)]
        else:
            scales = [float(conditioning_scale)] * n
        outs, m = self.zero_convs(skips, x, scales)
        down = [self._to_nchw(o, B * F, hh, ww, sample.dtype) for o, (hh, ww) in zip(outs, dims)]
        return down, self._to_nchw(m, B * F, hl, wl, sample.dtype)
