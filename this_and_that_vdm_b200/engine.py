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
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib

BF16 = torch.bfloat16
MAX_FRAMES = 32  # kTaMaxF of csrc/attn_temporal.cu (two m16 MMA tiles of frames per warp: SVD-XT runs 25)
PAD_IN = 64  # conv_in channels padded to one 64-wide K chunk


def _src(t: torch.Tensor, dev) -> torch.Tensor:
    """Parameter as the repack kernels read it: on the device, contiguous, in its own dtype (fp32 / fp16 / bf16)."""
    t = t.detach()
    if t.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        t = t.float()
    return t.to(device=dev).contiguous()


def _bf(t: torch.Tensor, dev) -> torch.Tensor:
    """-> bf16 through ttvdm_pack_linear (plain cast, no fold): [N, ...] becomes [N, K]; vectors keep their shape."""
    t = _src(t, dev)
    src = t.reshape(1, -1) if t.ndim < 2 else t.reshape(t.shape[0], -1)
    out = torch.empty(src.shape, dtype=BF16, device=dev)
    lib.pack_linear(src, out)
    return out.view(t.shape) if t.ndim < 2 else out


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    t = _src(t, dev)
    out = torch.empty(t.shape, dtype=torch.float32, device=dev)
    lib.pack_vector(t, out)
    return out


def _pack_conv(w: torch.Tensor, dev, pad_cin: int = 0) -> torch.Tensor:
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin'] / [C, C, 3, 1, 1] -> [C, 3*C]: tap-major then channel (Cin' = Cin zero-padded
    to pad_cin), bf16 — ttvdm_pack_conv_weight."""
    w = _src(w, dev)
    cout, cin = w.shape[:2]
    taps = w[0, 0].numel()
    cp = max(pad_cin, cin)
    out = torch.empty(cout, taps * cp, dtype=BF16, device=dev)
    lib.pack_conv_weight(w, out, cin_pad=cp)
    return out


def _pack_upsample_parity(w: torch.Tensor, dev) -> List[torch.Tensor]:
    """Conv2d weight [Cout, Cin, 3, 3] of an Upsample2D -> four bf16 [Cout, 4 * Cin] operands, parity p = 2 * py + px:
    after nearest x2, output row 2y + py reads source rows {y - 1: ky 0, y: ky 1 + 2} (py = 0) or {y: ky 0 + 1, y + 1: ky 2}
    (py = 1), likewise for columns, so the 3x3 taps collapse into a 2x2 window with summed weights (summed in fp32 on the
    host, then packed like any conv weight)."""
    w = w.detach().to("cpu", torch.float32)
    groups = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}
    out = []
    for py in (0, 1):
        for px in (0, 1):
            wp = torch.zeros(w.shape[0], w.shape[1], 2, 2)
            for ty in (0, 1):
                for tx in (0, 1):
                    for ky in groups[py][ty]:
                        for kx in groups[px][tx]:
                            wp[:, :, ty, tx] += w[:, :, ky, kx]
            out.append(_pack_conv(wp, dev))
    return out


_pack_conv3x3 = _pack_conv
_pack_tconv = _pack_conv


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
    """Attention projections. The LayerNorm in front of to_q/k/v is FOLDED into the packed weights: W' = W diag(gamma)
    (bf16), bias' = W beta (fp32), colsum[n] = sum_k W'[n, k] (fp32, of the rounded W') — the GEMM epilogue then applies
    rstd * (acc - mean * colsum) + bias' from the producer's row sums (include/ttvdm.h, ln_rowsums)."""
    wqkv: Optional[torch.Tensor] = None  # self-attn fused [3C, C], LN-folded
    wq: Optional[torch.Tensor] = None    # cross-attn query [C, C], LN-folded
    w_b: Optional[torch.Tensor] = None   # folded bias of wqkv / wq
    w_cs: Optional[torch.Tensor] = None  # column sums of wqkv / wq
    wk: Optional[torch.Tensor] = None    # cross-attn [C, 1024]
    wv: Optional[torch.Tensor] = None
    wo: torch.Tensor = None
    bo: torch.Tensor = None
    ln: Tuple[torch.Tensor, torch.Tensor] = None  # (gamma, beta) fp32 of the LayerNorm in front (unfused path)


@dataclass
class FFW:
    w1: torch.Tensor = None   # GEGLU proj, (hidden, gate) rows interleaved, LN-folded
    b1: torch.Tensor = None   # folded bias
    cs1: torch.Tensor = None  # column sums of w1
    w2: torch.Tensor = None
    b2: torch.Tensor = None
    ln: Tuple[torch.Tensor, torch.Tensor] = None  # (gamma, beta) fp32 of the LayerNorm in front (unfused path)


@dataclass
class TfLayer:
    s_attn1: AttnW = None
    s_attn2: AttnW = None
    s_ff: FFW = None
    t_ff_in: FFW = None
    t_attn1: AttnW = None
    t_attn2: AttnW = None
    t_ff: FFW = None
    pos_prevec: torch.Tensor = None  # fp32 [F, 8C]: frame positional embedding pushed through t_ff_in.w1 (input independent)


@dataclass
class TfW:
    C: int
    heads: int

    @property
    def hd(self) -> int:  # head dim: 64 for every shipped config, 128 at level 2 of the reference's class-default heads
        return self.C // self.heads

    gn_g: torch.Tensor = None
    gn_b: torch.Tensor = None
    w_in: torch.Tensor = None
    b_in: torch.Tensor = None
    layers: List[TfLayer] = field(default_factory=list)
    pos_emb: torch.Tensor = None  # fp32 [F, C], input independent
    alpha: float = 0.5
    w_out: torch.Tensor = None
    b_out: torch.Tensor = None
    pos_w1: torch.Tensor = None
    pos_b1: torch.Tensor = None
    pos_w2: torch.Tensor = None
    pos_b2: torch.Tensor = None


def _fold_ln(ws: Sequence[torch.Tensor], b: Optional[torch.Tensor], gamma: Optional[torch.Tensor],
             beta: Optional[torch.Tensor], dev, geglu: bool = False):
    """Linear(cat(ws), b) packed for ttvdm_gemm, optionally with the LayerNorm(gamma, beta) in front of it folded in:
    (W' bf16, bias' fp32 or None, colsum fp32 or None) with Linear(LN(x)) = rstd * (x W'^T - mean * colsum) + bias'
    (ttvdm_pack_linear; geglu: (hidden_j, gate_j) row interleave). gamma = None: plain repack, no fold."""
    ws = [_src(w, dev) for w in ws]
    fold = gamma is not None
    if fold:
        gamma, beta = _src(gamma, dev).to(ws[0].dtype), _src(beta, dev).to(ws[0].dtype)
    N, K = sum(w.shape[0] for w in ws), ws[0].shape[1]
    wf = torch.empty(N, K, dtype=BF16, device=dev)
    bias = torch.empty(N, dtype=torch.float32, device=dev) if (fold or b is not None) else None
    cs = torch.empty(N, dtype=torch.float32, device=dev) if fold else None
    row0 = 0
    for w in ws:
        lib.pack_linear(w, wf, bias=None if b is None else _src(b, dev).to(w.dtype), gamma=gamma if fold else None,
                        beta=beta if fold else None, out_bias=bias, out_colsum=cs, geglu=geglu, out_row0=row0)
        row0 += w.shape[0]
    return wf, bias, cs


def _sigmoid_scalar(t: torch.Tensor) -> float:
    return 1.0 / (1.0 + math.exp(-float(t.detach().float().cpu().reshape(-1)[0])))


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
            if c % 64 != 0 or c % h != 0 or c // h not in (64, 128):
                raise lib.TtvdmError(
                    f"sm_100a engine supports head_dim 64 (and 128 on the compatibility path) with channels %% 64 == 0 "
                    f"(got C={c}, heads={h})")
        self.temb_dim = self.chans[0] * 4
        self.add_dim = cfg.addition_time_embed_dim
        self._pos_cache: Dict[int, bool] = {}
        self._pos_rep: Dict[Tuple[int, int, int], torch.Tensor] = {}
        # statistics pool: every GroupNorm / LayerNorm sum a GEMM epilogue accumulates during one forward lives here and
        # is zeroed by ONE memset at the start of the forward (begin_step)
        self._pool: Optional[torch.Tensor] = None
        self._pool_used = 0
        self._pool_high = 0
        self.fuse_norm_stats = True  # False: standalone GroupNorm statistics pass (A/B switch for profiling / tests)
        # LayerNorm folded into the consuming GEMM (row sums from the producer's epilogue). Built, tested and measured
        # (profiles/r02_fusion_costs.json): on B200 the extra epilogue work costs more than the layernorm kernel it
        # removes at every level, so it is OFF unless TTVDM_FUSE_LN=1 — the weights are packed accordingly.
        self.fuse_layernorm = os.environ.get("TTVDM_FUSE_LN", "0") == "1"
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
                r.wsc = _bf(g(sp + ".conv_shortcut.weight"), dev)
                r.bsc = _f32(g(sp + ".conv_shortcut.bias"), dev)
            r.tn1_g, r.tn1_b = _f32(g(tp + ".norm1.weight"), dev), _f32(g(tp + ".norm1.bias"), dev)
            r.tw1, r.tb1 = _pack_tconv(g(tp + ".conv1.weight"), dev), _f32(g(tp + ".conv1.bias"), dev)
            r.tn2_g, r.tn2_b = _f32(g(tp + ".norm2.weight"), dev), _f32(g(tp + ".norm2.bias"), dev)
            r.tw2, r.tb2 = _pack_tconv(g(tp + ".conv2.weight"), dev), _f32(g(tp + ".conv2.bias"), dev)
            r.alpha = _sigmoid_scalar(g(prefix + ".time_mixer.mix_factor"))
            # all time_emb_proj layers are fused into ONE [sum C, 1280] GEMM per forward (K13 hoist)
            r.temb_off_s = self.temb_cols
            temb_w.append(g(sp + ".time_emb_proj.weight")); temb_b.append(g(sp + ".time_emb_proj.bias"))
            self.temb_cols += cout
            r.temb_off_t = self.temb_cols
            temb_w.append(g(tp + ".time_emb_proj.weight")); temb_b.append(g(tp + ".time_emb_proj.bias"))
            self.temb_cols += cout
            return r

        def attn(prefix: str, cross: bool, ln: str) -> AttnW:
            a = AttnW()
            gamma, beta = (g(ln + ".weight"), g(ln + ".bias")) if self.fuse_layernorm else (None, None)
            a.ln = (_f32(g(ln + ".weight"), dev), _f32(g(ln + ".bias"), dev))
            if cross:
                a.wq, a.w_b, a.w_cs = _fold_ln([g(prefix + ".to_q.weight")], None, gamma, beta, dev)
                a.wk = _bf(g(prefix + ".to_k.weight"), dev)
                a.wv = _bf(g(prefix + ".to_v.weight"), dev)
            else:
                a.wqkv, a.w_b, a.w_cs = _fold_ln([g(prefix + ".to_q.weight"), g(prefix + ".to_k.weight"),
                                                  g(prefix + ".to_v.weight")], None, gamma, beta, dev)
            a.wo, a.bo = _bf(g(prefix + ".to_out.0.weight"), dev), _f32(g(prefix + ".to_out.0.bias"), dev)
            return a

        def ff(prefix: str, ln: str) -> FFW:
            f = FFW()
            # GEGLU rows interleaved (hidden_j, gate_j) by the repack kernel: weights, bias and column sums stay aligned
            gamma, beta = (g(ln + ".weight"), g(ln + ".bias")) if self.fuse_layernorm else (None, None)
            f.ln = (_f32(g(ln + ".weight"), dev), _f32(g(ln + ".bias"), dev))
            f.w1, f.b1, f.cs1 = _fold_ln([g(prefix + ".net.0.proj.weight")], g(prefix + ".net.0.proj.bias"), gamma, beta, dev,
                                         geglu=True)
            f.w2, f.b2 = _bf(g(prefix + ".net.2.weight"), dev), _f32(g(prefix + ".net.2.bias"), dev)
            return f

        def tf(prefix: str, heads: int) -> TfW:
            C = g(prefix + ".proj_in.weight").shape[0]
            t = TfW(C=C, heads=heads)
            t.gn_g, t.gn_b = _f32(g(prefix + ".norm.weight"), dev), _f32(g(prefix + ".norm.bias"), dev)
            t.w_in, t.b_in = _bf(g(prefix + ".proj_in.weight"), dev), _f32(g(prefix + ".proj_in.bias"), dev)
            i = 0
            while (prefix + f".transformer_blocks.{i}.norm1.weight") in sd:  # transformer_layers_per_block >= 1
                sb, tb = prefix + f".transformer_blocks.{i}", prefix + f".temporal_transformer_blocks.{i}"
                L = TfLayer()
                L.s_attn1 = attn(sb + ".attn1", False, sb + ".norm1")
                L.s_attn2 = attn(sb + ".attn2", True, sb + ".norm2")
                L.s_ff = ff(sb + ".ff", sb + ".norm3")
                L.t_ff_in = ff(tb + ".ff_in", tb + ".norm_in")
                L.t_attn1 = attn(tb + ".attn1", False, tb + ".norm1")
                L.t_attn2 = attn(tb + ".attn2", True, tb + ".norm2")
                L.t_ff = ff(tb + ".ff", tb + ".norm3")
                t.layers.append(L)
                i += 1
            t.alpha = _sigmoid_scalar(g(prefix + ".time_mixer.mix_factor"))
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
                    blk["up_wp"] = _pack_upsample_parity(g(f"{p}.upsamplers.0.conv.weight"), dev)
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
                self.zero_w.append(_bf(w, dev))
                self.zero_b.append(_f32(g(f"controlnet_down_blocks.{i}.bias"), dev))
                i += 1
            w = g("controlnet_mid_block.weight")
            self.zero_mid_w = _bf(w, dev)
            self.zero_mid_b = _f32(g("controlnet_mid_block.bias"), dev)
        # ---- fused time_emb_proj
        self.temb_w = torch.empty(self.temb_cols, temb_w[0].shape[1], dtype=BF16, device=dev)
        self.temb_b = torch.empty(self.temb_cols, dtype=torch.float32, device=dev)
        row0 = 0
        for w, b in zip(temb_w, temb_b):
            lib.pack_linear(_src(w, dev), self.temb_w, out_row0=row0)
            lib.pack_vector(_src(b, dev), self.temb_b[row0:row0 + w.shape[0]])
            row0 += w.shape[0]

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

    # ---- statistics pool
    def begin_step(self) -> None:
        """Zero the statistics pool (one memset) and rewind it. Called once per network forward."""
        need = self._pool_high
        if self._pool is None or need > self._pool.numel():
            self._pool = torch.zeros(max(need + (need >> 2), 1 << 20), dtype=torch.uint8, device=self.device)
        else:
            self._pool[: max(need, 16)].zero_()
        self._pool_used = 0

    def _stat(self, numel: int, dtype) -> torch.Tensor:
        """`numel` zeroed elements of the step's statistics pool (falls back to a fresh zero tensor while the pool's
        size is still being learned: first forward of a new shape)."""
        nbytes = numel * (8 if dtype == torch.float64 else 4)
        off = (self._pool_used + 15) & ~15
        self._pool_used = off + nbytes
        self._pool_high = max(self._pool_high, self._pool_used)
        if self._pool is None or self._pool_used > self._pool.numel():
            return torch.zeros(numel, dtype=dtype, device=self.device)
        return self._pool[off:off + nbytes].view(dtype)

    def _norm_out_kw(self, out: torch.Tensor, M: int, N: int, gn_rpi: int, ln_out: bool, rs_add) -> dict:
        """Epilogue statistics requests for a GEMM that writes `out` [M, N]: GroupNorm pair sums for a consumer whose
        group instance spans gn_rpi rows, and / or LayerNorm row sums. The buffers ride on the tensor object."""
        kw = {}
        out.gn_stats = None
        out.ln_sums = None
        if gn_rpi and self.fuse_norm_stats and N % 64 == 0 and M % gn_rpi == 0:
            st = self._stat((M // gn_rpi) * N, torch.float64)
            kw.update(gn_stats_out=st, gn_rows_per_inst=gn_rpi)
            out.gn_stats = (st, gn_rpi)
        if ln_out:
            # [N / 32, M, 2] partial sums: one writer per slot, so plain (un-zeroed) scratch is enough
            rs = self._empty((N // 32) * M * 2, dtype=torch.float32)
            kw.update(row_sums_out=rs)
            if rs_add is not None:
                kw.update(rs_addvec=rs_add[0], rs_add_rows=rs_add[1], rs_add_mod=rs_add[2])
            out.ln_sums = rs
        return kw

    def _linear(self, a, w, *, M, bias=None, out=None, res1=None, s1=1.0, res2=None, s2=1.0, s0=1.0, geglu=False,
                a2=None, k2=0, act=0, out_fp32=False, lda=None, rowvec=None, rows_per_vec=0, ldrv=0, gn_rpi=0,
                ln_out=False, rs_add=None, ln=None, prevec=None):
        """ln = (row sums of `a`, column sums of `w`): `a` is the UN-normalised tensor and `w` carries the folded
        LayerNorm gain. prevec = (table [mod, N], rows, mod)."""
        N, K = w.shape[0], w.shape[1] - k2
        if out is None:
            out = self._empty(M, N // 2 if geglu else N, dtype=torch.float32 if out_fp32 else BF16)
        kw = self._norm_out_kw(out, M, N, gn_rpi, ln_out, rs_add) if not (geglu or out_fp32) else {}
        if ln is not None:
            kw.update(ln_rowsums=ln[0], ln_colsum=ln[1], ln_eps=1e-5)
            if prevec is not None:
                kw.update(prevec=prevec[0], prevec_rows=prevec[1], prevec_mod=prevec[2], ldpv=N)
        lib.gemm(a, w, out, M=M, N=N, k1=K, a2=a2, k2=k2, bias=bias, res1=res1, s1=s1, res2=res2, s2=s2, s0=s0,
                 geglu=geglu, act=act, out_fp32=out_fp32, lda=lda, rowvec=rowvec, rows_per_vec=rows_per_vec, ldrv=ldrv,
                 **kw)
        return out

    def _conv3(self, x, w, bias, *, n_img, H, W, cin, out=None, rowvec=None, rows_per_vec=0, ldrv=0, res1=None,
               out_fp32=False, gn_rpi=0, stride=1):
        """3x3 convolution, padding 1. H, W are the OUTPUT dims (stride 2: the input is [n_img, 2H, 2W, cin])."""
        N = w.shape[0]
        M = n_img * H * W
        if out is None:
            out = self._empty(M, N, dtype=torch.float32 if out_fp32 else BF16)
        kw = self._norm_out_kw(out, M, N, gn_rpi, False, None) if not out_fp32 else {}
        lib.gemm(x, w, out, M=M, N=N, k1=cin, mode=lib.A_CONV3X3, n_img=n_img, H=H, W=W, bias=bias, rowvec=rowvec,
                 rows_per_vec=rows_per_vec, ldrv=ldrv, res1=res1, out_fp32=out_fp32, conv_stride=stride, **kw)
        return out

    def _upsample_conv(self, x, blk, *, n_img, H, W):
        """Upsample2D = nearest x2 + 3x3 conv, WITHOUT the upsampled tensor: every output parity (py, px) is a 2x2-tap
        convolution of the LOW-resolution input with pre-summed weights (_pack_upsample_parity) — 16 instead of 36
        tap-pixels — and ttvdm_interleave2x puts the four results into place. The GroupNorm sums of the next ResBlock are
        accumulated by the four epilogues into one buffer (an instance is a frame either way)."""
        C = x.shape[1]
        N = blk["up_wp"][0].shape[0]
        M = n_img * H * W
        out = self._empty(4 * M, N)
        parts = self._empty(4, M, N)
        out.ln_sums = None
        out.gn_stats = None
        kw = {}
        if self.fuse_norm_stats and N % 64 == 0:
            st = self._stat(n_img * N, torch.float64)
            kw = dict(gn_stats_out=st, gn_rows_per_inst=H * W)
            out.gn_stats = (st, 4 * H * W)
        for pi, (py, px) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            lib.gemm(x, blk["up_wp"][pi], parts[pi], M=M, N=N, k1=C, mode=lib.A_CONV3X3, n_img=n_img, H=H, W=W,
                     bias=blk["up_b"], conv_taps=4, conv_dy0=py - 1, conv_dx0=px - 1, **kw)
        lib.interleave2x(parts, out, n_img=n_img, H=H, W=W, C=N)
        return out

    def _tconv(self, x, w, bias, *, B, F, S, C, out=None, rowvec=None, rows_per_vec=0, ldrv=0, s0=1.0, res1=None,
               s1=1.0, res2=None, s2=1.0, gn_rpi=0):
        M = B * F * S
        if out is None:
            out = self._empty(M, C)
        kw = self._norm_out_kw(out, M, C, gn_rpi, False, None)
        lib.gemm(x, w, out, M=M, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=bias, rowvec=rowvec,
                 rows_per_vec=rows_per_vec, ldrv=ldrv, s0=s0, res1=res1, s1=s1, res2=res2, s2=s2, **kw)
        return out

    def _ln(self, x, gb, *, rows, C, addvec=None, F=0, S=0, sum_out=None, out=None):
        if out is None:
            out = self._empty(rows, C)
        lib.layernorm(x, out, gb[0], gb[1], rows=rows, C=C, addvec=addvec, F=F, S=S, sum_out=sum_out)
        return out

    def _gn(self, x, gamma, beta, *, rows, rows_per_inst, eps, silu, x2=None):
        """GroupNorm(32) (+SiLU, + channel concat with x2). Sources whose producing GEMM left per-pair sums for this
        instance size (tensor.gn_stats) skip the statistics pass."""
        c1 = x.shape[1]
        c2 = x2.shape[1] if x2 is not None else 0
        out = self._empty(rows, c1 + c2)

        def pst(t):
            st = getattr(t, "gn_stats", None) if t is not None else None
            return st[0] if (st is not None and st[1] == rows_per_inst and ((c1 + c2) // 32) % 2 == 0) else None

        ps1, ps2 = pst(x), pst(x2)
        stats = None
        if ps1 is None or (x2 is not None and ps2 is None):
            stats = self._empty((rows // rows_per_inst) * 64, dtype=torch.float64)
        lib.groupnorm(x, out, stats, gamma, beta, c1=c1, rows=rows, rows_per_inst=rows_per_inst, eps=eps, silu=silu,
                      x2=x2, c2=c2, pstats1=ps1, pstats2=ps2)
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

    def context_kv(self, ehs: torch.Tensor, out: Optional[list] = None) -> List[List[Tuple[torch.Tensor, ...]]]:
        """Cross-attention K/V of the (constant) context for every transformer layer: computed once per video instead of
        per frame / per pixel (reference: svd/unet_spatio_temporal_condition.py:452 repeat_interleave,
        svd/diffusion_arch/transformer_temporal.py:316-319 broadcast). ehs [B, L, D] -> per transformer, per layer
        (k_s, v_s, k_t, v_t), each bf16 [B, L, C]. `out`: a previous result of the same shape to overwrite in place (the
        captured step graph keeps reading the same buffers for the next video)."""
        B, L, D = ehs.shape
        x = _bf(ehs.reshape(B * L, D), self.device)
        res = []
        for ti, t in enumerate(self.all_transformers()):
            per_layer = []
            for li, ly in enumerate(t.layers):
                old = out[ti][li] if out is not None else (None, None, None, None)
                ks = self._linear(x, ly.s_attn2.wk, M=B * L, out=old[0])
                vs = self._linear(x, ly.s_attn2.wv, M=B * L, out=old[1])
                kt = self._linear(x, ly.t_attn2.wk, M=B * L, out=old[2])
                vt = self._linear(x, ly.t_attn2.wv, M=B * L, out=old[3])
                per_layer.append((ks, vs, kt, vt))
            res.append(per_layer)
        return res

    def _ensure_pos_emb(self, F: int) -> None:
        """time_pos_embed(time_proj(arange(F))) is input independent (transformer_temporal.py:328-339): once. So is its
        image under the (LayerNorm-folded) ff_in projection, which the GEGLU GEMM adds before the LayerNorm scale."""
        if F in self._pos_cache:
            return
        frames = torch.arange(F, device=self.device, dtype=torch.float32)
        for t in self.all_transformers():
            ts = self._empty(F, t.C)
            lib.sinusoid(frames, ts, n=F, dim=t.C)
            h = self._linear(ts, t.pos_w1, M=F, bias=t.pos_b1, act=1)
            t.pos_emb = self._linear(h, t.pos_w2, M=F, bias=t.pos_b2, out_fp32=True)
            if self.fuse_layernorm:
                pe = _bf(t.pos_emb, self.device)
                for ly in t.layers:
                    ly.pos_prevec = self._linear(pe, ly.t_ff_in.w1, M=F, out_fp32=True)
        self._pos_cache = {F: True}
        self._pos_rep = {}

    def _pos_table(self, t: TfW, B: int, F: int) -> torch.Tensor:
        """fp32 [B*F, C]: the frame positional embedding as a per-(batch, frame) row vector (rowvec index = row / S)."""
        key = (id(t), B, F)
        if key not in self._pos_rep:
            self._pos_rep[key] = t.pos_emb.repeat(B, 1).contiguous()
        return self._pos_rep[key]

    # ============================================================================================ blocks
    def _resblock(self, r: ResW, x, skip, *, B, F, H, W, temb, out=None, out_res2=None, out_s2=0.0):
        """SpatioTemporalResBlock (diffusers; A.3-A.6). x [rows, c], skip optional second source (channel concat).
        temb fp32 [B, sum C]. Returns [rows, cout] carrying per-frame GroupNorm sums for its consumer.
        Every GroupNorm here reads statistics that the producing GEMM's epilogue accumulated (gn_rpi = rows per group
        instance of the CONSUMER: S for the 4-D norms, F*S for the 5-D temporal ones); only the apply pass remains."""
        S = H * W
        rows = B * F * S
        n_img = B * F
        ldrv = self.temb_cols
        y = self._gn(x, r.n1_g, r.n1_b, rows=rows, rows_per_inst=S, eps=r.eps, silu=True, x2=skip)
        h = self._conv3(y, r.w1, r.b1, n_img=n_img, H=H, W=W, cin=r.cin, rowvec=temb[:, r.temb_off_s:],
                        rows_per_vec=F * S, ldrv=ldrv, gn_rpi=S)
        y = self._gn(h, r.n2_g, r.n2_b, rows=rows, rows_per_inst=S, eps=r.eps, silu=True)
        if r.wsc is not None:
            if skip is not None:
                sc = self._linear(x, r.wsc, M=rows, bias=r.bsc, a2=skip, k2=skip.shape[1])
            else:
                sc = self._linear(x, r.wsc, M=rows, bias=r.bsc)
        else:
            sc = x
        hs = self._conv3(y, r.w2, r.b2, n_img=n_img, H=H, W=W, cin=r.cout, res1=sc, out=h, gn_rpi=F * S)
        # temporal branch: 5-D GroupNorm (stats over all frames of a video) + 3-tap conv over frames
        y = self._gn(hs, r.tn1_g, r.tn1_b, rows=rows, rows_per_inst=F * S, eps=r.eps, silu=True)
        t1 = self._tconv(y, r.tw1, r.tb1, B=B, F=F, S=S, C=r.cout, rowvec=temb[:, r.temb_off_t:], rows_per_vec=F * S,
                         ldrv=ldrv, gn_rpi=F * S)
        y = self._gn(t1, r.tn2_g, r.tn2_b, rows=rows, rows_per_inst=F * S, eps=r.eps, silu=True)
        # AlphaBlender: a*hs + (1-a)*(hs + conv) = hs + (1-a)*conv ; optional fused extra residual (ControlNet)
        if out is None:
            out = t1
        return self._tconv(y, r.tw2, r.tb2, B=B, F=F, S=S, C=r.cout, out=out, s0=1.0 - r.alpha, res1=hs, s1=1.0,
                           res2=out_res2, s2=out_s2, gn_rpi=S)

    def _self_attn_hd128(self, y, wqkv, *, n_img, S, C, heads, scale):
        """Spatial self-attention for head_dim 128 (level 2 of the reference UNet's CLASS-DEFAULT heads (5, 10, 10, 20),
        svd/unet_spatio_temporal_condition.py:99 — no shipped config uses it): compatibility path out of the GEMM and
        row-softmax kernels, one (image, head) at a time — Q K^T (fp32 scores) -> softmax -> P V with V^T produced by a
        swapped-operand GEMM per image. Correct, all sm_100a kernels, not tuned (the flash kernel's TMEM plan is d = 64)."""
        d = 128
        rows = n_img * S
        Sp = (S + 63) // 64 * 64
        wq, wk, wv = wqkv[:C], wqkv[C:2 * C], wqkv[2 * C:]
        q = self._linear(y, wq, M=rows)
        kh = self._empty(heads, rows, d)
        for h in range(heads):  # K per head, contiguous [rows, 128]: the W operand of the score GEMM
            lib.gemm(y, wk[h * d:(h + 1) * d], kh[h], M=rows, N=d, k1=C)
        vt = torch.zeros(C, Sp, dtype=BF16, device=self.device)  # columns >= S stay zero (K padding of P V)
        scores = self._empty(S, S, dtype=torch.float32)
        probs = self._empty(S, Sp)
        o = self._empty(rows, C)
        for n in range(n_img):
            yn = y[n * S:(n + 1) * S]
            lib.gemm(wv, yn, vt, M=C, N=S, k1=C, ldo=Sp)  # V^T of all heads of this image: [C, Sp]
            for h in range(heads):
                lib.gemm(q[n * S:, h * d:], kh[h][n * S:(n + 1) * S], scores, M=S, N=S, k1=d, lda=C, s0=scale,
                         out_fp32=True)
                lib.softmax_rows(scores, probs, rows=S, cols=S, ldx=S, ldo=Sp, cols_out=Sp)
                lib.gemm(probs, vt[h * d:(h + 1) * d], o[n * S:, h * d:], M=S, N=d, k1=Sp, ldo=C)
        return o

    def _transformer(self, t: TfW, x, kv, *, B, F, H, W, n_ctx, batch_offset):
        """TransformerSpatioTemporalModel.forward (svd/diffusion_arch/transformer_temporal.py:276-381)."""
        if self.fuse_layernorm and t.hd == 64:
            return self._transformer_ln_folded(t, x, kv, B=B, F=F, H=H, W=W, n_ctx=n_ctx, batch_offset=batch_offset)
        S = H * W
        rows = B * F * S
        C = t.C
        scale = float(t.hd) ** -0.5
        hd = t.hd
        y = self._gn(x, t.gn_g, t.gn_b, rows=rows, rows_per_inst=S, eps=1e-6, silu=False)
        h = self._linear(y, t.w_in, M=rows, bias=t.b_in)
        qkv = q = gg = hm = None
        for li, (ly, (ks, vs, kt, vt)) in enumerate(zip(t.layers, kv)):
            L = ks.shape[0] // n_ctx
            # ---- spatial BasicTransformerBlock
            y = self._ln(h, ly.s_attn1.ln, rows=rows, C=C, out=y)
            if hd == 64:
                qkv = self._linear(y, ly.s_attn1.wqkv, M=rows, out=qkv)
                o = y  # reuse
                lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], o, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=B * F,
                                 heads=t.heads, seq=S, scale=scale)
            else:
                o = self._self_attn_hd128(y, ly.s_attn1.wqkv, n_img=B * F, S=S, C=C, heads=t.heads, scale=scale)
            self._linear(o, ly.s_attn1.wo, M=rows, bias=ly.s_attn1.bo, res1=h, out=h)
            y = self._ln(h, ly.s_attn2.ln, rows=rows, C=C, out=y)
            q = self._linear(y, ly.s_attn2.wq, M=rows, out=q)
            lib.attn_cross(q, ks, vs, y, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                           temporal=False, batch_offset=batch_offset, scale=scale, head_dim=hd)
            self._linear(y, ly.s_attn2.wo, M=rows, bias=ly.s_attn2.bo, res1=h, out=h)
            y = self._ln(h, ly.s_ff.ln, rows=rows, C=C, out=y)
            gg = self._linear(y, ly.s_ff.w1, M=rows, bias=ly.s_ff.b1, geglu=True, out=gg)
            self._linear(gg, ly.s_ff.w2, M=rows, bias=ly.s_ff.b2, res1=h, out=h)  # h == x_spatial
            # ---- TemporalBasicTransformerBlock on (h + frame positional embedding); rows stay (b, f, s)
            if hm is None:
                hm = self._empty(rows, C)
            y = self._ln(h, ly.t_ff_in.ln, rows=rows, C=C, addvec=t.pos_emb, F=F, S=S, sum_out=hm, out=y)
            gg = self._linear(y, ly.t_ff_in.w1, M=rows, bias=ly.t_ff_in.b1, geglu=True, out=gg)
            self._linear(gg, ly.t_ff_in.w2, M=rows, bias=ly.t_ff_in.b2, res1=hm, out=hm)
            y = self._ln(hm, ly.t_attn1.ln, rows=rows, C=C, out=y)
            qkv = self._linear(y, ly.t_attn1.wqkv, M=rows, out=qkv)
            lib.attn_temporal(qkv, qkv[:, C:], qkv[:, 2 * C:], y, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, B=B, F=F, S=S,
                              heads=t.heads, scale=scale, head_dim=hd)
            self._linear(y, ly.t_attn1.wo, M=rows, bias=ly.t_attn1.bo, res1=hm, out=hm)
            y = self._ln(hm, ly.t_attn2.ln, rows=rows, C=C, out=y)
            q = self._linear(y, ly.t_attn2.wq, M=rows, out=q)
            lib.attn_cross(q, kt, vt, y, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                           temporal=True, batch_offset=batch_offset, scale=scale, head_dim=hd)
            self._linear(y, ly.t_attn2.wo, M=rows, bias=ly.t_attn2.bo, res1=hm, out=hm)
            y = self._ln(hm, ly.t_ff.ln, rows=rows, C=C, out=y)
            gg = self._linear(y, ly.t_ff.w1, M=rows, bias=ly.t_ff.b1, geglu=True, out=gg)
            # AlphaBlender fused: a*h + (1-a)*(ff + hm)
            a = t.alpha
            self._linear(gg, ly.t_ff.w2, M=rows, bias=ly.t_ff.b2, s0=1.0 - a, res1=hm, s1=1.0 - a, res2=h, s2=a, out=h)
        # ---- proj_out + input residual (its epilogue leaves the GroupNorm sums of the next ResBlock)
        if hm is None:
            hm = self._empty(rows, C)
        return self._linear(h, t.w_out, M=rows, bias=t.b_out, res1=x, out=hm, gn_rpi=S)

    def _transformer_ln_folded(self, t: TfW, x, kv, *, B, F, H, W, n_ctx, batch_offset):
        """TransformerSpatioTemporalModel.forward with every LayerNorm folded into its consumer (TTVDM_FUSE_LN=1).
        No LayerNorm kernel runs: each of the 7 LayerNorms per layer is folded into the GEMM that consumes it (folded
        weights + epilogue scale from the row sums the producing GEMM's epilogue accumulated), and `hidden + frame
        positional embedding` (:356) is never materialised — the embedding enters the statistics (rs_add), the ff_in
        projection (prevec) and the ff_in residual (rowvec) separately."""
        S = H * W
        rows = B * F * S
        C = t.C
        scale = 0.125
        y = self._gn(x, t.gn_g, t.gn_b, rows=rows, rows_per_inst=S, eps=1e-6, silu=False)
        h = self._linear(y, t.w_in, M=rows, bias=t.b_in, ln_out=True)
        o = y  # attention outputs reuse the normalised-input buffer
        qkv = q = gg = hm = None
        pos_rs = (t.pos_emb, S, F)
        pos_tab = self._pos_table(t, B, F)
        for li, (ly, (ks, vs, kt, vt)) in enumerate(zip(t.layers, kv)):
            L = ks.shape[0] // n_ctx
            last = li == len(t.layers) - 1
            # ---- spatial BasicTransformerBlock
            qkv = self._linear(h, ly.s_attn1.wqkv, M=rows, bias=ly.s_attn1.w_b, ln=(h.ln_sums, ly.s_attn1.w_cs), out=qkv)
            lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], o, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=B * F,
                             heads=t.heads, seq=S, scale=scale)
            self._linear(o, ly.s_attn1.wo, M=rows, bias=ly.s_attn1.bo, res1=h, out=h, ln_out=True)
            q = self._linear(h, ly.s_attn2.wq, M=rows, bias=ly.s_attn2.w_b, ln=(h.ln_sums, ly.s_attn2.w_cs), out=q)
            lib.attn_cross(q, ks, vs, o, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                           temporal=False, batch_offset=batch_offset, scale=scale)
            self._linear(o, ly.s_attn2.wo, M=rows, bias=ly.s_attn2.bo, res1=h, out=h, ln_out=True)
            gg = self._linear(h, ly.s_ff.w1, M=rows, bias=ly.s_ff.b1, geglu=True, ln=(h.ln_sums, ly.s_ff.cs1), out=gg)
            # h == x_spatial; its row sums are taken over (h + frame positional embedding) for norm_in
            self._linear(gg, ly.s_ff.w2, M=rows, bias=ly.s_ff.b2, res1=h, out=h, ln_out=True, rs_add=pos_rs)
            # ---- TemporalBasicTransformerBlock on (h + pos); rows stay (b, f, s)
            gg = self._linear(h, ly.t_ff_in.w1, M=rows, bias=ly.t_ff_in.b1, geglu=True, ln=(h.ln_sums, ly.t_ff_in.cs1),
                              prevec=(ly.pos_prevec, S, F), out=gg)
            hm = self._linear(gg, ly.t_ff_in.w2, M=rows, bias=ly.t_ff_in.b2, res1=h, rowvec=pos_tab, rows_per_vec=S,
                              ldrv=C, out=hm, ln_out=True)
            qkv = self._linear(hm, ly.t_attn1.wqkv, M=rows, bias=ly.t_attn1.w_b, ln=(hm.ln_sums, ly.t_attn1.w_cs), out=qkv)
            lib.attn_temporal(qkv, qkv[:, C:], qkv[:, 2 * C:], o, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, B=B, F=F, S=S,
                              heads=t.heads, scale=scale)
            self._linear(o, ly.t_attn1.wo, M=rows, bias=ly.t_attn1.bo, res1=hm, out=hm, ln_out=True)
            q = self._linear(hm, ly.t_attn2.wq, M=rows, bias=ly.t_attn2.w_b, ln=(hm.ln_sums, ly.t_attn2.w_cs), out=q)
            lib.attn_cross(q, kt, vt, o, ldq=C, ldo=C, rows=rows, heads=t.heads, L=L, F=F, S=S, n_ctx=n_ctx,
                           temporal=True, batch_offset=batch_offset, scale=scale)
            self._linear(o, ly.t_attn2.wo, M=rows, bias=ly.t_attn2.bo, res1=hm, out=hm, ln_out=True)
            gg = self._linear(hm, ly.t_ff.w1, M=rows, bias=ly.t_ff.b1, geglu=True, ln=(hm.ln_sums, ly.t_ff.cs1), out=gg)
            # AlphaBlender fused: a*h + (1-a)*(ff + hm); the next layer's norm1 needs the row sums of the blend
            a = t.alpha
            self._linear(gg, ly.t_ff.w2, M=rows, bias=ly.t_ff.b2, s0=1.0 - a, res1=hm, s1=1.0 - a, res2=h, s2=a, out=h,
                         ln_out=not last)
        # ---- proj_out + input residual
        return self._linear(h, t.w_out, M=rows, bias=t.b_out, res1=x, out=hm, gn_rpi=S)

    # ============================================================================================ network halves
    def encode(self, x_in, temb, kvs, *, B, F, H, W, n_ctx, batch_offset):
        """conv_in + down blocks. Returns (x, skips[12], dims[12], ti) — shared by UNet and ControlNet."""
        n_img = B * F
        x = self._conv3(x_in, self.conv_in_w, self.conv_in_b, n_img=n_img, H=H, W=W, cin=PAD_IN, gn_rpi=H * W)
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
                # Downsample2D: 3x3 conv, stride 2, padding 1 — implicit GEMM over a TMA box with element stride 2 (round 2;
                # round 1 gathered a 9x im2col buffer first)
                H, W = H // 2, W // 2
                if (2 * H, 2 * W) != dims[-1]:
                    raise lib.TtvdmError(f"Downsample2D needs even height / width, got {dims[-1]}")
                x = self._conv3(x, blk["down_w"], blk["down_b"], n_img=n_img, H=H, W=W, cin=x.shape[1], gn_rpi=H * W,
                                stride=2)
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
                if os.environ.get("TTVDM_UPSAMPLE_PARITY", "1") != "0":
                    x = self._upsample_conv(x, blk, n_img=n_img, H=H, W=W)
                    H, W = 2 * H, 2 * W
                else:  # A/B: materialise the upsampled tensor, then the 3x3 conv (round 1's schedule)
                    C = x.shape[1]
                    up = self._empty(n_img * 4 * H * W, C)
                    lib.upsample2x(x, up, n_img=n_img, H=H, W=W, C=C)
                    H, W = 2 * H, 2 * W
                    x = self._conv3(up, blk["up_w"], blk["up_b"], n_img=n_img, H=H, W=W, cin=C, gn_rpi=H * W)
        rows = n_img * H * W
        y = self._gn(x, self.out_g, self.out_b, rows=rows, rows_per_inst=H * W, eps=1e-5, silu=True)
        return self._conv3(y, self.conv_out_w, self.conv_out_b, n_img=n_img, H=H, W=W, cin=x.shape[1], out_fp32=True)

    def zero_convs(self, skips, mid, scales: Sequence[float], into: Optional[Sequence[torch.Tensor]] = None,
                   mid_into: Optional[torch.Tensor] = None, n_img: int = 0):
        """ControlNet 1x1 'zero' convs x conditioning scale (svd/temporal_controlnet.py:616-633). With `into`, the
        residual is accumulated straight into the UNet's skip tensors in the GEMM epilogue (U4 / K15)."""
        outs = []
        for i, s in enumerate(skips):
            rows = s.shape[0]
            if into is not None:
                # the merged skip feeds an up-block GroupNorm: its per-frame sums come out of this epilogue
                rpi = rows // n_img if n_img else 0
                outs.append(self._linear(s, self.zero_w[i], M=rows, bias=self.zero_b[i], s0=scales[i], res1=into[i],
                                         s1=1.0, out=into[i], gn_rpi=rpi))
            else:
                outs.append(self._linear(s, self.zero_w[i], M=rows, bias=self.zero_b[i], s0=scales[i]))
        rows = mid.shape[0]
        if mid_into is not None:
            m = self._linear(mid, self.zero_mid_w, M=rows, bias=self.zero_mid_b, s0=scales[-1], res1=mid_into, s1=1.0,
                             out=mid_into, gn_rpi=rows // n_img if n_img else 0)
        else:
            m = self._linear(mid, self.zero_mid_w, M=rows, bias=self.zero_mid_b, s0=scales[-1])
        return outs, m

    # ============================================================================================ boundary API
    def _prep_inputs(self, sample, timestep, ehs, added_time_ids, extra_channels=None):
        if sample.device != self.device:
            raise lib.TtvdmError(f"sample on {sample.device}, model on {self.device}")
        B, F, Cin, H, W = sample.shape
        if F > MAX_FRAMES:
            raise lib.TtvdmError(f"num_frames = {F}: the temporal-attention kernel holds at most {MAX_FRAMES} frames per "
                                 f"sequence (the reference runs 14, SVD-XT 25)")
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
        self.begin_step()
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
            x.gn_stats = None  # modified in place: the producer's GroupNorm sums no longer describe it
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
        self.begin_step()
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
