"""B200 execution engine of the VAE either side of the denoising loop (scope table row "next #1").

Host-side (Python) weight packing + kernel schedule for diffusers' AutoencoderKLTemporalDecoder as the reference
calls it:
  * encode : `vae.encode(x).latent_dist`   svd/pipeline_stable_video_diffusion_controlnet.py:199 (first frame), :652
             (the 14 gesture condition frames — the reference re-encodes them on every one of the 25 steps; the
             pipelines of this repo hoist that out of the loop)
  * decode : `vae.decode(z, num_frames)`   :257-283 (decode_latents, one call per chunk of frames)

Data layout in HBM: as in engine.py — bf16 channels-last token matrices [n_img*H*W, C], rows ordered (image, pixel);
the temporal layers of the decoder walk frames with a row stride of H*W. Every 3x3 convolution, 1x1 shortcut,
linear and 3-tap temporal convolution is one ttvdm_gemm call (tcgen05, fused bias / residual / blend epilogue), every
GroupNorm(+SiLU) one ttvdm_groupnorm call. The mid-block attention has ONE head of C = 512 dims over all H*W/64
tokens of a frame; it runs per frame as scores = gemm(Q, K) (fp32) -> ttvdm_softmax_rows -> gemm(P, V^T), with V^T
produced directly by a GEMM with swapped operands (no transpose pass) and the value bias added after P V (rows of P
sum to one). Weight folds done at pack time: quant_conv (1x1) into the encoder's conv_out; AlphaBlender into the
epilogue of the second temporal convolution (out = spatial + sigmoid(mix_factor) * conv).

torch is used only to own device memory and for the NCHW fp32/fp16 <-> channels-last bf16 conversion at the
module boundary. No CPU / eager fallback exists.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import os

import torch

from . import lib
from .engine import BF16, PAD_IN, _bf, _f32, _pack_conv3x3, _pack_tconv, _pack_upsample_parity

# kernels index single tensors with 64-bit offsets; per-image ops are still issued in groups of at most this many
# elements so that no launch sees a tensor beyond 2^31 elements (TMA box coordinates stay far from their limits)
_MAX_ELEMS = 1 << 30


@dataclass
class VResW:
    cin: int
    cout: int
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
    # temporal half (decoder only)
    temporal: bool = False
    tn1_g: torch.Tensor = None
    tn1_b: torch.Tensor = None
    tw1: torch.Tensor = None
    tb1: torch.Tensor = None
    tn2_g: torch.Tensor = None
    tn2_b: torch.Tensor = None
    tw2: torch.Tensor = None
    tb2: torch.Tensor = None
    t_scale: float = 0.5  # weight of the temporal conv branch = sigmoid(mix_factor)


@dataclass
class VAttnW:
    C: int
    gn_g: torch.Tensor = None
    gn_b: torch.Tensor = None
    wq: torch.Tensor = None
    bq: torch.Tensor = None
    wk: torch.Tensor = None
    bk: torch.Tensor = None
    wv: torch.Tensor = None
    bv: torch.Tensor = None
    wo: torch.Tensor = None
    bo: torch.Tensor = None


class VaeEngine:
    """Packed weights + kernel schedule of one AutoencoderKLTemporalDecoder."""

    def __init__(self, model):
        lib.init()
        self.device = model.device
        self.cfg = model.config
        chans = tuple(self.cfg.block_out_channels)
        if any(c % 64 != 0 for c in chans):
            raise lib.TtvdmError(f"sm_100a VAE engine needs block_out_channels %% 64 == 0 (got {chans})")
        if self.cfg.in_channels != 3 or self.cfg.out_channels != 3:
            raise lib.TtvdmError("sm_100a VAE engine supports RGB in / out (in_channels = out_channels = 3)")
        self.latent_channels = int(self.cfg.latent_channels)
        # GroupNorm statistics from the producing GEMM's epilogue (round 2; TTVDM_VAE_FUSE_GN=0: standalone statistics pass)
        self.fuse_norm_stats = os.environ.get("TTVDM_VAE_FUSE_GN", "1") != "0"
        self._pack(model)

    # ============================================================================================ packing
    def _pack(self, m) -> None:
        dev = self.device
        sd = dict(m.state_dict())
        g = lambda k: sd[k]  # noqa: E731

        def res2d(p: str) -> VResW:
            cout, cin = g(p + ".conv1.weight").shape[:2]
            r = VResW(cin=cin, cout=cout)
            r.n1_g, r.n1_b = _f32(g(p + ".norm1.weight"), dev), _f32(g(p + ".norm1.bias"), dev)
            r.w1, r.b1 = _pack_conv3x3(g(p + ".conv1.weight"), dev), _f32(g(p + ".conv1.bias"), dev)
            r.n2_g, r.n2_b = _f32(g(p + ".norm2.weight"), dev), _f32(g(p + ".norm2.bias"), dev)
            r.w2, r.b2 = _pack_conv3x3(g(p + ".conv2.weight"), dev), _f32(g(p + ".conv2.bias"), dev)
            if (p + ".conv_shortcut.weight") in sd:
                r.wsc = _bf(g(p + ".conv_shortcut.weight").reshape(cout, cin), dev)
                r.bsc = _f32(g(p + ".conv_shortcut.bias"), dev)
            return r

        def res3d(p: str) -> VResW:
            r = res2d(p + ".spatial_res_block")
            tp = p + ".temporal_res_block"
            r.temporal = True
            r.tn1_g, r.tn1_b = _f32(g(tp + ".norm1.weight"), dev), _f32(g(tp + ".norm1.bias"), dev)
            r.tw1, r.tb1 = _pack_tconv(g(tp + ".conv1.weight"), dev), _f32(g(tp + ".conv1.bias"), dev)
            r.tn2_g, r.tn2_b = _f32(g(tp + ".norm2.weight"), dev), _f32(g(tp + ".norm2.bias"), dev)
            r.tw2, r.tb2 = _pack_tconv(g(tp + ".conv2.weight"), dev), _f32(g(tp + ".conv2.bias"), dev)
            # AlphaBlender("learned", switch_spatial_to_temporal_mix=True): a = 1 - sigmoid(mix);
            # a*s + (1-a)*(s + conv) = s + sigmoid(mix) * conv
            r.t_scale = float(torch.sigmoid(g(p + ".time_mixer.mix_factor").float()).item())
            return r

        def attn(p: str) -> VAttnW:
            C = g(p + ".to_q.weight").shape[0]
            a = VAttnW(C=C)
            a.gn_g, a.gn_b = _f32(g(p + ".group_norm.weight"), dev), _f32(g(p + ".group_norm.bias"), dev)
            a.wq, a.bq = _bf(g(p + ".to_q.weight"), dev), _f32(g(p + ".to_q.bias"), dev)
            a.wk, a.bk = _bf(g(p + ".to_k.weight"), dev), _f32(g(p + ".to_k.bias"), dev)
            a.wv, a.bv = _bf(g(p + ".to_v.weight"), dev), _f32(g(p + ".to_v.bias"), dev)
            a.wo, a.bo = _bf(g(p + ".to_out.0.weight"), dev), _f32(g(p + ".to_out.0.bias"), dev)
            return a

        def count(fmt: str) -> int:
            n = 0
            while fmt.format(n) in sd:
                n += 1
            return n

        # ---- encoder
        self.e_conv_in_w = _pack_conv3x3(g("encoder.conv_in.weight"), dev, pad_cin=PAD_IN)
        self.e_conv_in_b = _f32(g("encoder.conv_in.bias"), dev)
        self.e_down: List[dict] = []
        for i in range(count("encoder.down_blocks.{}.resnets.0.norm1.weight")):
            p = f"encoder.down_blocks.{i}"
            blk = {"res": [res2d(f"{p}.resnets.{j}") for j in range(count(p + ".resnets.{}.norm1.weight"))],
                   "down_w": None, "down_b": None}
            if (p + ".downsamplers.0.conv.weight") in sd:
                blk["down_w"] = _pack_conv3x3(g(p + ".downsamplers.0.conv.weight"), dev)
                blk["down_b"] = _f32(g(p + ".downsamplers.0.conv.bias"), dev)
            self.e_down.append(blk)
        self.e_mid_res = [res2d("encoder.mid_block.resnets.0"), res2d("encoder.mid_block.resnets.1")]
        self.e_mid_attn = attn("encoder.mid_block.attentions.0")
        self.e_out_g, self.e_out_b = _f32(g("encoder.conv_norm_out.weight"), dev), _f32(g("encoder.conv_norm_out.bias"), dev)
        # quant_conv (1x1) o conv_out (3x3) is one 3x3 convolution: W' = Wq Wc, b' = Wq bc + bq (folded in fp32)
        wc, bc = g("encoder.conv_out.weight").float(), g("encoder.conv_out.bias").float()
        wq = g("quant_conv.weight").float().reshape(g("quant_conv.weight").shape[0], -1)
        bq = g("quant_conv.bias").float()
        self.e_conv_out_w = _pack_conv3x3(torch.einsum("om,mikl->oikl", wq, wc), dev)
        self.e_conv_out_b = _f32(wq @ bc + bq, dev)
        self.moment_channels = wq.shape[0]
        # ---- decoder
        self.d_conv_in_w = _pack_conv3x3(g("decoder.conv_in.weight"), dev, pad_cin=PAD_IN)
        self.d_conv_in_b = _f32(g("decoder.conv_in.bias"), dev)
        n_mid = count("decoder.mid_block.resnets.{}.spatial_res_block.norm1.weight")
        self.d_mid_res = [res3d(f"decoder.mid_block.resnets.{j}") for j in range(n_mid)]
        self.d_mid_attn = [attn(f"decoder.mid_block.attentions.{j}")
                           for j in range(count("decoder.mid_block.attentions.{}.to_q.weight"))]
        self.d_up: List[dict] = []
        for i in range(count("decoder.up_blocks.{}.resnets.0.spatial_res_block.norm1.weight")):
            p = f"decoder.up_blocks.{i}"
            n_res = count(p + ".resnets.{}.spatial_res_block.norm1.weight")
            blk = {"res": [res3d(f"{p}.resnets.{j}") for j in range(n_res)], "up_w": None, "up_b": None}
            if (p + ".upsamplers.0.conv.weight") in sd:
                blk["up_w"] = _pack_conv3x3(g(p + ".upsamplers.0.conv.weight"), dev)
                blk["up_wp"] = _pack_upsample_parity(g(p + ".upsamplers.0.conv.weight"), dev)
                blk["up_b"] = _f32(g(p + ".upsamplers.0.conv.bias"), dev)
            self.d_up.append(blk)
        self.d_out_g, self.d_out_b = _f32(g("decoder.conv_norm_out.weight"), dev), _f32(g("decoder.conv_norm_out.bias"), dev)
        # conv_out 128 -> 3, padded to 4 output features (fp32 [rows, 4], the shape class of the UNet's conv_out)
        w = g("decoder.conv_out.weight")
        self.d_conv_out_w = _pack_conv3x3(torch.cat([w, torch.zeros_like(w[:1])], 0), dev)
        self.d_conv_out_b = _f32(torch.cat([g("decoder.conv_out.bias"), torch.zeros(1, device=w.device, dtype=w.dtype)]), dev)
        # time_conv_out weights travel in kernel-parameter space: host fp32 [co, ci, t]
        self.d_time_w = g("decoder.time_conv_out.weight").detach().float().reshape(3, 3, 3).cpu().contiguous()
        self.d_time_b = g("decoder.time_conv_out.bias").detach().float().cpu().contiguous()

    # ============================================================================================ helpers
    def _empty(self, *shape, dtype=BF16) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    @staticmethod
    def _groups(n_img: int, rows_per_img: int, width: int):
        """Image ranges [i0, i1) whose tensors stay below _MAX_ELEMS elements (per-image ops only)."""
        per = max(1, _MAX_ELEMS // max(1, rows_per_img * width))
        return [(i, min(n_img, i + per)) for i in range(0, n_img, per)]

    def _gn(self, x, gamma, beta, *, n_img, S, eps, silu, frames_per_inst=1):
        """GroupNorm(32) (+SiLU). frames_per_inst = 1: per-image statistics; = F: the 5-D norm of TemporalResnetBlock.
        A tensor whose producing GEMM left per-(instance, channel pair) sums for this instance size (x.gn_stats, round 2)
        needs no statistics pass: the call is the apply pass only."""
        C = x.shape[1]
        rows = n_img * S
        out = self._empty(rows, C)
        rpi = frames_per_inst * S
        st = getattr(x, "gn_stats", None)
        ps = st[0] if (st is not None and st[1] == rpi and (C // 32) % 2 == 0) else None
        if frames_per_inst > 1 or ps is not None:
            stats = None if ps is not None else self._empty((n_img // frames_per_inst) * 64, dtype=torch.float64)
            lib.groupnorm(x, out, stats, gamma, beta, c1=C, rows=rows, rows_per_inst=rpi, eps=eps, silu=silu, pstats1=ps)
            return out
        for i0, i1 in self._groups(n_img, S, C):
            stats = self._empty((i1 - i0) * 64, dtype=torch.float64)
            lib.groupnorm(x[i0 * S:i1 * S], out[i0 * S:i1 * S], stats, gamma, beta, c1=C, rows=(i1 - i0) * S,
                          rows_per_inst=S, eps=eps, silu=silu)
        return out

    def _gn_sums(self, out, M: int, N: int, gn_rpi: int):
        """Zeroed fp64 [M / gn_rpi, N / 2, 2] buffer for the epilogue of the GEMM that writes `out` (ttvdm_gemm
        gn_stats_out) when its consumer is a GroupNorm over gn_rpi rows; rides on the tensor. None when not applicable."""
        out.gn_stats = None
        if not gn_rpi or not self.fuse_norm_stats or N % 64 != 0 or M % gn_rpi != 0 or out.dtype != BF16:
            return None
        st = torch.zeros((M // gn_rpi) * N, dtype=torch.float64, device=self.device)
        out.gn_stats = (st, gn_rpi)
        return st

    def _conv3(self, x, w, bias, *, n_img, H, W, cin, res1=None, out=None, out_fp32=False, gn_rpi=0):
        """gn_rpi: rows per GroupNorm instance of the CONSUMER of `out` (S: per image, F * S: the 5-D temporal norm): the
        epilogue then leaves the statistics (whole instances per launch only)."""
        N = w.shape[0]
        S = H * W
        if out is None:
            out = self._empty(n_img * S, N, dtype=torch.float32 if out_fp32 else BF16)
        st = self._gn_sums(out, n_img * S, N, gn_rpi)
        for i0, i1 in self._groups(n_img, S, max(cin, N)):
            sl = slice(i0 * S, i1 * S)
            kw = {}
            if st is not None:
                if ((i0 * S) % gn_rpi) or ((i1 - i0) * S) % gn_rpi:
                    out.gn_stats = None  # an instance would straddle two launches: leave it to the statistics pass
                    st = None
                else:
                    kw = dict(gn_stats_out=st[(i0 * S // gn_rpi) * N:], gn_rows_per_inst=gn_rpi)
            lib.gemm(x[sl], w, out[sl], M=(i1 - i0) * S, N=N, k1=cin, mode=lib.A_CONV3X3, n_img=i1 - i0, H=H, W=W,
                     bias=bias, res1=None if res1 is None else res1[sl], out_fp32=out_fp32, **kw)
        return out

    def _linear(self, a, w, *, M, bias=None, res1=None, out=None, out_fp32=False, gn_rpi=0):
        N, K = w.shape
        if out is None:
            out = self._empty(M, N, dtype=torch.float32 if out_fp32 else BF16)
        st = self._gn_sums(out, M, N, gn_rpi)
        rows_per = max(1, _MAX_ELEMS // max(N, K))
        if st is not None:
            rows_per = max(gn_rpi, rows_per // gn_rpi * gn_rpi)  # whole instances per launch
        for r0 in range(0, M, rows_per):
            r1 = min(M, r0 + rows_per)
            kw = dict(gn_stats_out=st[(r0 // gn_rpi) * N:], gn_rows_per_inst=gn_rpi) if st is not None else {}
            lib.gemm(a[r0:r1], w, out[r0:r1], M=r1 - r0, N=N, k1=K, bias=bias,
                     res1=None if res1 is None else res1[r0:r1], out_fp32=out_fp32, **kw)
        return out

    def _upsample_conv(self, x, blk, *, n_img, H, W):
        """Upsample2D (nearest x2 + 3x3 conv) as four 2x2-tap parity convolutions of the LOW-resolution input with pre-summed
        weights + ttvdm_interleave2x (engine._upsample_conv / _pack_upsample_parity): 16 instead of 36 tap-pixels and no
        upsampled tensor — the three upsample convolutions are 22 % of the decoder's FLOPs. Per-image GroupNorm sums of the
        next ResBlock come out of the four epilogues. TTVDM_UPSAMPLE_PARITY=0: round 1's upsample2x + conv."""
        C = x.shape[1]
        S = H * W
        if os.environ.get("TTVDM_UPSAMPLE_PARITY", "1") == "0":
            up = self._empty(n_img * 4 * S, C)
            for i0, i1 in self._groups(n_img, 4 * S, C):
                lib.upsample2x(x[i0 * S:i1 * S], up[i0 * 4 * S:i1 * 4 * S], n_img=i1 - i0, H=H, W=W, C=C)
            return self._conv3(up, blk["up_w"], blk["up_b"], n_img=n_img, H=2 * H, W=2 * W, cin=C, gn_rpi=4 * S)
        N = blk["up_wp"][0].shape[0]
        out = self._empty(n_img * 4 * S, N)
        st = self._gn_sums(out, n_img * 4 * S, N, 4 * S)  # [n_img, N / 2, 2]: an instance is an image either way
        for i0, i1 in self._groups(n_img, 4 * S, max(C, N)):
            ni = i1 - i0
            parts = self._empty(4, ni * S, N)
            kw = dict(gn_stats_out=st[i0 * N:], gn_rows_per_inst=S) if st is not None else {}
            for pi, (py, px) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                lib.gemm(x[i0 * S:i1 * S], blk["up_wp"][pi], parts[pi], M=ni * S, N=N, k1=C, mode=lib.A_CONV3X3, n_img=ni,
                         H=H, W=W, bias=blk["up_b"], conv_taps=4, conv_dy0=py - 1, conv_dx0=px - 1, **kw)
            lib.interleave2x(parts, out[i0 * 4 * S:i1 * 4 * S], n_img=ni, H=H, W=W, C=N)
            del parts
        return out

    # ============================================================================================ blocks
    def _resblock(self, r: VResW, x, *, n_img, F, H, W):
        """ResnetBlock2D(temb=None) [+ TemporalResnetBlock + AlphaBlender when r.temporal]. x [n_img*H*W, cin]."""
        S = H * W
        rows = n_img * S
        y = self._gn(x, r.n1_g, r.n1_b, n_img=n_img, S=S, eps=1e-6, silu=True)
        h = self._conv3(y, r.w1, r.b1, n_img=n_img, H=H, W=W, cin=r.cin, gn_rpi=S)  # -> norm2 (per image)
        y = self._gn(h, r.n2_g, r.n2_b, n_img=n_img, S=S, eps=1e-6, silu=True)
        sc = x if r.wsc is None else self._linear(x, r.wsc, M=rows, bias=r.bsc)
        # conv2 + shortcut feeds the temporal block's 5-D norm (F * S rows per instance) or the next block's per-image norm
        hs = self._conv3(y, r.w2, r.b2, n_img=n_img, H=H, W=W, cin=r.cout, res1=sc, out=h,
                         gn_rpi=(F * S if r.temporal else S))
        if not r.temporal:
            return hs
        B = n_img // F
        y = self._gn(hs, r.tn1_g, r.tn1_b, n_img=n_img, S=S, eps=1e-5, silu=True, frames_per_inst=F)
        t1 = self._empty(rows, r.cout)
        st = self._gn_sums(t1, rows, r.cout, F * S)  # -> tn2 (5-D)
        kw = dict(gn_stats_out=st, gn_rows_per_inst=F * S) if st is not None else {}
        lib.gemm(y, r.tw1, t1, M=rows, N=r.cout, k1=r.cout, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=r.tb1, **kw)
        y = self._gn(t1, r.tn2_g, r.tn2_b, n_img=n_img, S=S, eps=1e-5, silu=True, frames_per_inst=F)
        st = self._gn_sums(t1, rows, r.cout, S)  # the blended output -> the next block's per-image norm
        kw = dict(gn_stats_out=st, gn_rows_per_inst=S) if st is not None else {}
        lib.gemm(y, r.tw2, t1, M=rows, N=r.cout, k1=r.cout, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=r.tb2,
                 s0=r.t_scale, res1=hs, s1=1.0, **kw)
        return t1

    def _attention(self, a: VAttnW, x, *, n_img, H, W):
        """diffusers Attention on a 4-D input (heads = 1, dim_head = C): GroupNorm -> q,k,v -> softmax(q k^T / sqrt C) v
        -> to_out + residual. One (scores, softmax, P V) triple per image; the fp32 score buffer is reused."""
        S = H * W
        C = a.C
        rows = n_img * S
        Sp = (S + 63) // 64 * 64  # K of the P V GEMM must be a multiple of 64; the padding columns are zeros
        y = self._gn(x, a.gn_g, a.gn_b, n_img=n_img, S=S, eps=1e-6, silu=False)
        q = self._linear(y, a.wq, M=rows, bias=a.bq)
        k = self._linear(y, a.wk, M=rows, bias=a.bk)
        o = self._empty(rows, C)
        scores = self._empty(S, S, dtype=torch.float32)
        probs = self._empty(S, Sp)
        vt = torch.zeros(C, Sp, dtype=BF16, device=self.device)
        scale = float(C) ** -0.5
        for i in range(n_img):
            sl = slice(i * S, (i + 1) * S)
            # V^T[c, s] = sum_k Wv[c, k] y[s, k]: the weight is the A operand, the tokens are the "weight" rows
            lib.gemm(a.wv, y[sl], vt, M=C, N=S, k1=C, ldo=Sp)
            lib.gemm(q[sl], k[sl], scores, M=S, N=S, k1=C, s0=scale, out_fp32=True)
            lib.softmax_rows(scores, probs, rows=S, cols=S, ldx=S, ldo=Sp, cols_out=Sp)
            lib.gemm(probs, vt, o[sl], M=S, N=C, k1=Sp, bias=a.bv)  # + b_v: rows of P sum to one
        return self._linear(o, a.wo, M=rows, bias=a.bo, res1=x, out=q, gn_rpi=S)  # -> the next ResBlock's norm1

    # ============================================================================================ boundary API
    def _to_tokens(self, t: torch.Tensor, pad_to: int) -> torch.Tensor:
        """NCHW (any float dtype) -> channels-last bf16 [N*H*W, pad_to] with zero-padded channels."""
        if t.device != self.device:
            raise lib.TtvdmError(f"input on {t.device}, VAE on {self.device}")
        N, C, H, W = t.shape
        x = torch.zeros(N, H, W, pad_to, dtype=BF16, device=self.device)
        x[..., :C] = t.permute(0, 2, 3, 1)
        return x.view(N * H * W, pad_to)

    def encode(self, x: torch.Tensor, max_images_per_pass: int = 4, dedupe: bool = True) -> torch.Tensor:
        """x [N, 3, H, W] in [-1, 1] -> moments [N, 2*latent, H/f, W/f] in x.dtype (mean | logvar, after quant_conv).
        dedupe: bit-identical images are encoded once (the encoder has no cross-image op) — 12 of the 14 gesture
        condition frames of the reference are all-zero images (data_loader/video_this_that_dataset.py:28-130)."""
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f"expected an image batch [N, 3, H, W], got {tuple(x.shape)}")
        n_down = sum(1 for b in self.e_down if b["down_w"] is not None)
        N, _, H, W = x.shape
        if H % (1 << n_down) or W % (1 << n_down):
            raise ValueError(f"image height / width must be multiples of {1 << n_down}, got {H}x{W}")
        inverse = None
        if dedupe and N > 1:
            flat = x.reshape(N, -1)
            first = [0]  # index of the first occurrence of every distinct image, in order of appearance
            inv = [0] * N
            for i in range(1, N):
                for u, j in enumerate(first):
                    if torch.equal(flat[i], flat[j]):
                        inv[i] = u
                        break
                else:
                    inv[i] = len(first)
                    first.append(i)
            if len(first) < N:
                inverse = torch.tensor(inv, device=x.device)
                x = x[torch.tensor(first, device=x.device)]
        outs = []
        for i0 in range(0, x.shape[0], max_images_per_pass):
            outs.append(self._encode_pass(x[i0:i0 + max_images_per_pass]))
        mom = torch.cat(outs, 0).to(x.dtype)
        return mom if inverse is None else mom[inverse]

    def _encode_pass(self, x: torch.Tensor) -> torch.Tensor:
        n, _, H, W = x.shape
        h = self._conv3(self._to_tokens(x, PAD_IN), self.e_conv_in_w, self.e_conv_in_b, n_img=n, H=H, W=W, cin=PAD_IN,
                        gn_rpi=H * W)
        for blk in self.e_down:
            for r in blk["res"]:
                h = self._resblock(r, h, n_img=n, F=1, H=H, W=W)
            if blk["down_w"] is not None:
                C = h.shape[1]
                col = self._empty(n * (H // 2) * (W // 2), 9 * C)
                lib.im2col_s2_pad01(h, col, n_img=n, H=H, W=W, C=C)
                H, W = H // 2, W // 2
                h = self._linear(col, blk["down_w"], M=n * H * W, bias=blk["down_b"], gn_rpi=H * W)
        h = self._resblock(self.e_mid_res[0], h, n_img=n, F=1, H=H, W=W)
        h = self._attention(self.e_mid_attn, h, n_img=n, H=H, W=W)
        h = self._resblock(self.e_mid_res[1], h, n_img=n, F=1, H=H, W=W)
        y = self._gn(h, self.e_out_g, self.e_out_b, n_img=n, S=H * W, eps=1e-6, silu=True)
        mom = self._conv3(y, self.e_conv_out_w, self.e_conv_out_b, n_img=n, H=H, W=W, cin=h.shape[1], out_fp32=True)
        return mom.view(n, H, W, self.moment_channels).permute(0, 3, 1, 2).contiguous()

    def decode(self, z: torch.Tensor, num_frames: int) -> torch.Tensor:
        """z [B*num_frames, latent, h, w] (already divided by the scaling factor) -> [B*num_frames, 3, 8h, 8w] in
        z.dtype. Frames of one video are consecutive (diffusers' `sample[None, :].reshape(B, F, ...)`)."""
        if z.dim() != 4 or z.shape[1] != self.latent_channels:
            raise ValueError(f"expected latents [N, {self.latent_channels}, h, w], got {tuple(z.shape)}")
        n, _, H, W = z.shape
        if n % num_frames != 0:
            raise ValueError(f"{n} latent frames are not a multiple of num_frames={num_frames}")
        F = num_frames
        kw = dict(n_img=n, F=F)
        h = self._conv3(self._to_tokens(z, PAD_IN), self.d_conv_in_w, self.d_conv_in_b, n_img=n, H=H, W=W, cin=PAD_IN,
                        gn_rpi=H * W)
        h = self._resblock(self.d_mid_res[0], h, H=H, W=W, **kw)
        for att, r in zip(self.d_mid_attn, self.d_mid_res[1:]):
            h = self._attention(att, h, n_img=n, H=H, W=W)
            h = self._resblock(r, h, H=H, W=W, **kw)
        for blk in self.d_up:
            for r in blk["res"]:
                h = self._resblock(r, h, H=H, W=W, **kw)
            if blk["up_w"] is not None:
                h = self._upsample_conv(h, blk, n_img=n, H=H, W=W)
                H, W = 2 * H, 2 * W
        y = self._gn(h, self.d_out_g, self.d_out_b, n_img=n, S=H * W, eps=1e-6, silu=True)
        rgb = self._conv3(y, self.d_conv_out_w, self.d_conv_out_b, n_img=n, H=H, W=W, cin=h.shape[1], out_fp32=True)
        out = self._empty(n, 3, H, W, dtype=torch.float32)
        lib.vae_time_conv_out(rgb, self.d_time_w, self.d_time_b, out, B=n // F, F=F, H=H, W=W, ldx=rgb.shape[1])
        return out.to(z.dtype)
