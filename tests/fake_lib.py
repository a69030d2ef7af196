"""*** TEST INFRASTRUCTURE ONLY *** — a torch-CPU emulation of the C-ABI entry points of include/ttvdm.h.

Purpose: the `-m "not gpu"` suite has to cover the HOST logic of the engines (weight packing, buffer reuse, leading
dimensions, the kernel schedule, epilogue-fusion arguments) in a container without a GPU. The engines talk to the
kernels only through the wrappers of this_and_that_vdm_b200/lib.py, so the tests swap those wrappers for the
functions below, each of which does — in plain torch, on the same raw-pointer view of the buffers (data pointer +
leading dimension, never the tensor's own shape) — exactly what the header documents for the kernel. Results are
rounded to bf16 wherever the kernel stores bf16.

This is NOT a fallback: it lives under tests/, nothing in the product imports it (tests/test_boundary.py greps for
that), and the product still raises without libttvdm_sm100.so / an sm_100 device. The emulation itself is validated
by running the GPU-proven UNet / GestureNet engine through it against the oracle (tests/test_host_logic.py).
"""
from __future__ import annotations

import contextlib
import math

import torch
import torch.nn.functional as Fn

from this_and_that_vdm_b200 import lib

BF16 = torch.bfloat16
_launches = 0


def _mat(t: torch.Tensor, rows: int, cols: int, ld: int) -> torch.Tensor:
    """The [rows, cols] matrix with row stride ld that starts at t's data pointer (what the kernel sees)."""
    return torch.as_strided(t, (rows, cols), (ld, 1), t.storage_offset())


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def gemm(a, w, out, *, M, N, k1, mode=lib.A_LINEAR, lda=None, a2=None, k2=0, lda2=0, n_img=0, H=0, W=0, bias=None,
         rowvec=None, rows_per_vec=0, ldrv=0, s0=1.0, res1=None, ldr1=0, s1=1.0, res2=None, ldr2=0, s2=1.0,
         geglu=False, ldo=None, out_fp32=False, act=0, gn_stats_out=None, gn_rows_per_inst=0, row_sums_out=None, rs_addvec=None, rs_add_rows=0, rs_add_mod=0,
         ln_rowsums=None, ln_colsum=None, ln_eps=1e-5, prevec=None, prevec_rows=0, prevec_mod=0, ldpv=0,
         ln_row_add=None, conv_stride=1, conv_taps=0, conv_dy0=0, conv_dx0=0) -> None:
    assert k1 % 64 == 0 and k2 % 64 == 0, "gemm: k1 / k2 must be multiples of 64"
    assert a.dtype == BF16 and w.dtype == BF16
    lda = k1 if lda is None else lda
    taps = {lib.A_LINEAR: 1, lib.A_CONV3X3: 9, lib.A_TCONV3: 3}[mode]
    if mode == lib.A_CONV3X3 and conv_taps == 4:
        taps = 4
    ktot = taps * (k1 + k2)
    wm = _mat(w, N, ktot, ktot).float()
    if mode == lib.A_LINEAR:
        A = _mat(a, M, k1, lda).float()
        if a2 is not None:
            A = torch.cat([A, _mat(a2, M, k2, lda2 if lda2 else k2).float()], 1)
        acc = A @ wm.t()
    elif mode == lib.A_CONV3X3:
        assert n_img * H * W == M and a2 is None and conv_stride in (1, 2)
        cs = conv_stride  # H, W are OUTPUT dims; the input is [n_img, cs*H, cs*W, k1]
        x = _mat(a, M * cs * cs, k1, lda).float().view(n_img, H * cs, W * cs, k1).permute(0, 3, 1, 2)
        if taps == 4:
            # 2 x 2 window whose first tap sits at (conv_dy0, conv_dx0): embed it in a 3 x 3 kernel
            assert cs == 1 and conv_dy0 in (-1, 0) and conv_dx0 in (-1, 0)
            k3 = torch.zeros(N, 3, 3, k1)
            k3[:, conv_dy0 + 1:conv_dy0 + 3, conv_dx0 + 1:conv_dx0 + 3] = wm.view(N, 2, 2, k1)
        else:
            k3 = wm.view(N, 3, 3, k1)
        acc = Fn.conv2d(x, k3.permute(0, 3, 1, 2), padding=1, stride=cs).permute(0, 2, 3, 1).reshape(M, N)
    else:
        assert n_img * H * W == M and a2 is None  # n_img = B, H = F, W = S
        x = _mat(a, M, k1, lda).float().view(n_img, H, W, k1)
        xp = Fn.pad(x, (0, 0, 0, 0, 1, 1))
        acc = sum(xp[:, t:t + H] @ wm[:, t * k1:(t + 1) * k1].t() for t in range(3)).reshape(M, N)
    if ln_rowsums is not None:
        # LayerNorm of A folded into the epilogue: rstd * (acc + prevec - mean * colsum)  (include/ttvdm.h, ABI 3)
        assert ln_colsum is not None and N % 32 == 0 and not out_fp32
        parts = (k1 + k2) // 32
        rs = ln_rowsums.view(-1)[: parts * M * 2].view(parts, M, 2).float().sum(0)
        if prevec_mod > 0:
            pidx = (torch.arange(M) // prevec_rows) % prevec_mod
            if ln_row_add is not None:
                rs = rs + _mat(ln_row_add, prevec_mod, 2, 2).float()[pidx]
            if prevec is not None:
                acc = acc + _mat(prevec, prevec_mod, N, ldpv if ldpv else N).float()[pidx]
        mean = rs[:, 0] / (k1 + k2)
        var = (rs[:, 1] / (k1 + k2) - mean * mean).clamp_min(0)
        rstd = torch.rsqrt(var + ln_eps)
        acc = rstd[:, None] * (acc - mean[:, None] * ln_colsum.float()[None, :N])
    if bias is not None:
        acc = acc + bias.float()[:N]
    if rowvec is not None:
        assert rows_per_vec > 0
        n_vec = (M + rows_per_vec - 1) // rows_per_vec
        rv = _mat(rowvec, n_vec, N, ldrv if ldrv else N).float()
        acc = acc + rv.repeat_interleave(rows_per_vec, 0)[:M]
    if geglu:
        assert res1 is None and res2 is None and not out_fp32
        val = s0 * acc[:, 0::2] * Fn.gelu(acc[:, 1::2])
        n_out = N // 2
    else:
        val = s0 * acc
        n_out = N
        r1 = _mat(res1, M, N, ldr1 if ldr1 else N).float().clone() if res1 is not None else None
        r2 = _mat(res2, M, N, ldr2 if ldr2 else N).float().clone() if res2 is not None else None
        if r1 is not None:
            val = val + s1 * r1
        if r2 is not None:
            val = val + s2 * r2
        if act == 1:
            val = Fn.silu(val)
    ldo = ldo if ldo is not None else n_out
    assert out.dtype == (torch.float32 if out_fp32 else BF16)
    _mat(out, M, n_out, ldo).copy_(val.to(out.dtype))
    if gn_stats_out is not None or row_sums_out is not None:
        assert not geglu and not out_fp32 and N % 32 == 0
        stored = val.to(BF16).double()  # statistics of exactly what was stored
        if gn_stats_out is not None:
            assert gn_stats_out.dtype == torch.float64 and gn_rows_per_inst > 0 and M % gn_rows_per_inst == 0
            n_inst = M // gn_rows_per_inst
            v = stored.view(n_inst, gn_rows_per_inst, N // 2, 2)
            st = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1)  # [n_inst, N/2, 2]
            gn_stats_out.view(-1)[: n_inst * N].add_(st.reshape(-1))
        if row_sums_out is not None:
            assert row_sums_out.dtype == torch.float32
            if rs_addvec is not None:
                idx = (torch.arange(M) // rs_add_rows) % rs_add_mod
                stored = stored + _mat(rs_addvec, rs_add_mod, N, N).double()[idx]
            ch = stored.view(M, N // 32, 32)
            part = torch.stack([ch.sum(-1), (ch * ch).sum(-1)], -1).permute(1, 0, 2)  # [N/32, M, 2]
            row_sums_out.view(-1)[: (N // 32) * M * 2].copy_(part.float().reshape(-1))
    _count()


def _heads(t, rows, heads, ld, d=64):
    return _mat(t, rows, heads * d, ld).float().view(rows, heads, d)


def attn_spatial(q, k, v, out, *, ldq, ldk, ldv, ldo, n_img, heads, seq, scale) -> None:
    rows = n_img * seq
    qq, kk, vv = (_heads(t, rows, heads, ld).view(n_img, seq, heads, 64).permute(0, 2, 1, 3)
                  for t, ld in ((q, ldq), (k, ldk), (v, ldv)))
    p = torch.softmax(qq @ kk.transpose(-1, -2) * scale, -1)
    o = (p @ vv).permute(0, 2, 1, 3).reshape(rows, heads * 64)
    _mat(out, rows, heads * 64, ldo).copy_(o.to(BF16))
    _count()


def attn_cross(q, kc, vc, out, *, ldq, ldo, rows, heads, L, F, S, n_ctx, temporal, batch_offset, scale, head_dim=64) -> None:
    C = heads * head_dim
    qq = _heads(q, rows, heads, ldq, head_dim)
    kk = _mat(kc, n_ctx * L, C, C).float().view(n_ctx, L, heads, head_dim)
    vv = _mat(vc, n_ctx * L, C, C).float().view(n_ctx, L, heads, head_dim)
    r = torch.arange(rows)
    b = r // (F * S) + batch_offset
    ctx = ((b * S + r % S) % n_ctx) if temporal else b
    kr, vr = kk[ctx], vv[ctx]  # [rows, L, heads, 64]
    s = torch.einsum("rhd,rlhd->rhl", qq, kr) * scale
    o = torch.einsum("rhl,rlhd->rhd", torch.softmax(s, -1), vr).reshape(rows, C)
    _mat(out, rows, C, ldo).copy_(o.to(BF16))
    _count()


def attn_temporal(q, k, v, out, *, ldq, ldk, ldv, ldo, B, F, S, heads, scale, head_dim=64) -> None:
    rows = B * F * S
    qq, kk, vv = (_heads(t, rows, heads, ld, head_dim).view(B, F, S, heads, head_dim).permute(0, 2, 3, 1, 4)
                  for t, ld in ((q, ldq), (k, ldk), (v, ldv)))  # [B, S, heads, F, d]
    p = torch.softmax(qq @ kk.transpose(-1, -2) * scale, -1)
    o = (p @ vv).permute(0, 3, 1, 2, 4).reshape(rows, heads * head_dim)
    _mat(out, rows, heads * head_dim, ldo).copy_(o.to(BF16))
    _count()


def groupnorm(x1, out, stats, gamma, beta, *, c1, rows, rows_per_inst, eps, silu, x2=None, c2=0, ld1=None, ld2=None,
              ldo=None, pstats1=None, pstats2=None) -> None:
    assert (c1 + c2) % 32 == 0 and c1 % 8 == 0 and c2 % 8 == 0 and rows % rows_per_inst == 0
    need_pass = pstats1 is None or (x2 is not None and pstats2 is None)
    if need_pass:
        assert stats.dtype == torch.float64 and stats.numel() >= (rows // rows_per_inst) * 64
    x = _mat(x1, rows, c1, c1 if ld1 is None else ld1).float()
    if x2 is not None:
        x = torch.cat([x, _mat(x2, rows, c2, c2 if ld2 is None else ld2).float()], 1)
    C = c1 + c2
    n_inst = rows // rows_per_inst
    cpg = C // 32
    # group sums: from the tensor itself for sources without producer statistics, from the per-pair sums otherwise
    xi = x.double().view(n_inst, rows_per_inst, C)
    ch_s, ch_q = xi.sum(1), (xi * xi).sum(1)  # [n_inst, C]
    for ps, off, c in ((pstats1, 0, c1), (pstats2 if x2 is not None else None, c1, c2)):
        if ps is not None:
            assert cpg % 2 == 0 and ps.dtype == torch.float64
            pv = ps.view(-1)[: n_inst * c].view(n_inst, c // 2, 2)
            # spread each pair's sums over its two channels (only the group totals matter)
            ch_s[:, off:off + c] = (pv[:, :, 0] / 2).repeat_interleave(2, 1)
            ch_q[:, off:off + c] = (pv[:, :, 1] / 2).repeat_interleave(2, 1)
    n = rows_per_inst * cpg
    mean = ch_s.view(n_inst, 32, cpg).sum(-1) / n
    var = (ch_q.view(n_inst, 32, cpg).sum(-1) / n - mean * mean).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + eps)
    mean_c = mean.float().repeat_interleave(cpg, 1)[:, None, :]
    rstd_c = rstd.float().repeat_interleave(cpg, 1)[:, None, :]
    y = ((x.view(n_inst, rows_per_inst, C) - mean_c) * rstd_c * gamma.float() + beta.float()).reshape(rows, C)
    if silu:
        y = Fn.silu(y)
    _mat(out, rows, C, C if ldo is None else ldo).copy_(y.to(BF16))
    _count(3 if need_pass else 1)  # memset + stats + apply / apply only


def layernorm(x, out, gamma, beta, *, rows, C, eps=1e-5, addvec=None, F=0, S=0, sum_out=None, ldx=None, ldo=None,
              ldsum=None) -> None:
    xx = _mat(x, rows, C, C if ldx is None else ldx).float()
    if addvec is not None:
        idx = (torch.arange(rows) // S) % F
        xx = xx + _mat(addvec, F, C, C).float()[idx]
    if sum_out is not None:
        _mat(sum_out, rows, C, C if ldsum is None else ldsum).copy_(xx.to(BF16))
    y = Fn.layer_norm(xx, (C,), gamma.float(), beta.float(), eps)
    _mat(out, rows, C, C if ldo is None else ldo).copy_(y.to(BF16))
    _count()


def im2col_s2(x, out, *, n_img, H, W, C) -> None:
    xi = _mat(x, n_img * H * W, C, C).float().view(n_img, H, W, C).permute(0, 3, 1, 2)
    cols = Fn.unfold(xi, 3, padding=1, stride=2)  # [n, C*9, Ho*Wo], channel-major then tap
    Ho, Wo = H // 2, W // 2
    cols = cols.view(n_img, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(n_img * Ho * Wo, 9 * C)
    _mat(out, n_img * Ho * Wo, 9 * C, 9 * C).copy_(cols.to(BF16))
    _count()


def interleave2x(parts, out, *, n_img, H, W, C) -> None:
    pp = _mat(parts, 4 * n_img * H * W, C, C).view(2, 2, n_img, H, W, C)  # [py, px, n, y, x, c]
    _mat(out, n_img * 4 * H * W, C, C).copy_(pp.permute(2, 3, 0, 4, 1, 5).reshape(n_img * 4 * H * W, C))
    _count()


def upsample2x(x, out, *, n_img, H, W, C) -> None:
    xi = _mat(x, n_img * H * W, C, C).view(n_img, H, W, C)
    up = xi.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(n_img * 4 * H * W, C)
    _mat(out, n_img * 4 * H * W, C, C).copy_(up)
    _count()


def sinusoid(t, out, *, n, dim) -> None:
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t.float().reshape(-1)[:n, None] * freq[None]
    _mat(out, n, dim, dim).copy_(torch.cat([torch.cos(arg), torch.sin(arg)], 1).to(BF16))
    _count()


def axpy(a, b, out, scale, n) -> None:
    val = a.reshape(-1)[:n].float() + scale * b.reshape(-1)[:n].float()
    out.view(-1)[:n].copy_(val.to(out.dtype))
    _count()


def sampler_prepare(latents, image_latents, cond, model_in, *, c_pad, B_local, batch_offset, F, h, w, sigma) -> None:
    x = torch.zeros(B_local, F, h, w, c_pad)
    lat = (latents.float() / math.sqrt(sigma * sigma + 1.0)).permute(0, 2, 3, 1)  # [F, h, w, 4]
    for b in range(B_local):
        x[b, ..., 0:4] = lat
        x[b, ..., 4:8] = image_latents[batch_offset + b].float().permute(1, 2, 0)[None]
        if cond is not None:
            x[b, ..., 8:12] = cond.float().permute(0, 2, 3, 1)
    _mat(model_in, B_local * F * h * w, c_pad, c_pad).copy_(x.view(-1, c_pad).to(BF16))
    _count()


def sampler_euler_step(latents, eps_u, eps_c, guidance, *, ld_eps, F, h, w, sigma, sigma_next) -> None:
    n = F * h * w
    eu = _mat(eps_u, n, 4, ld_eps).float().view(F, h, w, 4).permute(0, 3, 1, 2)
    ec = _mat(eps_c, n, 4, ld_eps).float().view(F, h, w, 4).permute(0, 3, 1, 2)
    eps = eu + guidance.float().view(F, 1, 1, 1) * (ec - eu)
    x = latents.float()
    x0 = eps * (-sigma / math.sqrt(sigma * sigma + 1.0)) + x / (sigma * sigma + 1.0)
    latents.copy_(x + (x - x0) / sigma * (sigma_next - sigma))
    _count()


# ---- VAE entry points (include/ttvdm.h "VAE" section)
def softmax_rows(x, out, *, rows, cols, ldx, ldo, cols_out, causal=False) -> None:
    sc = _mat(x, rows, cols, ldx).float()
    period = rows if causal is True else int(causal)
    if period:
        sc = sc.masked_fill(torch.arange(cols)[None, :] > (torch.arange(rows) % period)[:, None], float("-inf"))
    p = torch.softmax(sc, -1)
    o = torch.zeros(rows, cols_out)
    o[:, :cols] = p
    _mat(out, rows, cols_out, ldo).copy_(o.to(BF16))
    _count()


def im2col_s2_pad01(x, out, *, n_img, H, W, C) -> None:
    """Downsample2D of the VAE encoder: F.pad(x, (0, 1, 0, 1)) then Conv2d(3, stride 2, padding 0)."""
    xi = _mat(x, n_img * H * W, C, C).float().view(n_img, H, W, C).permute(0, 3, 1, 2)
    cols = Fn.unfold(Fn.pad(xi, (0, 1, 0, 1)), 3, padding=0, stride=2)
    Ho, Wo = H // 2, W // 2
    cols = cols.view(n_img, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(n_img * Ho * Wo, 9 * C)
    _mat(out, n_img * Ho * Wo, 9 * C, 9 * C).copy_(cols.to(BF16))
    _count()


def vae_time_conv_out(x, w, bias, out, *, B, F, H, W, ldx) -> None:
    """x fp32 [(b, f, s), ldx] (3 real channels) -> out fp32 NCHW [B*F, 3, H, W]; Conv3d(3, 3, (3,1,1), pad (1,0,0))."""
    S = H * W
    xi = _mat(x, B * F * S, 3, ldx).float().view(B, F, S, 3).permute(0, 3, 1, 2)[..., None]  # [B, 3, F, S, 1]
    y = Fn.conv3d(xi, w.float().view(3, 3, 3, 1, 1), bias.float(), padding=(1, 0, 0))  # [B, 3, F, S, 1]
    out.view(B, F, 3, S).copy_(y[..., 0].permute(0, 2, 1, 3))
    _count()


def act_inplace(x, kind) -> None:
    v = x.float()
    x.copy_((Fn.gelu(v) if kind == 2 else v * torch.sigmoid(1.702 * v)).to(BF16))
    _count()


def layernorm_flat(x, out, *, rows, n, eps=1e-5) -> None:
    out.view(rows, n).copy_(Fn.layer_norm(x.view(rows, n).float(), (n,), None, None, eps))
    _count()


# ---- weight repack entry points
def pack_conv_weight(w, out, *, cin_pad=0) -> None:
    cout, cin = w.shape[:2]
    taps = w[0, 0].numel()
    cp = max(cin_pad, cin)
    v = w.detach().float().reshape(cout, cin, taps).permute(0, 2, 1)  # [cout, taps, cin]
    if cp > cin:
        v = Fn.pad(v, (0, cp - cin))
    out.view(-1)[: cout * taps * cp].copy_(v.reshape(-1).to(BF16))
    _count()


def pack_linear(w, out_w, *, bias=None, gamma=None, beta=None, out_bias=None, out_colsum=None, geglu=False,
                out_row0=0) -> None:
    N = w.shape[0]
    w32 = w.detach().float().reshape(N, -1)
    wf = (w32 * gamma.detach().float()[None, :] if gamma is not None else w32).to(BF16)
    rows = torch.arange(N)
    if geglu:
        half = N // 2
        rows = torch.where(rows < half, 2 * rows, 2 * (rows - half) + 1)
    rows = rows + out_row0
    out_w[rows] = wf
    if out_colsum is not None:
        out_colsum[rows] = wf.float().sum(1)
    if out_bias is not None:
        b = w32 @ beta.detach().float() if beta is not None else torch.zeros(N)
        if bias is not None:
            b = b + bias.detach().float()
        out_bias[rows] = b
    _count()


def pack_vector(src, out) -> None:
    out.view(-1)[: src.numel()].copy_(src.detach().float().reshape(-1))
    _count()


def launch_count() -> int:
    return _launches


_PATCHED = ["gemm", "attn_spatial", "attn_cross", "attn_temporal", "groupnorm", "layernorm", "im2col_s2", "upsample2x",
            "sinusoid", "axpy", "sampler_prepare", "sampler_euler_step", "softmax_rows", "im2col_s2_pad01",
            "vae_time_conv_out", "interleave2x", "act_inplace", "layernorm_flat", "launch_count", "pack_conv_weight", "pack_linear",
            "pack_vector"]


@contextlib.contextmanager
def installed():
    """Swap the ctypes wrappers of this_and_that_vdm_b200.lib for the emulation (and make init() a no-op)."""
    saved = {n: getattr(lib, n, None) for n in _PATCHED + ["init"]}
    g = globals()
    try:
        for n in _PATCHED:
            setattr(lib, n, g[n])
        lib.init = lambda device=None: None
        yield
    finally:
        for n, f in saved.items():
            if f is None:
                if hasattr(lib, n):
                    delattr(lib, n)
            else:
                setattr(lib, n, f)
