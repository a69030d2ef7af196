"""Kernel-level numerics + timing sweep on a B200 (run under gpurun). Each case compares one C-ABI entry point
with the same op in plain torch fp32 on the GPU, fed the identical bf16 inputs. Never raises: prints one line
per case so a single GPU call reports on every kernel. `pytest -m gpu` runs the same cases with asserts.
"""
from __future__ import annotations

import json
import math
import sys
import time
import traceback
from pathlib import Path

import torch
import torch.nn.functional as Fn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from this_and_that_vdm_b200 import lib  # noqa: E402

DEV = "cuda"


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.float()
    b = b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def bf(x):
    return x.to(torch.bfloat16)


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=gen) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ cases
def case_gemm_linear(M=1000, N=320, K=320, bias=True, res=False, geglu=False, a2=False, seed=0):
    a = bf(g(M, K, seed=seed))
    w = bf(g(N, K + (64 if a2 else 0), seed=seed + 1, scale=K ** -0.5))
    b = g(N, seed=seed + 2) if bias else None
    a_2 = bf(g(M, 64, seed=seed + 3)) if a2 else None
    r1 = bf(g(M, N, seed=seed + 4)) if res else None
    r2 = bf(g(M, N, seed=seed + 5)) if res else None
    out = torch.empty(M, N // 2 if geglu else N, dtype=torch.bfloat16, device=DEV)
    lib.gemm(a, w, out, M=M, N=N, k1=K, a2=a_2, k2=64 if a2 else 0, bias=b, res1=r1, s1=0.5, res2=r2, s2=0.25,
             s0=0.75 if res else 1.0, geglu=geglu)
    af = torch.cat([a, a_2], 1).float() if a2 else a.float()
    ref = af @ w.float().t()
    if bias:
        ref = ref + b
    if geglu:
        ref = ref[:, 0::2] * Fn.gelu(ref[:, 1::2])
    if res:
        ref = 0.75 * ref + 0.5 * r1.float() + 0.25 * r2.float()
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_conv3x3(n=3, H=9, W=16, Cin=64, Cout=64, temb=True, fp32_out=False, seed=0, stride=1):
    """H, W: OUTPUT dims; stride 2 (Downsample2D) reads a [n, 2H, 2W, Cin] input."""
    x = bf(g(n, H * stride, W * stride, Cin, seed=seed))  # NHWC
    w = bf(g(Cout, Cin, 3, 3, seed=seed + 1, scale=(9 * Cin) ** -0.5))
    b = g(Cout, seed=seed + 2)
    tv = g(n, Cout, seed=seed + 3) if temb else None
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = torch.empty(n * H * W, Cout, dtype=torch.float32 if fp32_out else torch.bfloat16, device=DEV)
    lib.gemm(x, wk, out, M=n * H * W, N=Cout, k1=Cin, mode=lib.A_CONV3X3, n_img=n, H=H, W=W, bias=b, rowvec=tv,
             rows_per_vec=H * W, out_fp32=fp32_out, conv_stride=stride)
    ref = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1, stride=stride)
    if temb:
        ref = ref + tv[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1).reshape(n * H * W, Cout)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_upsample_conv(n=3, H=9, W=16, Cin=64, Cout=128, seed=0):
    """Upsample2D (nearest x2 + 3x3 conv) as four 2x2-tap parity convolutions on the low-resolution input + interleave."""
    from this_and_that_vdm_b200.engine import _pack_upsample_parity
    x = bf(g(n, H, W, Cin, seed=seed))
    w = bf(g(Cout, Cin, 3, 3, seed=seed + 1, scale=(9 * Cin) ** -0.5))
    b = g(Cout, seed=seed + 2)
    wp = _pack_upsample_parity(w, DEV)
    M = n * H * W
    parts = torch.empty(4, M, Cout, dtype=torch.bfloat16, device=DEV)
    out = torch.empty(4 * M, Cout, dtype=torch.bfloat16, device=DEV)
    for pi, (py, px) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        lib.gemm(x, wp[pi], parts[pi], M=M, N=Cout, k1=Cin, mode=lib.A_CONV3X3, n_img=n, H=H, W=W, bias=b, conv_taps=4,
                 conv_dy0=py - 1, conv_dx0=px - 1)
    lib.interleave2x(parts, out, n_img=n, H=H, W=W, C=Cout)
    up = Fn.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = Fn.conv2d(up, w.float(), b, padding=1).permute(0, 2, 3, 1).reshape(4 * M, Cout)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_tconv(B=2, F=14, S=24, C=64, seed=0):
    x = bf(g(B, F, S, C, seed=seed))
    w = bf(g(C, C, 3, seed=seed + 1, scale=(3 * C) ** -0.5))  # [Cout, Cin, 3]
    b = g(C, seed=seed + 2)
    tv = g(B * F, C, seed=seed + 3)
    res = bf(g(B * F * S, C, seed=seed + 4))
    wk = w.permute(0, 2, 1).reshape(C, 3 * C).contiguous()
    out = torch.empty(B * F * S, C, dtype=torch.bfloat16, device=DEV)
    lib.gemm(x, wk, out, M=B * F * S, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, rowvec=tv,
             rows_per_vec=S, res1=res, s1=1.0, s0=0.5)
    xr = x.float().permute(0, 3, 1, 2).reshape(B, C, F, S, 1)
    ref = Fn.conv3d(xr, w.float()[:, :, :, None, None], b, padding=(1, 0, 0))
    ref = ref.reshape(B, C, F, S).permute(0, 2, 3, 1).reshape(B * F * S, C) + tv.repeat_interleave(S, 0)
    ref = 0.5 * ref + res.float()
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_attn_spatial(n=3, heads=5, S=384, seed=0):
    C = heads * 64
    qkv = bf(g(n * S, 3 * C, seed=seed))
    out = torch.empty(n * S, C, dtype=torch.bfloat16, device=DEV)
    lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=n,
                     heads=heads, seq=S, scale=0.125)
    q, k, v = [t.float().reshape(n, S, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
    ref = Fn.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n * S, C)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_attn_cross(B=2, F=3, S=96, heads=5, L=78, temporal=False, batch_offset=0, n_ctx=2, seed=0, hd=64):
    C = heads * hd
    rows = B * F * S
    q = bf(g(rows, C, seed=seed))
    kc = bf(g(n_ctx, L, C, seed=seed + 1))
    vc = bf(g(n_ctx, L, C, seed=seed + 2))
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
    lib.attn_cross(q, kc, vc, out, ldq=C, ldo=C, rows=rows, heads=heads, L=L, F=F, S=S, n_ctx=n_ctx,
                   temporal=temporal, batch_offset=batch_offset, scale=hd ** -0.5, head_dim=hd)
    r = torch.arange(rows, device=DEV)
    b = r // (F * S) + batch_offset
    s = r % S
    ctx = ((b * S + s) % n_ctx) if temporal else b
    qh = q.float().reshape(rows, heads, 1, hd)
    kh = kc.float()[ctx].reshape(rows, L, heads, hd).transpose(1, 2)
    vh = vc.float()[ctx].reshape(rows, L, heads, hd).transpose(1, 2)
    ref = Fn.scaled_dot_product_attention(qh, kh, vh).reshape(rows, C)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_attn_temporal(B=2, F=14, S=40, heads=5, seed=0, hd=64):
    C = heads * hd
    rows = B * F * S
    qkv = bf(g(rows, 3 * C, seed=seed))
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
    lib.attn_temporal(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, B=B, F=F, S=S,
                      heads=heads, scale=hd ** -0.5, head_dim=hd)
    q, k, v = [t.float().reshape(B, F, S, heads, hd).permute(0, 2, 3, 1, 4) for t in qkv.split(C, dim=1)]
    ref = Fn.scaled_dot_product_attention(q, k, v).permute(0, 3, 1, 2, 4).reshape(rows, C)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_groupnorm(n_inst=4, rows_per_inst=150, c1=320, c2=0, silu=True, eps=1e-6, seed=0):
    rows = n_inst * rows_per_inst
    x1 = bf(g(rows, c1, seed=seed) * 2 + 0.5)
    x2 = bf(g(rows, c2, seed=seed + 1)) if c2 else None
    C = c1 + c2
    gamma = g(C, seed=seed + 2) * 0.1 + 1
    beta = g(C, seed=seed + 3) * 0.1
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
    stats = torch.empty(n_inst * 64, dtype=torch.float64, device=DEV)
    lib.groupnorm(x1, out, stats, gamma, beta, c1=c1, rows=rows, rows_per_inst=rows_per_inst, eps=eps, silu=silu,
                  x2=x2, c2=c2)
    x = torch.cat([x1, x2], 1).float() if c2 else x1.float()
    xr = x.reshape(n_inst, rows_per_inst, C).permute(0, 2, 1)
    ref = Fn.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = Fn.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(rows, C)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_layernorm(rows=1000, C=320, add=False, F=14, S=10, seed=0):
    x = bf(g(rows, C, seed=seed) * 1.5 + 0.3)
    gamma = g(C, seed=seed + 1) * 0.1 + 1
    beta = g(C, seed=seed + 2) * 0.1
    av = g(F, C, seed=seed + 3) if add else None
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
    so = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV) if add else None
    lib.layernorm(x, out, gamma, beta, rows=rows, C=C, addvec=av, F=F, S=S, sum_out=so)
    xf = x.float()
    err2 = 0.0
    if add:
        xf = xf + av[(torch.arange(rows, device=DEV) // S) % F]
        err2 = rel_l2(so, xf)
        xf = bf(xf).float()
    ref = Fn.layer_norm(xf, (C,), gamma, beta, 1e-5)
    torch.cuda.synchronize()
    return max(rel_l2(out, ref), err2)


def case_im2col_s2(n=2, H=8, W=12, C=64, seed=0):
    x = bf(g(n, H, W, C, seed=seed))
    w = bf(g(C, C, 3, 3, seed=seed + 1, scale=(9 * C) ** -0.5))
    col = torch.empty(n * (H // 2) * (W // 2), 9 * C, dtype=torch.bfloat16, device=DEV)
    lib.im2col_s2(x, col, n_img=n, H=H, W=W, C=C)
    wk = w.permute(0, 2, 3, 1).reshape(C, 9 * C).contiguous()
    M = n * (H // 2) * (W // 2)
    out = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    lib.gemm(col, wk, out, M=M, N=C, k1=9 * C)
    ref = Fn.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1).permute(0, 2, 3, 1).reshape(M, C)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_upsample(n=2, H=5, W=6, C=64, seed=0):
    x = bf(g(n, H, W, C, seed=seed))
    out = torch.empty(n, 2 * H, 2 * W, C, dtype=torch.bfloat16, device=DEV)
    lib.upsample2x(x, out, n_img=n, H=H, W=W, C=C)
    ref = Fn.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_axpy(n=8 * 1000, seed=0):
    a, b = bf(g(n, seed=seed)), bf(g(n, seed=seed + 1))
    out = torch.empty_like(a)
    lib.axpy(a, b, out, 0.5, n)
    torch.cuda.synchronize()
    return rel_l2(out, a.float() + 0.5 * b.float())


def case_sampler(F=14, h=8, w=12, seed=0):
    lat = g(F, 4, h, w, seed=seed) * 50
    img = g(2, 4, h, w, seed=seed + 1)
    cond = g(F, 4, h, w, seed=seed + 2)
    sigma, sigma_next = 50.0, 30.0
    mi = torch.empty(2, F, h, w, 64, dtype=torch.bfloat16, device=DEV)
    lib.sampler_prepare(lat, img, cond, mi, c_pad=64, B_local=2, batch_offset=0, F=F, h=h, w=w, sigma=sigma)
    ref = torch.zeros(2, F, h, w, 64, device=DEV)
    ref[..., 0:4] = (lat / math.sqrt(sigma ** 2 + 1)).permute(0, 2, 3, 1)[None]
    ref[..., 4:8] = img.permute(0, 2, 3, 1)[:, None]
    ref[..., 8:12] = cond.permute(0, 2, 3, 1)[None]
    e1 = rel_l2(mi, ref)
    eu, ec = g(F * h * w, 4, seed=seed + 3), g(F * h * w, 4, seed=seed + 4)
    gd = torch.linspace(1, 3, F, device=DEV)
    lat2 = lat.clone()
    lib.sampler_euler_step(lat2, eu, ec, gd, ld_eps=4, F=F, h=h, w=w, sigma=sigma, sigma_next=sigma_next)
    eps = eu + gd.repeat_interleave(h * w)[:, None] * (ec - eu)
    eps = eps.reshape(F, h, w, 4).permute(0, 3, 1, 2)
    x0 = eps * (-sigma / math.sqrt(sigma ** 2 + 1)) + lat / (sigma ** 2 + 1)
    refl = lat + (lat - x0) / sigma * (sigma_next - sigma)
    torch.cuda.synchronize()
    return max(e1, rel_l2(lat2, refl))


# ------------------------------------------------------------------------------------------------ norm fusions (ABI 3)
def _gn_ref(x, n_inst, rpi, C, gamma, beta, eps, silu):
    xr = x.float().reshape(n_inst, rpi, C).permute(0, 2, 1)
    ref = Fn.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        ref = Fn.silu(ref)
    return ref.permute(0, 2, 1).reshape(n_inst * rpi, C)


def case_gemm_gn_stats(mode="linear", M=0, N=320, K=320, rpi=0, n=2, H=24, W=40, Cin=64, B=2, F=5, S=200, res=True,
                       concat=False, seed=0):
    """A GEMM whose epilogue accumulates the GroupNorm pair sums of its output, followed by ttvdm_groupnorm that uses
    them (no statistics pass): compared with torch group_norm of the GEMM's own bf16 output. Also returns the error of
    the pair sums themselves against fp64 sums of the stored tensor (max of the two)."""
    if mode == "linear":
        a = bf(g(M, K, seed=seed)); w = bf(g(N, K, seed=seed + 1, scale=K ** -0.5)); b = g(N, seed=seed + 2)
        r1 = bf(g(M, N, seed=seed + 3)) if res else None
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        st = torch.zeros((M // rpi) * N, dtype=torch.float64, device=DEV)
        lib.gemm(a, w, out, M=M, N=N, k1=K, bias=b, res1=r1, gn_stats_out=st, gn_rows_per_inst=rpi)
    elif mode == "conv":
        M, rpi_ = n * H * W, (rpi or H * W)
        rpi = rpi_
        x = bf(g(n, H, W, Cin, seed=seed)); wk = bf(g(N, 9 * Cin, seed=seed + 1, scale=(9 * Cin) ** -0.5)); b = g(N, seed=seed + 2)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        st = torch.zeros((M // rpi) * N, dtype=torch.float64, device=DEV)
        lib.gemm(x, wk, out, M=M, N=N, k1=Cin, mode=lib.A_CONV3X3, n_img=n, H=H, W=W, bias=b, gn_stats_out=st,
                 gn_rows_per_inst=rpi)
    else:
        M = B * F * S
        rpi = rpi or F * S
        x = bf(g(B, F, S, N, seed=seed)); wk = bf(g(N, 3 * N, seed=seed + 1, scale=(3 * N) ** -0.5)); b = g(N, seed=seed + 2)
        r1 = bf(g(M, N, seed=seed + 3))
        out = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        st = torch.zeros((M // rpi) * N, dtype=torch.float64, device=DEV)
        lib.gemm(x, wk, out, M=M, N=N, k1=N, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, res1=r1, gn_stats_out=st,
                 gn_rows_per_inst=rpi)
    n_inst = M // rpi
    v = out.double().view(n_inst, rpi, N // 2, 2)
    ref_st = torch.stack([v.sum(dim=(1, 3)), (v * v).sum(dim=(1, 3))], dim=-1).reshape(-1)
    e_st = float((st - ref_st).abs().max() / ref_st.abs().max())
    # consumer: GroupNorm (+ optional second source WITHOUT producer statistics: mixed mode)
    c2 = 64 if concat else 0
    x2 = bf(g(M, c2, seed=seed + 7)) if concat else None
    C = N + c2
    gamma = g(C, seed=seed + 4) * 0.1 + 1
    beta = g(C, seed=seed + 5) * 0.1
    y = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    ws = torch.empty(n_inst * 64, dtype=torch.float64, device=DEV) if concat else None
    n0 = lib.launch_count()
    lib.groupnorm(out, y, ws, gamma, beta, c1=N, rows=M, rows_per_inst=rpi, eps=1e-6, silu=True, x2=x2, c2=c2, pstats1=st)
    launches = lib.launch_count() - n0
    assert launches == (2 if concat else 1), f"groupnorm launched {launches} kernels"
    xin = torch.cat([out, x2], 1) if concat else out
    ref = _gn_ref(xin, n_inst, rpi, C, gamma, beta, 1e-6, True)
    torch.cuda.synchronize()
    return max(rel_l2(y, ref), e_st * 100)  # e_st must be < 1e-4 * ... (fp32 partials): scaled so TOL applies


def case_gemm_ln_fold(M=1000, C=320, N=960, geglu=False, posemb=False, F=7, S=0, direct_producer=False, seed=0):
    """Producer GEMM writes h (+ row sums, optionally over h + positional embedding); consumer GEMM applies the folded
    LayerNorm in its epilogue. Reference: torch layer_norm of the producer's bf16 output, then the linear (+ GEGLU)."""
    Kp = 4 * C if direct_producer else C
    a = bf(g(M, Kp, seed=seed)); wp = bf(g(C, Kp, seed=seed + 1, scale=Kp ** -0.5)); bp = g(C, seed=seed + 2)
    r1 = bf(g(M, C, seed=seed + 3) * 2)
    h = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    rs = torch.empty(C // 32, M, 2, dtype=torch.float32, device=DEV)
    pos = g(F, C, seed=seed + 9) * 0.5 if posemb else None
    S = S or (M // F if posemb else 0)
    lib.gemm(a, wp, h, M=M, N=C, k1=Kp, bias=bp, res1=r1, row_sums_out=rs, rs_addvec=pos, rs_add_rows=S, rs_add_mod=F)
    gamma = g(C, seed=seed + 4) * 0.2 + 1
    beta = g(C, seed=seed + 5) * 0.2
    w = g(N, C, seed=seed + 6, scale=C ** -0.5)
    b = g(N, seed=seed + 7)
    wf = bf(w * gamma[None, :])
    bias_f = (b + w @ beta).contiguous()
    cs = wf.float().sum(1).contiguous()
    out = torch.empty(M, N // 2 if geglu else N, dtype=torch.bfloat16, device=DEV)
    kw = {}
    if posemb:
        pv = (bf(pos).float() @ wf.float().t()).contiguous()
        kw = dict(prevec=pv, prevec_rows=S, prevec_mod=F)
    lib.gemm(h, wf, out, M=M, N=N, k1=C, bias=bias_f, geglu=geglu, ln_rowsums=rs, ln_colsum=cs, ln_eps=1e-5, **kw)
    x = h.float()
    if posemb:
        x = x + pos[(torch.arange(M, device=DEV) // S) % F]
    ref = Fn.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + b
    if geglu:
        ref = ref[:, 0::2] * Fn.gelu(ref[:, 1::2])
    torch.cuda.synchronize()
    return rel_l2(out, ref)


def case_pack(dtype=torch.float32, seed=0):
    """Repack entry points against plain torch: conv tap-major layout with channel padding, LayerNorm-folded GEGLU
    interleave (+ folded bias, column sums), fused row offsets, fp32 vector cast. Returns the worst relative error."""
    errs = []
    w = g(40, 24, 3, 3, seed=seed).to(dtype)
    out = torch.empty(40, 9 * 64, dtype=torch.bfloat16, device=DEV)
    lib.pack_conv_weight(w, out, cin_pad=64)
    ref = torch.zeros(40, 3, 3, 64, device=DEV)
    ref[..., :24] = w.float().permute(0, 2, 3, 1)
    errs.append(rel_l2(out, bf(ref.reshape(40, -1))))
    tw = g(48, 48, 3, 1, 1, seed=seed + 1).to(dtype)
    out = torch.empty(48, 3 * 48, dtype=torch.bfloat16, device=DEV)
    lib.pack_conv_weight(tw, out)
    errs.append(rel_l2(out, bf(tw.float().reshape(48, 48, 3).permute(0, 2, 1).reshape(48, -1))))
    N, K = 96, 320
    lw, lb = g(N, K, seed=seed + 2, scale=K ** -0.5).to(dtype), g(N, seed=seed + 3).to(dtype)
    gamma, beta = (g(K, seed=seed + 4) * 0.2 + 1).to(dtype), (g(K, seed=seed + 5) * 0.2).to(dtype)
    ow = torch.zeros(N + 10, K, dtype=torch.bfloat16, device=DEV)
    ob = torch.zeros(N + 10, dtype=torch.float32, device=DEV)
    oc = torch.zeros(N + 10, dtype=torch.float32, device=DEV)
    lib.pack_linear(lw, ow, bias=lb, gamma=gamma, beta=beta, out_bias=ob, out_colsum=oc, geglu=True, out_row0=10)
    order = torch.stack([torch.arange(N // 2), torch.arange(N // 2) + N // 2], 1).reshape(-1).to(DEV)
    wf = bf(lw.float() * gamma.float()[None, :])[order]
    errs.append(rel_l2(ow[10:], wf))
    errs.append(rel_l2(ob[10:], (lb.float() + lw.float() @ beta.float())[order]))
    errs.append(rel_l2(oc[10:], wf.float().sum(1)))
    errs.append(float(ow[:10].float().abs().max()))  # rows before out_row0 untouched
    v = g(777, seed=seed + 6).to(dtype)
    o = torch.empty(777, dtype=torch.float32, device=DEV)
    lib.pack_vector(v, o)
    errs.append(rel_l2(o, v.float()))
    torch.cuda.synchronize()
    return max(errs)


CASES = [
    ("pack_fp32", lambda: case_pack(torch.float32)),
    ("pack_fp16", lambda: case_pack(torch.float16)),
    ("pack_bf16", lambda: case_pack(torch.bfloat16)),
    ("gemm_gnstats_linear_tma", lambda: case_gemm_gn_stats("linear", M=28 * 150, N=320, K=320, rpi=150)),
    ("gemm_gnstats_linear_straddle", lambda: case_gemm_gn_stats("linear", M=20 * 24, N=1280, K=1280, rpi=24)),
    ("gemm_gnstats_linear_direct", lambda: case_gemm_gn_stats("linear", M=7 * 144, N=320, K=2880, rpi=144)),
    ("gemm_gnstats_linear_mixed_concat", lambda: case_gemm_gn_stats("linear", M=6 * 100, N=64, K=64, rpi=100, concat=True)),
    ("gemm_gnstats_conv_direct", lambda: case_gemm_gn_stats("conv", n=3, H=18, W=32, Cin=320, N=320)),
    ("gemm_gnstats_conv_tma_5d", lambda: case_gemm_gn_stats("conv", n=4, H=9, W=16, Cin=64, N=128, rpi=2 * 9 * 16)),
    ("gemm_gnstats_conv_ragged", lambda: case_gemm_gn_stats("conv", n=2, H=7, W=13, Cin=64, N=64)),
    ("gemm_gnstats_tconv_5d", lambda: case_gemm_gn_stats("tconv", B=2, F=5, S=200, N=320)),
    ("gemm_gnstats_tconv_4d", lambda: case_gemm_gn_stats("tconv", B=2, F=4, S=96, N=640, rpi=96)),
    ("gemm_gnstats_tconv_direct", lambda: case_gemm_gn_stats("tconv", B=1, F=3, S=150, N=1280)),
    ("gemm_lnfold_qkv", lambda: case_gemm_ln_fold()),
    ("gemm_lnfold_q_640", lambda: case_gemm_ln_fold(M=777, C=640, N=640)),
    ("gemm_lnfold_geglu", lambda: case_gemm_ln_fold(M=515, C=320, N=2560, geglu=True)),
    ("gemm_lnfold_geglu_posemb", lambda: case_gemm_ln_fold(M=7 * 90, C=320, N=2560, geglu=True, posemb=True)),
    ("gemm_lnfold_direct_producer_posemb", lambda: case_gemm_ln_fold(M=7 * 64, C=320, N=2560, geglu=True, posemb=True,
                                                                     direct_producer=True)),
    ("gemm_lnfold_1280", lambda: case_gemm_ln_fold(M=300, C=1280, N=3840)),
    ("gemm_lnfold_c64", lambda: case_gemm_ln_fold(M=200, C=64, N=192)),
] + [
    ("gemm_linear_320", lambda: case_gemm_linear()),
    ("gemm_linear_bigN_res", lambda: case_gemm_linear(M=777, N=1280, K=640, res=True)),
    ("gemm_linear_geglu", lambda: case_gemm_linear(M=515, N=2560, K=320, geglu=True)),
    ("gemm_linear_geglu_bresident", lambda: case_gemm_linear(M=128 * 117 + 5, N=2560, K=320, geglu=True)),
    ("gemm_linear_concatK", lambda: case_gemm_linear(M=300, N=640, K=1280, a2=True)),
    ("gemm_linear_raggedN", lambda: case_gemm_linear(M=130, N=200, K=64)),
    # N % 32 != 0 with BOTH residuals (the last 32-column chunk must not lose s2 * res2)
    ("gemm_linear_raggedN_res12", lambda: case_gemm_linear(M=300, N=200, K=128, res=True)),
    ("gemm_linear_res_partialN", lambda: case_gemm_linear(M=128 * 300 + 17, N=320, K=320, res=True)),
    ("gemm_linear_res_640", lambda: case_gemm_linear(M=128 * 150, N=640, K=640, res=True)),
    ("gemm_linear_multitile", lambda: case_gemm_linear(M=128 * 200, N=1920, K=640, bias=False)),
    ("conv3x3_small", lambda: case_conv3x3()),
    # Downsample2D: stride 2, padding 1, through a TMA box with element stride 2 (no im2col)
    ("conv3x3_stride2_small", lambda: case_conv3x3(n=3, H=9, W=16, Cin=64, Cout=64, temb=False, stride=2)),
    ("conv3x3_stride2_L0", lambda: case_conv3x3(n=2, H=36, W=64, Cin=320, Cout=320, temb=False, stride=2)),
    ("conv3x3_stride2_odd_tiles", lambda: case_conv3x3(n=2, H=5, W=7, Cin=128, Cout=192, temb=False, stride=2)),
    ("conv3x3_stride2_L2", lambda: case_conv3x3(n=4, H=9, W=16, Cin=1280, Cout=1280, temb=False, stride=2)),
    ("upsample_conv_small", lambda: case_upsample_conv()),
    ("upsample_conv_odd", lambda: case_upsample_conv(n=2, H=5, W=7, Cin=128, Cout=64)),
    ("upsample_conv_L1", lambda: case_upsample_conv(n=2, H=18, W=32, Cin=1280, Cout=1280)),
    ("conv3x3_L0", lambda: case_conv3x3(n=2, H=32, W=48, Cin=320, Cout=320)),
    ("conv3x3_wide", lambda: case_conv3x3(n=1, H=7, W=130, Cin=128, Cout=192)),
    ("conv3x3_out4_fp32", lambda: case_conv3x3(n=2, H=8, W=12, Cin=320, Cout=4, temb=False, fp32_out=True)),
    ("tconv_small", lambda: case_tconv()),
    ("tconv_L1", lambda: case_tconv(B=2, F=14, S=384, C=640)),
    ("attn_spatial_384", lambda: case_attn_spatial()),
    ("attn_spatial_ragged", lambda: case_attn_spatial(n=2, heads=2, S=200)),
    ("attn_spatial_tiny", lambda: case_attn_spatial(n=4, heads=20, S=24)),
    ("attn_spatial_1536", lambda: case_attn_spatial(n=2, heads=5, S=1536)),
    ("attn_cross_spatial", lambda: case_attn_cross()),
    ("attn_cross_spatial_L1", lambda: case_attn_cross(L=1)),
    ("attn_cross_temporal", lambda: case_attn_cross(temporal=True)),
    ("attn_cross_temporal_shard", lambda: case_attn_cross(B=1, temporal=True, batch_offset=1, S=97 - 1)),
    ("attn_cross_temporal_oddS", lambda: case_attn_cross(B=2, F=2, S=135, temporal=True)),
    ("attn_cross_spatial_L128", lambda: case_attn_cross(B=2, F=2, S=50, heads=3, L=128)),
    ("attn_cross_spatial_L17_tinyS", lambda: case_attn_cross(B=3, F=2, S=5, heads=2, L=17, n_ctx=3)),
    ("attn_cross_temporal_3ctx", lambda: case_attn_cross(B=3, F=2, S=70, heads=2, L=33, temporal=True, n_ctx=3)),
    ("attn_cross_temporal_3ctx_shard", lambda: case_attn_cross(B=1, F=3, S=49, heads=2, L=64, temporal=True, n_ctx=3, batch_offset=2)),
    # S / n_ctx >= 512 rows per (unit, context): these take the tcgen05 / TMA cross mode of the flash kernel
    ("attn_cross_spatial_tc", lambda: case_attn_cross(B=2, F=2, S=600, heads=5, L=78)),
    ("attn_cross_temporal_tc", lambda: case_attn_cross(B=2, F=2, S=1100, heads=5, L=78, temporal=True)),
    ("attn_cross_temporal_tc_oddS_shard", lambda: case_attn_cross(B=1, F=3, S=1101, heads=2, L=77, temporal=True, batch_offset=1)),
    ("attn_cross_spatial_tc_L128_shard", lambda: case_attn_cross(B=1, F=2, S=520, heads=3, L=128, batch_offset=1)),
    ("attn_cross_spatial_tc_L1", lambda: case_attn_cross(B=2, F=1, S=512, heads=2, L=1)),
    # head_dim 128 (reference class-default heads at level 2): warp-level kernels, templated on the head dim
    ("attn_cross_spatial_hd128", lambda: case_attn_cross(B=2, F=3, S=96, heads=3, L=78, hd=128)),
    ("attn_cross_temporal_hd128_oddS", lambda: case_attn_cross(B=2, F=2, S=135, heads=2, L=77, temporal=True, hd=128)),
    ("attn_cross_temporal_hd128_big", lambda: case_attn_cross(B=2, F=2, S=1100, heads=2, L=78, temporal=True, hd=128)),
    ("attn_temporal_hd128", lambda: case_attn_temporal(B=2, F=14, S=37, heads=3, hd=128)),
    ("attn_temporal_hd128_F4", lambda: case_attn_temporal(B=1, F=4, S=9, heads=2, hd=128)),
    ("attn_temporal_F25_svd_xt", lambda: case_attn_temporal(B=2, F=25, S=21, heads=5)),
    ("attn_temporal_F17", lambda: case_attn_temporal(B=1, F=17, S=10, heads=2)),
    ("attn_temporal_F32", lambda: case_attn_temporal(B=1, F=32, S=6, heads=3)),
    ("attn_temporal_F25_hd128", lambda: case_attn_temporal(B=1, F=25, S=11, heads=2, hd=128)),
    ("attn_temporal", lambda: case_attn_temporal()),
    ("attn_temporal_F16", lambda: case_attn_temporal(B=1, F=16, S=33, heads=2)),
    ("attn_temporal_F3_ragged", lambda: case_attn_temporal(B=2, F=3, S=7, heads=3)),
    ("layernorm_generic_C200", lambda: case_layernorm(rows=333, C=200)),
    ("layernorm_320_add", lambda: case_layernorm(rows=14 * 9 * 2 + 0, C=320, add=True, S=9)),
    ("groupnorm_4d", lambda: case_groupnorm()),
    ("groupnorm_c64", lambda: case_groupnorm(n_inst=3, rows_per_inst=77, c1=64, silu=False)),
    ("groupnorm_concat", lambda: case_groupnorm(c1=1280, c2=640, silu=True)),
    ("groupnorm_5d_nosilu", lambda: case_groupnorm(n_inst=2, rows_per_inst=14 * 96, c1=640, silu=False, eps=1e-5)),
    ("layernorm_320", lambda: case_layernorm()),
    ("layernorm_1280_add", lambda: case_layernorm(rows=14 * 10 * 3, C=1280, add=True)),
    ("layernorm_640", lambda: case_layernorm(C=640)),
    ("im2col_s2+gemm", lambda: case_im2col_s2()),
    ("upsample2x", lambda: case_upsample()),
    ("axpy", lambda: case_axpy()),
    ("sampler_glue", lambda: case_sampler()),
]

TOL = 1.5e-2  # rel-L2 of a bf16-output kernel vs the fp32 torch op on identical bf16 inputs (bf16 eps = 7.8e-3)


def time_it(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timings():
    res = {}
    # GEMM shapes of the 576x1024 workload (M = 28*S)
    for name, M, N, K, geglu in [
        ("gemm_L0_cxc", 258048, 320, 320, False), ("gemm_L0_geglu", 258048, 2560, 320, True),
        ("gemm_L0_ffout", 258048, 320, 1280, False), ("gemm_L1_qkv", 64512, 1920, 640, False),
        ("gemm_L2_geglu", 16128, 10240, 1280, True), ("gemm_L2_ffout", 16128, 1280, 5120, False),
    ]:
        a = bf(g(M, K)); w = bf(g(N, K, scale=K ** -0.5)); b = g(N)
        out = torch.empty(M, N // 2 if geglu else N, dtype=torch.bfloat16, device=DEV)
        ms = time_it(lambda: lib.gemm(a, w, out, M=M, N=N, k1=K, bias=b, geglu=geglu))
        res[name] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
        a2 = a.float(); del a2
        ms_t = time_it(lambda: Fn.linear(a, w))
        res[name]["torch_ms"] = ms_t
    for name, n, H, W, Ci, Co in [("conv_L0", 28, 72, 128, 320, 320), ("conv_L1", 28, 36, 64, 640, 640),
                                   ("conv_L2", 28, 18, 32, 1280, 1280), ("conv_L3", 28, 9, 16, 1280, 1280)]:
        x = bf(g(n, H, W, Ci)); wk = bf(g(Co, 9 * Ci, scale=(9 * Ci) ** -0.5)); b = g(Co)
        out = torch.empty(n * H * W, Co, dtype=torch.bfloat16, device=DEV)
        ms = time_it(lambda: lib.gemm(x, wk, out, M=n * H * W, N=Co, k1=Ci, mode=lib.A_CONV3X3, n_img=n, H=H, W=W, bias=b))
        res[name] = {"ms": ms, "tflops": 2.0 * n * H * W * Co * 9 * Ci / ms / 1e9}
        xc = x.permute(0, 3, 1, 2); wc = wk.reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        res[name]["torch_ms"] = time_it(lambda: Fn.conv2d(xc, wc, None, padding=1))
    for name, n, heads, S in [("attn_L0", 28, 5, 9216), ("attn_L1", 28, 10, 2304), ("attn_L2", 28, 20, 576)]:
        C = heads * 64
        qkv = bf(g(n * S, 3 * C)); out = torch.empty(n * S, C, dtype=torch.bfloat16, device=DEV)
        ms = time_it(lambda: lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C,
                                              ldo=C, n_img=n, heads=heads, seq=S, scale=0.125), iters=5)
        res[name] = {"ms": ms, "tflops": 4.0 * n * heads * S * S * 64 / ms / 1e9}
        q, k, v = [t.reshape(n, S, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
        res[name]["torch_ms"] = time_it(lambda: Fn.scaled_dot_product_attention(q, k, v), iters=5)
    M, C = 258048, 320
    x = bf(g(M, C)); gm = g(C); bt = g(C); out = torch.empty_like(x)
    stats = torch.empty(28 * 64, dtype=torch.float64, device=DEV)
    ms = time_it(lambda: lib.groupnorm(x, out, stats, gm, bt, c1=C, rows=M, rows_per_inst=9216, eps=1e-6, silu=True))
    res["groupnorm_L0"] = {"ms": ms, "gbs": M * C * 2 * 3 / ms / 1e6}
    ms = time_it(lambda: lib.layernorm(x, out, gm, bt, rows=M, C=C))
    res["layernorm_L0"] = {"ms": ms, "gbs": M * C * 2 * 2 / ms / 1e6}
    qkv = bf(g(M, 3 * C))
    ms = time_it(lambda: lib.attn_temporal(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C,
                                           B=2, F=14, S=9216, heads=5, scale=0.125))
    res["attn_temporal_L0"] = {"ms": ms, "gbs": M * C * 2 * 4 / ms / 1e6}
    kc = bf(g(2, 78, C)); vc = bf(g(2, 78, C))
    ms = time_it(lambda: lib.attn_cross(x, kc, vc, out, ldq=C, ldo=C, rows=M, heads=5, L=78, F=14, S=9216, n_ctx=2,
                                        temporal=True, batch_offset=0, scale=0.125))
    res["attn_cross_temporal_L0"] = {"ms": ms, "gbs": M * C * 2 * 2 / ms / 1e6}
    return res


def main():
    lib.init()
    print("device:", torch.cuda.get_device_name(0), flush=True)
    results = {}
    only = sys.argv[1:] if len(sys.argv) > 1 and sys.argv[1] != "--time" else None
    for name, fn in CASES:
        if only and not any(o in name for o in only):
            continue
        t0 = time.time()
        try:
            err = fn()
            ok = err == err and err < TOL
            results[name] = err
            print(f"{'PASS' if ok else 'FAIL'} {name:32s} rel_l2={err:.3e}  ({time.time() - t0:.2f}s)", flush=True)
        except Exception as e:  # noqa: BLE001
            results[name] = f"EXC {e}"
            print(f"EXC  {name:32s} {e}", flush=True)
            traceback.print_exc()
    out_dir = Path(__file__).resolve().parents[1] / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    (out_dir / "kernel_check.json").write_text(json.dumps(results, indent=1))
    if "--time" in sys.argv:
        try:
            t = timings()
            for k, v in t.items():
                print("TIME", k, json.dumps(v), flush=True)
            (out_dir / "kernel_times.json").write_text(json.dumps(t, indent=1))
        except Exception:  # noqa: BLE001
            traceback.print_exc()


if __name__ == "__main__":
    main()
