"""Algorithmic FLOP census of the SVD UNet + GestureNet hot path (SURVEY.md Appendix B restated as a module).

2*MAC convention; cross-attention K/V of the constant context counted once per (batch element, token); norms and
elementwise ops count 0. `census(h, w, B)` walks the same structure the engine executes and is the single source
of the numerator of bench.py's `roofline.achieved`. Known answers: UNet 72x128 B=2 = 89.854 TF,
GestureNet = 32.953 TF, VGL video (25 steps) = 3070.2 TF; params 1,524,623,082 / 680,946,577.
"""
from __future__ import annotations

import collections
import sys

F_FRAMES, L_CTX, D_CTX, TEMB = 14, 78, 1024, 1280
HEADS = (5, 10, 20, 20)
CHANS = (320, 640, 1280, 1280)


def census(h: int, w: int, B: int, controlnet: bool = False, frames: int = F_FRAMES, L: int = L_CTX):
    fl, pr = collections.Counter(), collections.Counter()
    BF = B * frames

    def conv2d(cin, cout, k, hh, ww, tag="conv2d", stride=1):
        fl[tag] += 2 * BF * (hh // stride) * (ww // stride) * cin * cout * k * k
        pr[tag] += cin * cout * k * k + cout

    def conv_t(c, hh, ww):
        fl["conv_temporal"] += 2 * BF * hh * ww * c * c * 3
        pr["conv_temporal"] += c * c * 3 + c

    def lin(rows, cin, cout, tag, bias=True):
        fl[tag] += 2 * rows * cin * cout
        pr[tag] += cin * cout + (cout if bias else 0)

    def stres(cin, cout, hh, ww):
        pr["norm"] += 2 * cin + 2 * cout + 4 * cout
        conv2d(cin, cout, 3, hh, ww)
        lin(BF, TEMB, cout, "temb_proj")
        conv2d(cout, cout, 3, hh, ww)
        if cin != cout:
            conv2d(cin, cout, 1, hh, ww, "conv1x1_shortcut")
        conv_t(cout, hh, ww)
        lin(BF, TEMB, cout, "temb_proj")
        conv_t(cout, hh, ww)
        pr["alpha"] += 1

    def transformer(C, hh, ww):
        S = hh * ww
        rows = BF * S
        pr["norm"] += 2 * C + 14 * C
        lin(rows, C, C, "proj_in_out")
        for _ in range(3):
            lin(rows, C, C, "attn_qkvo", bias=False)
        lin(rows, C, C, "attn_qkvo")
        fl["spatial_self_attn"] += 4 * BF * S * S * C
        lin(rows, C, C, "attn_qkvo", bias=False)
        lin(rows, C, C, "attn_qkvo")
        lin(B * L, D_CTX, C, "xattn_kv", bias=False)
        lin(B * L, D_CTX, C, "xattn_kv", bias=False)
        fl["spatial_cross_attn"] += 4 * BF * S * L * C
        lin(rows, C, 8 * C, "ff")
        lin(rows, 4 * C, C, "ff")
        lin(BF, C, 4 * C, "time_pos_embed")
        lin(BF, 4 * C, C, "time_pos_embed")
        lin(rows, C, 8 * C, "ff")
        lin(rows, 4 * C, C, "ff")
        for _ in range(3):
            lin(rows, C, C, "attn_qkvo", bias=False)
        lin(rows, C, C, "attn_qkvo")
        fl["temporal_self_attn"] += 4 * B * S * frames * frames * C
        lin(rows, C, C, "attn_qkvo", bias=False)
        lin(rows, C, C, "attn_qkvo")
        lin(B * L, D_CTX, C, "xattn_kv", bias=False)
        lin(B * L, D_CTX, C, "xattn_kv", bias=False)
        fl["temporal_cross_attn"] += 4 * B * S * frames * L * C
        lin(rows, C, 8 * C, "ff")
        lin(rows, 4 * C, C, "ff")
        pr["alpha"] += 1
        lin(rows, C, C, "proj_in_out")

    lin(B, 320, TEMB, "time_embed")
    lin(B, TEMB, TEMB, "time_embed")
    lin(B, 768, TEMB, "time_embed")
    lin(B, TEMB, TEMB, "time_embed")
    hh, ww = h, w
    conv2d(12 if controlnet else 8, 320, 3, hh, ww)
    skips = [(320, hh, ww)]
    cout = 320
    for i in range(4):
        cin, cout = cout, CHANS[i]
        for j in range(2):
            stres(cin if j == 0 else cout, cout, hh, ww)
            if i < 3:
                transformer(cout, hh, ww)
            skips.append((cout, hh, ww))
        if i < 3:
            conv2d(cout, cout, 3, hh, ww, "conv2d", stride=2)
            hh //= 2
            ww //= 2
            skips.append((cout, hh, ww))
    stres(1280, 1280, hh, ww)
    transformer(1280, hh, ww)
    stres(1280, 1280, hh, ww)
    if controlnet:
        for (c, a, b) in skips:
            conv2d(c, c, 1, a, b, "zero_conv1x1")
        conv2d(1280, 1280, 1, hh, ww, "zero_conv1x1")
        return fl, pr
    rev = list(reversed(CHANS))
    prev = rev[0]
    for i in range(4):
        out = rev[i]
        for j in range(3):
            sc, _, _ = skips.pop()
            stres((prev if j == 0 else out) + sc, out, hh, ww)
            if i > 0:
                transformer(out, hh, ww)
        prev = out
        if i < 3:
            hh *= 2
            ww *= 2
            conv2d(out, out, 3, hh, ww)
    pr["norm"] += 2 * 320
    conv2d(320, 4, 3, hh, ww)
    return fl, pr


def step_flops(h: int, w: int, B: int = 2, vgl: bool = True) -> float:
    fu, _ = census(h, w, B)
    tot = sum(fu.values())
    if vgl:
        fc, _ = census(h, w, B, controlnet=True)
        tot += sum(fc.values())
    return float(tot)


def step_flops_split(h: int, w: int, B: int = 2, vgl: bool = True):
    """(flops executed by the GEMM/conv kernel family, flops executed by the flash-attention kernel, rest)."""
    tot = collections.Counter()
    for cn in ([False, True] if vgl else [False]):
        f, _ = census(h, w, B, controlnet=cn)
        tot.update(f)
    attn = tot["spatial_self_attn"] + tot["spatial_cross_attn"] + tot["temporal_cross_attn"]
    rest = tot["temporal_self_attn"]
    gemm = sum(tot.values()) - attn - rest
    return float(gemm), float(attn), float(rest)


if __name__ == "__main__":
    for (h, w) in [(32, 48), (72, 128)]:
        for B in (1, 2):
            fu, pu = census(h, w, B)
            fc, pc = census(h, w, B, controlnet=True)
            print(f"{h}x{w} B={B}: UNet {sum(fu.values()) / 1e12:.3f} TF ({sum(pu.values()):,} params)  "
                  f"GestureNet {sum(fc.values()) / 1e12:.3f} TF ({sum(pc.values()):,} params)  "
                  f"VGL step {(sum(fu.values()) + sum(fc.values())) / 1e12:.3f} TF  video {25 * (sum(fu.values()) + sum(fc.values())) / 1e12:.1f} TF")
    if "-v" in sys.argv:
        fu, _ = census(72, 128, 2)
        for k, v in sorted(fu.items(), key=lambda kv: -kv[1]):
            print(f"   {k:22s} {v / 1e12:9.4f} TF")


# ---------------------------------------------------------------------------------------------- VAE
def vae_decode_flops(h: int, w: int, frames: int, chans=(128, 256, 512, 512), layers_per_block: int = 2) -> float:
    """Algorithmic FLOPs (2 x MACs; norms / activations = 0) of AutoencoderKLTemporalDecoder.decode for `frames` latent
    frames of h x w: conv_in, mid block (SpatioTemporalResBlock, 1-head attention, SpatioTemporalResBlock), 4 up blocks
    of (layers_per_block + 1) SpatioTemporalResBlocks + nearest x2 conv, conv_out 128->3, time_conv_out."""
    def st_res(s, cin, cout):
        f = 2.0 * s * 9 * (cin * cout + cout * cout)          # spatial conv1 + conv2
        if cin != cout:
            f += 2.0 * s * cin * cout                          # 1x1 shortcut
        return f + 2.0 * s * 3 * cout * cout * 2               # two (3,1,1) temporal convs

    s = h * w
    top = chans[-1]
    f = 2.0 * s * 9 * 4 * top                                  # conv_in
    f += st_res(s, top, top)
    for _ in range(layers_per_block - 1):
        f += 4 * 2.0 * s * top * top + 4.0 * s * s * top       # q, k, v, out linears + QK^T + PV (one head)
        f += st_res(s, top, top)
    rev = list(reversed(chans))
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(layers_per_block + 1):
            f += st_res(s, c if j == 0 else co, co)
        c = co
        if i != len(rev) - 1:
            s *= 4
            f += 2.0 * s * 9 * co * co                         # Upsample2D conv
    f += 2.0 * s * 9 * chans[0] * 3 + 2.0 * s * 3 * 3 * 3      # conv_out + time_conv_out
    return f * frames


def vae_encode_flops(H: int, W: int, images: int, chans=(128, 256, 512, 512), layers_per_block: int = 2) -> float:
    """Same for AutoencoderKLTemporalDecoder.encode of `images` H x W images (quant_conv folded into conv_out)."""
    def res(s, cin, cout):
        return 2.0 * s * 9 * (cin * cout + cout * cout) + (2.0 * s * cin * cout if cin != cout else 0.0)

    s = H * W
    f = 2.0 * s * 9 * 3 * chans[0]
    c = chans[0]
    for i, co in enumerate(chans):
        for j in range(layers_per_block):
            f += res(s, c if j == 0 else co, co)
        c = co
        if i != len(chans) - 1:
            s //= 4
            f += 2.0 * s * 9 * co * co                         # stride-2 conv
    f += 2 * res(s, c, c) + 4 * 2.0 * s * c * c + 4.0 * s * s * c
    f += 2.0 * s * 9 * c * 8 + 2.0 * s * 8 * 8                 # conv_out 512 -> 8, quant_conv 8 -> 8
    return f * images
