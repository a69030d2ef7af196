"""Own kernel vs the torch library op (cuBLAS / cuDNN / SDPA) on the SAME B200, same process, for every hot shape of one
VGL Euler step at 14x576x1024 (B = 2): the per-shape table VERDICT r1 asked for (profiles/r02_shape_table.json).

Own kernels are timed WITH their production epilogue (bias, residuals, GEGLU ...); the torch op is the bare contraction
(the library gets the easier job). 192 MB L2 flush between timed launches. Never raises per row.

    python tools/shape_table.py [--quick]
"""
from __future__ import annotations

import json
import sys
import traceback
from pathlib import Path

import torch
import torch.nn.functional as Fn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from this_and_that_vdm_b200 import lib  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def rnd(*shape, scale=1.0, dtype=BF):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


_flush = None


def time_it(fn, iters=10, warm=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def row_linear(M, N, K, *, geglu=False, res=0, per_step=0, note=""):
    a, w, b = rnd(M, K), rnd(N, K, scale=K ** -0.5), torch.randn(N, device=DEV)
    r1 = rnd(M, N) if res >= 1 else None
    r2 = rnd(M, N) if res >= 2 else None
    out = torch.empty(M, N // 2 if geglu else N, dtype=BF, device=DEV)
    own = time_it(lambda: lib.gemm(a, w, out, M=M, N=N, k1=K, bias=b, geglu=geglu, res1=r1, res2=r2))
    ref = time_it(lambda: Fn.linear(a, w))
    fl = 2.0 * M * N * K
    byt = 2.0 * (M * K + N * K + M * (N // 2 if geglu else N) + res * M * N)
    return {"op": f"linear{'+geglu' if geglu else ''}{'+res' * res}", "M": M, "N": N, "K": K, "per_step": per_step,
            "own_ms": round(own, 4), "torch_ms": round(ref, 4), "own_tflops": round(fl / own / 1e9, 1),
            "torch_tflops": round(fl / ref / 1e9, 1), "own_gbs": round(byt / own / 1e6, 0), "alg_mb": round(byt / 1e6, 1),
            "note": note}


def row_fusion_variants(M, C, S, F=14):
    """Cost of the epilogue fusions in isolation, L-level shapes: producer C x C (+res) plain / + LayerNorm row sums /
    + GroupNorm pair sums; consumer qkv and GEGLU with the LayerNorm fold; 3x3 conv and temporal conv with GroupNorm sums."""
    rows = []
    a, w, b, r1 = rnd(M, C), rnd(C, C, scale=C ** -0.5), torch.randn(C, device=DEV), rnd(M, C)
    out = torch.empty(M, C, dtype=BF, device=DEV)
    rs = torch.empty(C // 32, M, 2, dtype=torch.float32, device=DEV)
    st = torch.zeros((M // S) * C, dtype=torch.float64, device=DEV)
    base = time_it(lambda: lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1))
    t_rs = time_it(lambda: lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1, row_sums_out=rs))
    t_gn = time_it(lambda: lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1, gn_stats_out=st, gn_rows_per_inst=S))
    rows.append({"op": "CxC+res: plain / +row_sums / +gn_stats", "M": M, "N": C, "K": C, "plain_ms": round(base, 4),
                 "row_sums_ms": round(t_rs, 4), "gn_stats_ms": round(t_gn, 4)})
    for N, geglu in ((3 * C, False), (8 * C, True)):
        wq, bq, cs = rnd(N, C, scale=C ** -0.5), torch.randn(N, device=DEV), torch.randn(N, device=DEV)
        o2 = torch.empty(M, N // 2 if geglu else N, dtype=BF, device=DEV)
        base = time_it(lambda: lib.gemm(a, wq, o2, M=M, N=N, k1=C, bias=bq, geglu=geglu))
        t_ln = time_it(lambda: lib.gemm(a, wq, o2, M=M, N=N, k1=C, bias=bq, geglu=geglu, ln_rowsums=rs, ln_colsum=cs))
        rows.append({"op": f"{'GEGLU' if geglu else 'qkv'}: plain / +LayerNorm fold", "M": M, "N": N, "K": C,
                     "plain_ms": round(base, 4), "ln_fold_ms": round(t_ln, 4)})
        del o2, wq
    n, H = M // S, int(S ** 0.5 * (9 / 16) ** 0.5 + 0.5)
    Wd = S // H
    if H * Wd == S:
        x, wk = rnd(n, H, Wd, C), rnd(C, 9 * C, scale=(9 * C) ** -0.5)
        base = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_CONV3X3, n_img=n, H=H, W=Wd, bias=b))
        t4 = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_CONV3X3, n_img=n, H=H, W=Wd, bias=b,
                                      gn_stats_out=st, gn_rows_per_inst=S))
        t5 = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_CONV3X3, n_img=n, H=H, W=Wd, bias=b,
                                      gn_stats_out=st, gn_rows_per_inst=F * S))
        rows.append({"op": "conv3x3: plain / +gn_stats 4-D / 5-D", "M": M, "N": C, "K": 9 * C, "plain_ms": round(base, 4),
                     "gn4_ms": round(t4, 4), "gn5_ms": round(t5, 4)})
        del wk
    B = n // F
    x, wk = rnd(B, F, S, C), rnd(C, 3 * C, scale=(3 * C) ** -0.5)
    base = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, res1=r1))
    t4 = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, res1=r1,
                                  gn_stats_out=st, gn_rows_per_inst=S))
    t5 = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, res1=r1,
                                  gn_stats_out=st, gn_rows_per_inst=F * S))
    rows.append({"op": "tconv3+res: plain / +gn_stats 4-D / 5-D", "M": M, "N": C, "K": 3 * C, "plain_ms": round(base, 4),
                 "gn4_ms": round(t4, 4), "gn5_ms": round(t5, 4)})
    gm, bt = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    ws = torch.empty((M // S) * 64, dtype=torch.float64, device=DEV)
    y = torch.empty_like(out)
    t_full = time_it(lambda: lib.groupnorm(out, y, ws, gm, bt, c1=C, rows=M, rows_per_inst=S, eps=1e-6, silu=True))
    t_app = time_it(lambda: lib.groupnorm(out, y, None, gm, bt, c1=C, rows=M, rows_per_inst=S, eps=1e-6, silu=True, pstats1=st))
    rows.append({"op": "groupnorm: stats+apply / apply only (producer sums)", "M": M, "N": C, "K": 0,
                 "full_ms": round(t_full, 4), "apply_only_ms": round(t_app, 4)})
    return rows


def row_conv(n, H, W, Ci, Co, *, per_step=0):
    x, wk, b = rnd(n, H, W, Ci), rnd(Co, 9 * Ci, scale=(9 * Ci) ** -0.5), torch.randn(Co, device=DEV)
    tv = torch.randn(2, Co, device=DEV)
    M = n * H * W
    out = torch.empty(M, Co, dtype=BF, device=DEV)
    own = time_it(lambda: lib.gemm(x, wk, out, M=M, N=Co, k1=Ci, mode=lib.A_CONV3X3, n_img=n, H=H, W=W, bias=b, rowvec=tv,
                                   rows_per_vec=M // 2))
    xc = x.permute(0, 3, 1, 2)
    wc = wk.reshape(Co, 3, 3, Ci).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    ref = time_it(lambda: Fn.conv2d(xc, wc, None, padding=1))
    fl = 2.0 * M * Co * 9 * Ci
    byt = 2.0 * (M * Ci + 9 * Ci * Co + M * Co)
    return {"op": "conv3x3", "M": M, "N": Co, "K": 9 * Ci, "per_step": per_step, "own_ms": round(own, 4),
            "torch_ms": round(ref, 4), "own_tflops": round(fl / own / 1e9, 1), "torch_tflops": round(fl / ref / 1e9, 1),
            "own_gbs": round(byt / own / 1e6, 0), "alg_mb": round(byt / 1e6, 1), "note": f"{n}x{H}x{W}"}


def row_tconv(B, F, S, C, *, per_step=0):
    x, wk, b = rnd(B, F, S, C), rnd(C, 3 * C, scale=(3 * C) ** -0.5), torch.randn(C, device=DEV)
    res = rnd(B * F * S, C)
    M = B * F * S
    out = torch.empty(M, C, dtype=BF, device=DEV)
    own = time_it(lambda: lib.gemm(x, wk, out, M=M, N=C, k1=C, mode=lib.A_TCONV3, n_img=B, H=F, W=S, bias=b, res1=res))
    x5 = x.permute(0, 3, 1, 2).reshape(B, C, F, S, 1).contiguous(memory_format=torch.channels_last_3d)
    w5 = wk.reshape(C, 3, C).permute(0, 2, 1).reshape(C, C, 3, 1, 1).contiguous(memory_format=torch.channels_last_3d)
    ref = time_it(lambda: Fn.conv3d(x5, w5, None, padding=(1, 0, 0)))
    fl = 2.0 * M * C * 3 * C
    byt = 2.0 * (3 * M * C + 3 * C * C)
    return {"op": "tconv3+res", "M": M, "N": C, "K": 3 * C, "per_step": per_step, "own_ms": round(own, 4),
            "torch_ms": round(ref, 4), "own_tflops": round(fl / own / 1e9, 1), "torch_tflops": round(fl / ref / 1e9, 1),
            "own_gbs": round(byt / own / 1e6, 0), "alg_mb": round(byt / 1e6, 1), "note": f"B{B} F{F} S{S}"}


def row_attn(n, heads, S, *, per_step=0):
    C = heads * 64
    qkv = rnd(n * S, 3 * C)
    out = torch.empty(n * S, C, dtype=BF, device=DEV)
    own = time_it(lambda: lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C,
                                           n_img=n, heads=heads, seq=S, scale=0.125), iters=5)
    q, k, v = [t.reshape(n, S, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
    ref = time_it(lambda: Fn.scaled_dot_product_attention(q, k, v), iters=5)
    fl = 4.0 * n * heads * S * S * 64
    return {"op": "self-attention d64", "M": n * S, "N": heads, "K": S, "per_step": per_step, "own_ms": round(own, 4),
            "torch_ms": round(ref, 4), "own_tflops": round(fl / own / 1e9, 1), "torch_tflops": round(fl / ref / 1e9, 1),
            "note": f"{n} images x {heads} heads x S={S}; torch = F.scaled_dot_product_attention (cuDNN / flash backend)"}


def row_norms(M, C, S):
    x, gm, bt = rnd(M, C), torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    out = torch.empty_like(x)
    stats = torch.empty((M // S) * 64, dtype=torch.float64, device=DEV)
    rows = []
    own = time_it(lambda: lib.groupnorm(x, out, stats, gm, bt, c1=C, rows=M, rows_per_inst=S, eps=1e-6, silu=True))
    xc = x.view(M // S, S, C).permute(0, 2, 1).unsqueeze(-1)  # [n, C, S, 1] channels-last storage
    ref = time_it(lambda: Fn.silu(Fn.group_norm(xc, 32, gm.to(BF), bt.to(BF), 1e-6)))
    rows.append({"op": "groupnorm+silu", "M": M, "N": C, "K": 0, "own_ms": round(own, 4), "torch_ms": round(ref, 4),
                 "own_gbs": round(M * C * 6 / own / 1e6, 0), "alg_mb": round(M * C * 6 / 1e6, 1),
                 "note": "own = stats + apply passes (6 B/elem); torch = group_norm + silu (2 kernels)"})
    own = time_it(lambda: lib.layernorm(x, out, gm, bt, rows=M, C=C))
    ref = time_it(lambda: Fn.layer_norm(x, (C,), gm.to(BF), bt.to(BF), 1e-5))
    rows.append({"op": "layernorm", "M": M, "N": C, "K": 0, "own_ms": round(own, 4), "torch_ms": round(ref, 4),
                 "own_gbs": round(M * C * 4 / own / 1e6, 0), "alg_mb": round(M * C * 4 / 1e6, 1), "note": ""})
    return rows


def main():
    lib.init()
    quick = "--quick" in sys.argv
    S0, S1, S2, S3 = 9216, 2304, 576, 144
    n = 28
    jobs = [
        lambda: row_linear(n * S0, 2560, 320, geglu=True, per_step=21, note="L0 GEGLU proj"),
        lambda: row_linear(n * S0, 320, 320, res=1, per_step=59, note="L0 CxC (out-proj / q / proj_in / proj_out)"),
        lambda: row_linear(n * S0, 320, 1280, res=1, per_step=21, note="L0 FF out"),
        lambda: row_linear(n * S0, 960, 320, per_step=14, note="L0 fused qkv"),
        lambda: row_conv(n, 72, 128, 320, 320, per_step=11),
        lambda: row_tconv(2, 14, S0, 320, per_step=14),
        lambda: row_linear(n * S1, 5120, 640, geglu=True, per_step=21, note="L1 GEGLU proj"),
        lambda: row_linear(n * S1, 640, 2560, res=1, per_step=21, note="L1 FF out"),
        lambda: row_linear(n * S1, 640, 640, res=1, per_step=58, note="L1 CxC"),
        lambda: row_linear(n * S1, 1920, 640, per_step=14, note="L1 fused qkv"),
        lambda: row_conv(n, 36, 64, 640, 640, per_step=9),
        lambda: row_tconv(2, 14, S1, 640, per_step=14),
        lambda: row_linear(n * S2, 10240, 1280, geglu=True, per_step=21, note="L2 GEGLU proj"),
        lambda: row_linear(n * S2, 1280, 5120, res=1, per_step=21, note="L2 FF out"),
        lambda: row_linear(n * S2, 1280, 1280, res=1, per_step=58, note="L2 CxC"),
        lambda: row_conv(n, 18, 32, 1280, 1280, per_step=10),
        lambda: row_conv(n, 9, 16, 1280, 1280, per_step=12),
        lambda: row_attn(n, 5, S0, per_step=7),
        lambda: row_attn(n, 10, S1, per_step=7),
        lambda: row_attn(n, 20, S2, per_step=7),
        lambda: row_attn(n, 20, S3, per_step=2),
    ]
    if quick:
        jobs = jobs[:3] + jobs[17:18]
    if "--fusions" in sys.argv:
        jobs = []
    if "--attn" in sys.argv:
        jobs = jobs[17:]
    if "--gemm-only" in sys.argv:
        jobs = jobs[:17]
    table = []
    for j in jobs:
        try:
            r = j()
            table.append(r)
            print("ROW", json.dumps(r), flush=True)
        except Exception as e:  # noqa: BLE001
            traceback.print_exc()
            print("ROW", json.dumps({"error": str(e)[:200]}), flush=True)
        torch.cuda.empty_cache()
    if "--fusions" in sys.argv:
        table = []
        for M, C, S in ((n * S0, 320, S0), (n * S1, 640, S1), (n * S2, 1280, S2)):
            try:
                for r in row_fusion_variants(M, C, S):
                    table.append(r)
                    print("ROW", json.dumps(r), flush=True)
            except Exception:  # noqa: BLE001
                traceback.print_exc()
            torch.cuda.empty_cache()
        out = Path(__file__).resolve().parents[1] / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / "shape_table_fusions.json").write_text(json.dumps({"device": torch.cuda.get_device_name(0), "rows": table}, indent=1))
        return
    if "--attn" in sys.argv:
        return
    try:
        for r in row_norms(n * S0, 320, S0) + row_norms(n * S1, 640, S1):
            table.append(r)
            print("ROW", json.dumps(r), flush=True)
    except Exception:  # noqa: BLE001
        traceback.print_exc()
    out = Path(__file__).resolve().parents[1] / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "shape_table.json").write_text(json.dumps({"device": torch.cuda.get_device_name(0), "rows": table}, indent=1))


if __name__ == "__main__":
    main()
