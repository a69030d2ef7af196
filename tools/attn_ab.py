"""Same-process A/B of the spatial self-attention kernel against torch SDPA (cuDNN / flash backend) on the same B200:
accuracy vs fp32 SDPA on N(0,1) and on peaky (x3 logits) inputs next to torch's own bf16 SDPA error, and CUDA-event
times at the three UNet levels with a 192 MB L2 flush between launches. The share of exponentials taken off the MUFU
comes from TTVDM_ATTN_POLY (pairs of every 8, read once per process), so the driver loop at the bottom re-runs this file
per setting.

    python tools/attn_ab.py [poly[:split[:pingpong]] ...]   # default sweep 0 / 2 / 3 / 4 in subprocesses, writes gpurun_out/attn_ab.json
"""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child():
    import torch
    import torch.nn.functional as Fn
    from this_and_that_vdm_b200 import lib
    lib.init()
    dev = "cuda"

    def rel(a, b):
        return float((a.float() - b.float()).norm() / b.float().norm())

    def run(n, heads, S, gain, seed):
        C = heads * 64
        gen = torch.Generator("cpu").manual_seed(seed)
        qkv = torch.randn(n * S, 3 * C, generator=gen)
        qkv[:, :C] *= gain
        qkv = qkv.to(dev).bfloat16()
        out = torch.empty(n * S, C, dtype=torch.bfloat16, device=dev)
        lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=n, heads=heads,
                         seq=S, scale=0.125)
        q, k, v = [t.reshape(n, S, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
        ref = Fn.scaled_dot_product_attention(q.float(), k.float(), v.float()).transpose(1, 2).reshape(n * S, C)
        tb = Fn.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n * S, C)
        torch.cuda.synchronize()
        return {"own": rel(out, ref), "torch_bf16": rel(tb, ref)}

    res = {"poly_of_8": os.environ.get("TTVDM_ATTN_POLY", "default"), "split": os.environ.get("TTVDM_ATTN_SPLIT", "default"), "pingpong": os.environ.get("TTVDM_ATTN_PINGPONG", "default"), "acc": {}, "time": {}}
    for name, (n, h, S, gain) in {"S384": (3, 5, 384, 1.0), "S200_ragged": (2, 2, 200, 1.0), "S24": (4, 20, 24, 1.0),
                                  "S1536": (2, 5, 1536, 1.0), "S1536_peaky": (2, 5, 1536, 4.0),
                                  "S2304_peaky8": (2, 10, 2304, 8.0), "S9216": (1, 5, 9216, 1.0),
                                  "S9216_peaky": (1, 5, 9216, 4.0)}.items():
        res["acc"][name] = run(n, h, S, gain, 0)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)

    def time_it(fn, it=8):
        fn(); fn()
        ts = []
        for _ in range(it):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    for name, n, heads, S in [("L0", 28, 5, 9216), ("L1", 28, 10, 2304), ("L2", 28, 20, 576), ("mid", 28, 20, 144)]:
        C = heads * 64
        qkv = torch.randn(n * S, 3 * C, device=dev).bfloat16()
        out = torch.empty(n * S, C, dtype=torch.bfloat16, device=dev)
        own = time_it(lambda: lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C,
                                               n_img=n, heads=heads, seq=S, scale=0.125))
        q, k, v = [t.reshape(n, S, heads, 64).transpose(1, 2) for t in qkv.split(C, dim=1)]
        tt = time_it(lambda: Fn.scaled_dot_product_attention(q, k, v))
        fl = 4.0 * n * heads * S * S * 64
        res["time"][name] = {"own_ms": round(own, 4), "torch_ms": round(tt, 4), "own_tflops": round(fl / own / 1e9, 1),
                             "torch_tflops": round(fl / tt / 1e9, 1)}
    print("ATTN_AB " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        rows = []
        for thr in sys.argv[1:] or ["0", "2", "3", "4"]:
            env = dict(os.environ, TTVDM_ATTN_POLY=thr.split(":")[0], TTVDM_ATTN_SPLIT=(thr.split(":") + ["1", "0"])[1], TTVDM_ATTN_PINGPONG=(thr.split(":") + ["1", "0"])[2])
            try:
                r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=120)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("ATTN_AB ")]
                rows.append(json.loads(line[0][8:]) if line else {"poly_of_8": thr, "error": (r.stdout + r.stderr)[-2000:]})
            except subprocess.TimeoutExpired:
                rows.append({"poly_of_8": thr, "error": "timeout (hang?)"})
            print(json.dumps(rows[-1]), flush=True)
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / "attn_ab.json").write_text(json.dumps(rows, indent=1))
