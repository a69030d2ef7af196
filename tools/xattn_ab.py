"""Cross-attention A/B on one B200: tcgen05 / TMA path (flash kernel, cross mode) vs the warp-level mma.sync kernel
(TTVDM_XATTN_LEGACY=1), spatial and temporal, at the three UNet levels of a VGL step at 14x576x1024; accuracy of both
against an fp32 torch restatement on a small case. One subprocess per setting (the switch is read once)."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def child():
    import torch
    from this_and_that_vdm_b200 import lib
    from tools import gpu_kernel_check as K
    lib.init()
    res = {"legacy": os.environ.get("TTVDM_XATTN_LEGACY", "0"), "acc": {}, "time": {}}
    for name, fn in K.CASES:
        if name.startswith("attn_cross"):
            res["acc"][name] = fn()
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")

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

    for lvl, heads, S in (("L0", 5, 9216), ("L1", 10, 2304), ("L2", 20, 576)):
        C, M = heads * 64, 2 * 14 * S
        x = torch.randn(M, C, device="cuda").bfloat16()
        kc = torch.randn(2 * 78, C, device="cuda").bfloat16()
        vc = torch.randn(2 * 78, C, device="cuda").bfloat16()
        out = torch.empty_like(x)
        for temporal in (False, True):
            ms = time_it(lambda: lib.attn_cross(x, kc, vc, out, ldq=C, ldo=C, rows=M, heads=heads, L=78, F=14, S=S, n_ctx=2,
                                                temporal=temporal, batch_offset=0, scale=0.125))
            res["time"][f"{lvl}_{'temporal' if temporal else 'spatial'}"] = {"ms": round(ms, 4), "gbs": round(M * C * 4 / ms / 1e6)}
    print("XATTN_AB " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        rows = []
        for legacy in ("0", "1"):
            env = dict(os.environ, TTVDM_XATTN_LEGACY=legacy)
            try:
                r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=150)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("XATTN_AB ")]
                rows.append(json.loads(line[0][9:]) if line else {"legacy": legacy, "error": (r.stdout + r.stderr)[-1500:]})
            except subprocess.TimeoutExpired:
                rows.append({"legacy": legacy, "error": "timeout (hang?)"})
            print(json.dumps(rows[-1]), flush=True)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / "xattn_ab.json").write_text(json.dumps(rows, indent=1))
