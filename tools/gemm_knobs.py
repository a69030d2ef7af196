"""Level-0 linear shapes under experiment knobs (TTVDM_GEMM_MAX_STAGES, TTVDM_GEMM_NO_BRES, TTVDM_GEMM_L2_PREFETCH): one
subprocess per setting."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

def child():
    import torch
    from this_and_that_vdm_b200 import lib
    lib.init()
    M = 258048
    flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
    bf = torch.bfloat16
    res = {}
    for name, N, K, r, geglu in [("CxC+res", 320, 320, 1, False), ("FFout+res", 320, 1280, 1, False), ("qkv", 960, 320, 0, False), ("GEGLU", 2560, 320, 0, True), ("L1 CxC", 640, 640, 1, False), ("L1 FFout", 640, 2560, 1, False)]:
        Mx = M if not name.startswith("L1") else M // 4
        a = torch.randn(Mx, K, device="cuda").to(bf); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(bf)
        b = torch.randn(N, device="cuda"); r1 = torch.randn(Mx, N, device="cuda").to(bf) if r else None
        out = torch.empty(Mx, N // 2 if geglu else N, dtype=bf, device="cuda")
        fn = lambda: lib.gemm(a, w, out, M=Mx, N=N, k1=K, bias=b, geglu=geglu, res1=r1)
        fn(); fn(); ts = []
        for _ in range(8):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ts.sort(); res[name] = round(ts[len(ts) // 2], 4)
    print("KNOBS " + json.dumps({"stages": os.environ.get("TTVDM_GEMM_MAX_STAGES", "-"), "no_bres": os.environ.get("TTVDM_GEMM_NO_BRES", "-"), "l2_prefetch": os.environ.get("TTVDM_GEMM_L2_PREFETCH", "-"), "ms": res}), flush=True)

if __name__ == "__main__":
    if "--child" in sys.argv: child()
    else:
        sweeps = [("0", None, "0"), ("0", None, "2"), ("0", None, "4"), ("0", None, "8"), ("0", None, "16"), ("0", "1", "8"), ("3", None, "8")]
        if "--stages" in sys.argv:
            sweeps = [("0", None, "0"), ("2", None, "0"), ("3", None, "0"), ("0", "1", "0"), ("2", "1", "0"), ("3", "1", "0")]
        for st, nb, pf in sweeps:
            env = dict(os.environ, TTVDM_GEMM_MAX_STAGES=st, TTVDM_GEMM_L2_PREFETCH=pf)
            if nb: env["TTVDM_GEMM_NO_BRES"] = nb
            r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True, timeout=120)
            print([l for l in r.stdout.splitlines() if l.startswith("KNOBS")] or r.stderr[-500:], flush=True)
