"""LayerNorm kernel timing at the three UNet levels (CUDA events, 192 MB L2 flush) + the layernorm kernel-check cases."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from this_and_that_vdm_b200 import lib
from tools import gpu_kernel_check as K
lib.init()
acc = {n: f() for n, f in K.CASES if "layernorm" in n}
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, M, C in (("L0", 258048, 320), ("L1", 64512, 640), ("L2", 16128, 1280)):
    x = torch.randn(M, C, device="cuda").bfloat16(); out = torch.empty_like(x)
    g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    fn = lambda: lib.layernorm(x, out, g, b, rows=M, C=C, eps=1e-5)
    fn(); fn(); ts = []
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); res[name] = {"ms": round(ts[5], 4), "gbs": round(M * C * 4 / ts[5] / 1e6)}
print("LN_TIME", json.dumps({"acc": acc, "time": res}))
