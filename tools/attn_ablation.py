import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from this_and_that_vdm_b200 import lib
lib.init()
n, heads, S = 28, 5, 9216
C = heads * 64
g = torch.Generator().manual_seed(0)
qkv = torch.randn(n * S, 3 * C, generator=g).to("cuda", torch.bfloat16)
out = torch.empty(n * S, C, dtype=torch.bfloat16, device="cuda")
def run():
    lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=n, heads=heads, seq=S, scale=0.125)
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
print("dbg", os.environ.get("TTVDM_ATTN_DBG", "0"), "ms", e0.elapsed_time(e1) / 5)
