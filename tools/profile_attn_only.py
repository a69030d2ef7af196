import sys, torch
sys.path.insert(0, "/root/repo")
from this_and_that_vdm_b200 import lib
lib.init()
n, heads, S = 28, 5, 9216
C = heads * 64
g = torch.Generator().manual_seed(0)
qkv = torch.randn(n * S, 3 * C, generator=g).to("cuda", torch.bfloat16)
out = torch.empty(n * S, C, dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=n, heads=heads, seq=S, scale=0.125)
torch.cuda.synchronize()
