"""Three launches of the S = 9216 spatial self-attention (28 images x 5 heads): the target of `ncu -k regex:attn_flash`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from this_and_that_vdm_b200 import lib
lib.init()
n, heads, S = 28, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 9216
C = heads * 64
qkv = torch.randn(n * S, 3 * C, generator=torch.Generator().manual_seed(0)).to("cuda", torch.bfloat16)
out = torch.empty(n * S, C, dtype=torch.bfloat16, device="cuda")
for _ in range(3):
    lib.attn_spatial(qkv, qkv[:, C:], qkv[:, 2 * C:], out, ldq=3 * C, ldk=3 * C, ldv=3 * C, ldo=C, n_img=n, heads=heads, seq=S, scale=0.125)
torch.cuda.synchronize()
print("ok")
