"""The four level-0 linear shapes of a VGL step at 14x576x1024, two launches each with an L2 flush in between: the target
of `ncu --set full -k regex:gemm_kernel` when a shape has to be understood in isolation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from this_and_that_vdm_b200 import lib
lib.init()
M = 258048
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
bf = torch.bfloat16
for (N, K, res, geglu) in [(320, 320, 1, False), (320, 1280, 1, False), (960, 320, 0, False), (2560, 320, 0, True)]:
    a = torch.randn(M, K, device="cuda").to(bf)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(bf)
    b = torch.randn(N, device="cuda")
    r1 = torch.randn(M, N, device="cuda").to(bf) if res else None
    out = torch.empty(M, N // 2 if geglu else N, dtype=bf, device="cuda")
    for _ in range(2):
        flush.zero_()
        lib.gemm(a, w, out, M=M, N=N, k1=K, bias=b, geglu=geglu, res1=r1)
    torch.cuda.synchronize()
print("ok")
