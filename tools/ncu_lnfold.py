"""ncu target: the level-0 GEGLU and qkv GEMMs, plain and with the LayerNorm fold, plus the C x C producer variants."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from this_and_that_vdm_b200 import lib
lib.init()
DEV, BF = "cuda", torch.bfloat16
M, C = 258048, 320
a = torch.randn(M, C, device=DEV).to(BF)
rs = torch.randn(C // 32, M, 2, device=DEV).abs() + 1
rs[:, :, 1] += 400.0
for N, geglu in ((8 * C, True), (3 * C, False)):
    w = (torch.randn(N, C, device=DEV) * C ** -0.5).to(BF)
    b, cs = torch.randn(N, device=DEV), torch.randn(N, device=DEV)
    out = torch.empty(M, N // 2 if geglu else N, dtype=BF, device=DEV)
    for rep in range(2):
        lib.gemm(a, w, out, M=M, N=N, k1=C, bias=b, geglu=geglu)
        lib.gemm(a, w, out, M=M, N=N, k1=C, bias=b, geglu=geglu, ln_rowsums=rs, ln_colsum=cs)
    torch.cuda.synchronize()
w = (torch.randn(C, C, device=DEV) * C ** -0.5).to(BF)
b = torch.randn(C, device=DEV)
r1 = torch.randn(M, C, device=DEV).to(BF)
out = torch.empty(M, C, dtype=BF, device=DEV)
st = torch.zeros(28 * C, dtype=torch.float64, device=DEV)
for rep in range(2):
    lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1)
    lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1, row_sums_out=rs)
    lib.gemm(a, w, out, M=M, N=C, k1=C, bias=b, res1=r1, gn_stats_out=st, gn_rows_per_inst=9216)
torch.cuda.synchronize()
print("done")
