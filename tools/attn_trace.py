"""Event trace of one CTA of the spatial self-attention kernel (debug aid, needs a trace build:
TTVDM_EXTRA_NVCC_FLAGS=-DTTVDM_ATTN_TRACE python this_and_that_vdm_b200/build.py --force).
Prints, for a window of KV tiles, what the MMA scheduler thread issued when, and where the softmax warps waited."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CAP = 4096
tr = torch.zeros(3 * CAP, dtype=torch.int64, device="cuda")
os.environ["TTVDM_ATTN_TRACE"] = str(tr.data_ptr())
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
M64 = (1 << 64) - 1
t = [[int(x) & M64 for x in row] for row in tr.cpu().view(3, CAP).tolist()]
base = min(x >> 8 for row in t for x in row if x)
names = {1: "pre_sfull", 2: "got_sfull", 3: "s_in_regs", 4: "max_done", 5: "exp_done", 6: "p_pub", 7: "token",
         0x10: "QK0{", 0x11: "QK1{", 0x18: "}QK0", 0x19: "}QK1", 0x20: "PV0{", 0x21: "PV1{", 0x28: "}PV0", 0x29: "}PV1"}
for q_ in range(64): names[0x40 + q_] = "idle(%02x)" % q_
lo, hi = int(sys.argv[1]) if len(sys.argv) > 1 else 30000, int(sys.argv[2]) if len(sys.argv) > 2 else 42000
for cls, label in enumerate(["MMA", "T0", "T1"]):
    ev = [(x >> 8, x & 255) for x in t[cls] if x]
    print(label, " ".join("%s@%d" % (names.get(tag, hex(tag)), c - base) for c, tag in ev if lo <= c - base <= hi))
ends = [(x >> 8) - base for x in t[1] if x and (x & 255) == 6]
if len(ends) > 60:
    print("cycles per KV tile (tile 0 softmax, steady state): %.0f" % ((ends[60] - ends[20]) / 40.0))
# ---- per-phase averages of the softmax thread of each Q tile over the steady state (KV tiles 20..60 of the trace)
for cls, label in ((1, "T0"), (2, "T1")):
    ev = [((x >> 8) - base, x & 255) for x in t[cls] if x]
    seg = {}
    last = None
    n6 = 0
    for c, tag in ev:
        if tag == 6:
            n6 += 1
        if last is not None and 20 <= n6 <= 60:
            seg.setdefault((last[1], tag), []).append(c - last[0])
        last = (c, tag)
    print(label, "phase averages:", "  ".join("%s->%s %.0f" % (names.get(a, hex(a)), names.get(b, hex(b)), sum(v) / len(v))
                                             for (a, b), v in sorted(seg.items())))
ev = [((x >> 8) - base, x & 255) for x in t[0] if x]
seg = {}
last = None
for c, tag in ev[200:1000]:
    if last is not None:
        seg.setdefault((last[1], tag), []).append(c - last[0])
    last = (c, tag)
print("MMA issuer phase averages:", "  ".join("%s->%s %.0f(n=%d)" % (names.get(a, hex(a)), names.get(b, hex(b)), sum(v) / len(v), len(v))
                                           for (a, b), v in sorted(seg.items())))
# ---- relative phase of the two Q tiles: start of exp phase (tag 4) of T0 vs T1, modulo the period
e0 = [((x >> 8) - base) for x in t[1] if x and (x & 255) == 4][20:60]
e1 = [((x >> 8) - base) for x in t[2] if x and (x & 255) == 4][20:60]
if e0 and e1:
    print("exp-phase start offsets T1 - T0 (cycles):", " ".join(str(b - a) for a, b in list(zip(e0, e1))[:24]))
