"""Debug aid (GPU): run-to-run determinism and GPU-vs-emulation agreement of the engine, block by block.
    python tools/gpu_determinism.py [svd]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests import fake_lib  # noqa: E402
from tests.common import SVD, TINY, build_models, make_inputs, rel_l2  # noqa: E402
from this_and_that_vdm_b200.engine import DenoiserEngine  # noqa: E402

kind = SVD if (len(sys.argv) > 1 and sys.argv[1] == "svd") else TINY
B, F, h, w = 2, 14, 16, 24
unet, _ = build_models(kind, controlnet=False)
with fake_lib.installed():
    cpu_eng = DenoiserEngine(unet, "unet")
sample, ehs, ati, cond = make_inputs(B, F, h, w)
T0 = torch.tensor(1.63777)
unet.to("cuda")
eng = unet._get_engine()


def fwd():
    with torch.no_grad():
        return eng.unet_forward(sample.cuda(), T0.cuda(), ehs.cuda(), ati.cuda()).float().cpu()


for fuse in (True, False):
    eng.fuse_norm_stats = fuse
    a, b = fwd(), fwd()
    print(f"whole forward, fuse_norm_stats={fuse}: run-to-run rel_l2 = {rel_l2(a, b):.3e}")
eng.fuse_norm_stats = True
with torch.no_grad(), fake_lib.installed():
    ref = cpu_eng.unet_forward(sample, T0, ehs, ati)
print(f"GPU vs emulation (whole forward): {rel_l2(fwd(), ref):.3e}")

# ---- block level: first resblock and first transformer of every level
with torch.no_grad():
    eng._ensure_pos_emb(F)
    with fake_lib.installed():
        cpu_eng._ensure_pos_emb(F)
        temb_c = cpu_eng.time_embeddings(T0.expand(B).contiguous(), ati)
        kv_c = cpu_eng.context_kv(ehs)
    temb_g = eng.time_embeddings(T0.expand(B).contiguous().cuda(), ati.cuda())
    kv_g = eng.context_kv(ehs.cuda())
    ti = 0
    H, W = h, w
    for lvl, (bg, bc) in enumerate(zip(eng.down, cpu_eng.down)):
        C = bg["res"][0].cin
        g = torch.Generator().manual_seed(lvl)
        x = torch.randn(B * F * H * W, C, generator=g).to(torch.bfloat16)
        for name in ("res", "tf"):
            if not bg[name]:
                continue
            outs = []
            for rep in range(2):
                eng.begin_step()
                xg = x.cuda()
                if name == "res":
                    o = eng._resblock(bg["res"][0], xg, None, B=B, F=F, H=H, W=W, temb=temb_g)
                else:
                    xg.gn_stats = None
                    o = eng._transformer(bg["tf"][0], xg, kv_g[ti], B=B, F=F, H=H, W=W, n_ctx=B, batch_offset=0)
                torch.cuda.synchronize()
                outs.append(o.float().cpu())
            with fake_lib.installed():
                cpu_eng.begin_step()
                if name == "res":
                    oc = cpu_eng._resblock(bc["res"][0], x.clone(), None, B=B, F=F, H=H, W=W, temb=temb_c)
                else:
                    oc = cpu_eng._transformer(bc["tf"][0], x.clone(), kv_c[ti], B=B, F=F, H=H, W=W, n_ctx=B, batch_offset=0)
            print(f"level {lvl} {name}: C={C} {H}x{W}  run-to-run {rel_l2(outs[0], outs[1]):.3e}   vs emulation "
                  f"{rel_l2(outs[0], oc.float()):.3e}")
        if bg["tf"]:
            ti += len(bg["tf"])
        if bg["down_w"] is not None:
            H, W = H // 2, W // 2
