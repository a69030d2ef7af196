"""Whole-network parity report on a B200: sm_100a engine (bf16) vs the CPU fp32 oracle, with torch-eager bf16 on the
same GPU (the oracle functions run on CUDA bf16 tensors) as the yard-stick. Prints rel-L2 errors; never asserts."""
from __future__ import annotations

import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import svd_oracle as O  # noqa: E402
from tests.common import SVD, TINY, build_models, make_inputs, oracle_cfg, rel_l2, state  # noqa: E402


def run(kind, name, h, w, steps=2):
    cfg = oracle_cfg(kind)
    t0 = time.time()
    unet, cn = build_models(kind)
    print(f"[{name}] models built in {time.time() - t0:.1f}s", flush=True)
    B, F = 2, 14
    sample, ehs, ati, cond = make_inputs(B, F, h, w)
    t = torch.tensor(1.63777)
    usd, csd = state(unet), state(cn)
    with torch.no_grad():
        t0 = time.time()
        cc = torch.cat([cond, cond])
        d_ref, m_ref = O.controlnet_forward(csd, cfg, sample, t, ehs, ati, cc, 1.0)
        y_ref, inter = O.unet_forward(usd, cfg, sample, t, ehs, ati, d_ref, m_ref, return_intermediates=True)
        y_ref0 = O.unet_forward(usd, cfg, sample, t, ehs, ati)
        print(f"[{name}] oracle fp32 CPU: {time.time() - t0:.1f}s  ({torch.get_num_threads()} threads)", flush=True)
        # torch-eager bf16 on the GPU (same functions, bf16 tensors)
        dev = "cuda"
        usd_b = {k: v.to(dev, torch.bfloat16) for k, v in usd.items()}
        csd_b = {k: v.to(dev, torch.bfloat16) for k, v in csd.items()}
        sb, eb, ab, cb = sample.to(dev, torch.bfloat16), ehs.to(dev, torch.bfloat16), ati.to(dev, torch.bfloat16), cc.to(dev, torch.bfloat16)
        d_e, m_e = O.controlnet_forward(csd_b, cfg, sb, t.to(dev), eb, ab, cb, 1.0)
        y_e = O.unet_forward(usd_b, cfg, sb, t.to(dev), eb, ab, d_e, m_e)
        y_e0 = O.unet_forward(usd_b, cfg, sb, t.to(dev), eb, ab)
        del usd_b, csd_b
        print(f"[{name}] eager-bf16  unet: {rel_l2(y_e0, y_ref0):.3e}  cn mid: {rel_l2(m_e, m_ref):.3e}  unet+cn: {rel_l2(y_e, y_ref):.3e}")
        unet.to(dev)
        cn.to(dev)
        sg, eg, ag, cg = sample.to(dev), ehs.to(dev), ati.to(dev), cc.to(dev)
        y0 = unet(sg, t.to(dev), eg, ag, return_dict=False)[0]
        print(f"[{name}] engine      unet: {rel_l2(y0, y_ref0):.3e}", flush=True)
        d, m = cn(sg, t.to(dev), eg, ag, controlnet_cond=cg, return_dict=False)
        print(f"[{name}] engine      cn mid: {rel_l2(m, m_ref):.3e}  cn down: " + " ".join(f"{rel_l2(a, b):.1e}" for a, b in zip(d, d_ref)))
        y = unet(sg, t.to(dev), eg, ag, down_block_additional_residuals=d, mid_block_additional_residual=m, return_dict=False)[0]
        print(f"[{name}] engine      unet+cn: {rel_l2(y, y_ref):.3e}", flush=True)
        # fused loop (few steps) vs oracle loop
        if steps:
            from this_and_that_vdm_b200.sampler import FusedDenoiser
            g = torch.Generator().manual_seed(7)
            sig = O.karras_sigmas(25)
            lat0 = torch.randn(1, F, 4, h, w, generator=g) * O.init_noise_sigma(sig)
            img = torch.randn(1, 4, h, w, generator=g)
            img2 = torch.cat([torch.zeros_like(img), img])
            ref = O.denoise_loop(usd, cfg, lat0, img2[:, None].repeat(1, F, 1, 1, 1), ehs, ati, 25, 1.0, 3.0, csd, cfg, cond, 1.0, max_steps=steps)
            den = FusedDenoiser(unet._get_engine(), cn._get_engine())
            den.prepare(eg, img2.to(dev), ag, sig, O.euler_timesteps(sig), torch.linspace(1, 3, F), num_frames=F, height=h, width=w, controlnet_cond=cond.to(dev))
            st = lat0[0].to(dev).clone()
            for i in range(steps):
                den.step(i, st)
            torch.cuda.synchronize()
            print(f"[{name}] fused VGL {steps} steps: {rel_l2(st, ref[0]):.3e}  (latent update only: {rel_l2(st - lat0[0].to(dev), ref[0] - lat0[0]):.3e})", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    if which in ("tiny", "all"):
        run(TINY, "tiny 16x24", 16, 24)
    if which in ("svd", "all"):
        run(SVD, "svd 32x48", 32, 48, steps=0)
