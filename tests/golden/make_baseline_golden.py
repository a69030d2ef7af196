"""Oracle outputs at the BASELINE.json configurations (SVD widths), committed as golden vectors for the GPU parity tests.

    python -m tests.golden.make_baseline_golden [--only vl25,vgl25,fwd72]

  vl25  : configs[1]  VL  (UNet only)        25-step Euler, 14 x 256 x 384 (latent 32 x 48), CFG pair
  vgl25 : configs[2]  VGL (UNet + GestureNet) 25-step Euler, 14 x 256 x 384, CFG pair
  fwd72 : one UNet forward, B = 1, at the headline resolution 14 x 576 x 1024 (latent 72 x 128): real tile counts,
          conv halo at W = 128, S = 9216 attention, odd CTA-pair tails
The oracle (oracle/svd_oracle.py) is pinned to the reference's own forward code by tests/test_reference_pin.py. Weights
and inputs are pure functions of seeds (tests/refpin.py), only outputs are stored (fp32). CPU cost on 8 cores:
vl25 ~ 10 min, vgl25 ~ 14 min, fwd72 ~ 3 min — which is why these are fixtures and not computed inside the tests.
"""
import argparse
import json
import time
from pathlib import Path

import torch

from oracle import svd_oracle as O
from tests import refpin
from tests.test_reference_pin import _models

HERE = Path(__file__).parent
KEEP_STEPS = (1, 5, 12, 25)  # latents after this many Euler steps


def loop_inputs(F=14, h=32, w=48, seed=21):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(1, F, 4, h, w, generator=g)
    img = torch.randn(1, 4, h, w, generator=g)
    _, ehs, ati, cond = refpin.make_inputs(2, F, h, w, seed=seed + 1)
    return noise, torch.cat([torch.zeros_like(img), img]), ehs, ati, cond


def run_loop(vgl: bool):
    kind = refpin.CASES["svd"][0]
    usd, csd = _models(kind, 14)
    cfg = dict(O.SVD_CONFIG)
    noise, img2, ehs, ati, cond = loop_inputs()
    F = 14
    sig = O.karras_sigmas(25)
    ts = O.euler_timesteps(sig)
    gd = torch.linspace(1.0, 3.0, F)[None, :, None, None, None]
    lat = noise * O.init_noise_sigma(sig)
    imgF = img2[:, None].repeat(1, F, 1, 1, 1)
    out = {}
    t0 = time.time()
    with torch.no_grad():
        for i in range(25):
            s, sn = float(sig[i]), float(sig[i + 1])
            x = torch.cat([torch.cat([lat] * 2) / (s * s + 1) ** 0.5, imgF], dim=2)
            d = m = None
            if vgl:
                d, m = O.controlnet_forward(csd, cfg, x, ts[i], ehs, ati, torch.cat([cond, cond]), 1.0)
            eps = O.unet_forward(usd, cfg, x, ts[i], ehs, ati, d, m)
            eu, ec = eps.chunk(2)
            lat = O.euler_step(eu + gd * (ec - eu), lat, s, sn)
            if i + 1 in KEEP_STEPS:
                out[f"step{i + 1}"] = lat.clone()
            print(f"  {'vgl' if vgl else 'vl'} step {i + 1}/25  {time.time() - t0:.0f} s  |lat| = {float(lat.norm()):.4g}",
                  flush=True)
    return out


def run_fwd72():
    kind = refpin.CASES["svd"][0]
    usd, _ = _models(kind, 14)
    cfg = dict(O.SVD_CONFIG)
    sample, ehs, ati, _ = refpin.make_inputs(1, 14, 72, 128, seed=31)
    t0 = time.time()
    with torch.no_grad():
        y = O.unet_forward(usd, cfg, sample, refpin.TIMESTEP, ehs, ati)
    print(f"  fwd72 {time.time() - t0:.0f} s", flush=True)
    return {"unet": y}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="fwd72,vl25,vgl25")
    a = ap.parse_args()
    for name in a.only.split(","):
        res = run_fwd72() if name == "fwd72" else run_loop(vgl=(name == "vgl25"))
        torch.save({k: v.float() for k, v in res.items()}, HERE / f"baseline_{name}.pt")
        (HERE / f"baseline_{name}.json").write_text(json.dumps(
            {"case": name, "tensors": {k: list(v.shape) for k, v in res.items()}, "weights": "tests/refpin.py seeds 11 / 12",
             "config": "SVD (320,640,1280,1280)/(5,10,20,20)", "generator": "tests/golden/make_baseline_golden.py",
             "torch": torch.__version__}, indent=1))
        print("wrote", name, flush=True)
