"""Generates tests/golden/tiny_vgl.{pt,json}: oracle outputs for the seeded tiny-config VGL case.

Run:  python -m tests.golden.make_golden
The fixture pins (a) the oracle against itself over time and (b) the CUDA path on the GPU box, where
/root/reference and this container's RNG state are not available — inputs and weights are regenerated from
seeds (torch CPU RNG is deterministic), only the expected outputs are stored (fp32, < 1 MB)."""
import json
from pathlib import Path

import torch

from oracle import svd_oracle as O
from tests.common import TINY, build_models, make_inputs, oracle_cfg, state

HERE = Path(__file__).parent
B, FR, H, W = 2, 14, 16, 24


def compute():
    cfg = oracle_cfg(TINY)
    unet, cn = build_models(TINY)
    sample, ehs, ati, cond = make_inputs(B, FR, H, W)
    usd, csd = state(unet), state(cn)
    t = torch.tensor(1.63777)
    with torch.no_grad():
        d, m = O.controlnet_forward(csd, cfg, sample, t, ehs, ati, torch.cat([cond, cond]), 1.0)
        y_vl = O.unet_forward(usd, cfg, sample, t, ehs, ati)
        y_vgl = O.unet_forward(usd, cfg, sample, t, ehs, ati, d, m)
    return {"unet_vl": y_vl, "unet_vgl": y_vgl, "cn_mid": m, "cn_down11": d[11]}


if __name__ == "__main__":
    out = compute()
    torch.save({k: v.to(torch.float32) for k, v in out.items()}, HERE / "tiny_vgl.pt")
    (HERE / "tiny_vgl.json").write_text(json.dumps(
        {"tensors": list(out.keys()), "config": "TINY (64,128,256,256)/(1,2,4,4)", "shape": [B, FR, H, W],
         "seeds": {"weights": 1234, "inputs": 0}, "generator": "tests/golden/make_golden.py (oracle/svd_oracle.py)"},
        indent=1))
    print({k: tuple(v.shape) for k, v in out.items()})
