"""Generates tests/golden/tiny_vae.{pt,json}: oracle outputs (oracle/vae_oracle.py) for the seeded tiny VAE case.

Run:  python -m tests.golden.make_vae_golden
Weights and inputs are regenerated from seeds on the GPU box; only the expected outputs are stored."""
import json
from pathlib import Path

import torch

from oracle import vae_oracle as VO
from tests.common import TINY_VAE, build_vae, state, vae_inputs

HERE = Path(__file__).parent
N_VIDEOS, FR, LH, LW = 2, 4, 8, 12


def compute():
    vae = build_vae(TINY_VAE)
    z, x = vae_inputs(N_VIDEOS * FR, LH, LW)
    sd = state(vae)
    with torch.no_grad():
        return {"decode": VO.decode(sd, z, FR), "encode_mean": VO.encode(sd, x)}


if __name__ == "__main__":
    out = compute()
    torch.save({k: v.to(torch.float32) for k, v in out.items()}, HERE / "tiny_vae.pt")
    (HERE / "tiny_vae.json").write_text(json.dumps(
        {"tensors": list(out.keys()), "config": "TINY_VAE (64,128,128,128), layers_per_block 1",
         "latent": [N_VIDEOS * FR, 4, LH, LW], "num_frames": FR, "seeds": {"weights": 4321, "inputs": 5},
         "generator": "tests/golden/make_vae_golden.py (oracle/vae_oracle.py)"}, indent=1))
    print({k: tuple(v.shape) for k, v in out.items()})
