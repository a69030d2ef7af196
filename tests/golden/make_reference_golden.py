"""Generate tests/golden/reference_pin.pt by EXECUTING THE REFERENCE'S OWN hot-path files, unmodified:

    /root/reference/svd/unet_spatio_temporal_condition.py      (UNetSpatioTemporalConditionModel.forward :363-536)
    /root/reference/svd/temporal_controlnet.py                 (ControlNetModel.forward :455-641)
    /root/reference/svd/diffusion_arch/unet_3d_blocks.py       (SpatioTemporal blocks :1870-2396)
    /root/reference/svd/diffusion_arch/transformer_temporal.py (TransformerSpatioTemporalModel :201-381)

Their only missing import is diffusers==0.25.1; tests/diffusers_shim supplies torch.nn restatements of the leaf layers
(see its README). Runs only where /root/reference exists (the authoring container); the outputs are committed and
consumed everywhere by tests/test_reference_pin.py (oracle vs golden) and tests/test_parity_gpu.py.

    python tests/golden/make_reference_golden.py [--out PATH] [--cases tiny,svd,...]

Run from the repo root. This process must never import the repo's own `svd` package: sys.path is arranged so that
`svd` resolves to /root/reference/svd.
"""
import argparse
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve()
REPO = HERE.parents[2]
REF = Path("/root/reference")
# shim first, then the reference (its `svd` package wins); the repo root is NOT on the path (refpin is loaded by file)
sys.path = [str(REPO / "tests" / "diffusers_shim"), str(REF)] + [p for p in sys.path if Path(p or ".").resolve() not in
                                                                (REPO, HERE.parent)]

import importlib.util  # noqa: E402

import torch  # noqa: E402

_spec = importlib.util.spec_from_file_location("refpin", REPO / "tests" / "refpin.py")
refpin = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(refpin)


def run_case(name: str):
    import svd
    assert Path(svd.__file__ or svd.__path__[0]).resolve().is_relative_to(REF), f"wrong svd package: {svd}"
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    kind, B, F, h, w = refpin.CASES[name]
    unet = UNetSpatioTemporalConditionModel(num_frames=F, **kind).eval()
    cn = ControlNetModel(**kind).eval()
    usd = refpin.fill_state_dict(((k, v.shape) for k, v in unet.state_dict().items()), seed=11)
    csd = refpin.fill_state_dict(((k, v.shape) for k, v in cn.state_dict().items()), seed=12)
    unet.load_state_dict(usd, strict=True)
    cn.load_state_dict(csd, strict=True)
    sample, ehs, ati, cond = refpin.make_inputs(B, F, h, w)
    cc = torch.cat([cond] * B)
    out = {}
    t0 = time.time()
    with torch.no_grad():
        out["unet"] = unet(sample, refpin.TIMESTEP, ehs, ati, return_dict=False)[0]
        down, mid = cn(sample, refpin.TIMESTEP, ehs, ati, controlnet_cond=cc, conditioning_scale=0.75,
                       return_dict=False)
        out["cn_mid"] = mid
        out["cn_down_fp"] = torch.stack([refpin.fingerprint(d) for d in down])
        out["cn_down_shapes"] = [tuple(d.shape) for d in down]
        out["vgl"] = unet(sample, torch.tensor(refpin.TIMESTEP), ehs, ati, down_block_additional_residuals=down,
                          mid_block_additional_residual=mid).sample
        if name == "tiny":
            gd, gm = cn(sample, torch.tensor([refpin.TIMESTEP]), ehs, ati, controlnet_cond=cc, conditioning_scale=1.0,
                        guess_mode=True, return_dict=False)
            out["guess_mid"] = gm
            out["guess_down_fp"] = torch.stack([refpin.fingerprint(d) for d in gd])
    out["unet_keys"] = len(usd)
    out["cn_keys"] = len(csd)
    out["unet_params"] = sum(v.numel() for v in usd.values())
    out["cn_params"] = sum(v.numel() for v in csd.values())
    print(f"{name}: reference forward x3 in {time.time() - t0:.1f} s; unet params {out['unet_params']:,}, "
          f"controlnet params {out['cn_params']:,}", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(HERE.parent / "reference_pin.pt"))
    ap.add_argument("--cases", default=",".join(refpin.CASES))
    args = ap.parse_args()
    torch.manual_seed(0)
    res = {c: run_case(c) for c in args.cases.split(",")}
    res["_meta"] = {"torch": torch.__version__, "reference": str(REF),
                    "files": ["svd/unet_spatio_temporal_condition.py", "svd/temporal_controlnet.py",
                              "svd/diffusion_arch/unet_3d_blocks.py", "svd/diffusion_arch/transformer_temporal.py"]}
    torch.save(res, args.out)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
