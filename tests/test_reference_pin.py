"""PIN of the denoiser oracle against the reference's OWN hot-path code.

Two layers:
  * golden (runs everywhere): tests/golden/reference_pin.pt was produced by tests/golden/make_reference_golden.py, which
    imports /root/reference/svd/{unet_spatio_temporal_condition,temporal_controlnet}.py and
    svd/diffusion_arch/{unet_3d_blocks,transformer_temporal}.py UNMODIFIED (diffusers leaf layers from tests/diffusers_shim)
    and runs UNetSpatioTemporalConditionModel.forward / ControlNetModel.forward. The oracle must reproduce those outputs
    to fp32 round-off on the same deterministic weights / inputs (tests/refpin.py).
  * live (only where /root/reference exists): regenerate one case in a subprocess and compare with the committed file,
    so the golden cannot drift from the reference source.
What this pins: the time_context quirk (transformer_temporal.py:309-319), skip ordering (unet_3d_blocks.py:2242,2352),
per-block eps, the residual merge (unet_spatio_temporal_condition.py:481-502), conv_in_concat / zero convs / scales /
guess_mode (temporal_controlnet.py:576-633), timestep forms (:399-414), the transformer_layers_per_block loop and the
head_dim-128 class default. The leaf layers stay a restatement of diffusers 0.25.1 (shim and oracle are written
independently: nn.Module vs functional), cross-checked against torch built-ins in tests/test_oracle.py.
"""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

from oracle import svd_oracle as O
from tests import refpin

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden" / "reference_pin.pt"
TOL = 2e-5  # fp32 vs fp32, different summation order only (measured 1e-6 .. 6e-6)


def _models(kind, F):
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    with torch.device("meta"):
        unet = UNetSpatioTemporalConditionModel(num_frames=F, **kind)
        cn = ControlNetModel(**kind)
    usd = refpin.fill_state_dict(((k, v.shape) for k, v in unet.state_dict().items()), seed=11)
    csd = refpin.fill_state_dict(((k, v.shape) for k, v in cn.state_dict().items()), seed=12)
    return usd, csd


def _rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def oracle_case(name):
    kind, B, F, h, w = refpin.CASES[name]
    usd, csd = _models(kind, F)
    cfg = dict(O.SVD_CONFIG)
    cfg.update({k: v for k, v in kind.items() if k in cfg})
    sample, ehs, ati, cond = refpin.make_inputs(B, F, h, w)
    cc = torch.cat([cond] * B)
    out = {}
    with torch.no_grad():
        out["unet"] = O.unet_forward(usd, cfg, sample, refpin.TIMESTEP, ehs, ati)
        down, mid = O.controlnet_forward(csd, cfg, sample, refpin.TIMESTEP, ehs, ati, cc, 0.75)
        out["cn_mid"] = mid
        out["cn_down_fp"] = torch.stack([refpin.fingerprint(d) for d in down])
        out["cn_down_shapes"] = [tuple(d.shape) for d in down]
        out["vgl"] = O.unet_forward(usd, cfg, sample, torch.tensor(refpin.TIMESTEP), ehs, ati, down, mid)
        if name == "tiny":
            gd, gm = O.controlnet_forward(csd, cfg, sample, torch.tensor([refpin.TIMESTEP]), ehs, ati, cc, 1.0, True)
            out["guess_mid"] = gm
            out["guess_down_fp"] = torch.stack([refpin.fingerprint(d) for d in gd])
    out["unet_keys"], out["cn_keys"] = len(usd), len(csd)
    out["unet_params"] = sum(v.numel() for v in usd.values())
    out["cn_params"] = sum(v.numel() for v in csd.values())
    return out


def _compare(ours, ref, name):
    assert ours["unet_keys"] == ref["unet_keys"] and ours["cn_keys"] == ref["cn_keys"], "state-dict key count differs"
    assert ours["unet_params"] == ref["unet_params"] and ours["cn_params"] == ref["cn_params"]
    assert [tuple(s) for s in ours["cn_down_shapes"]] == [tuple(s) for s in ref["cn_down_shapes"]]
    errs = {}
    for k in ("unet", "cn_mid", "vgl", "guess_mid"):
        if k in ref:
            assert ours[k].shape == ref[k].shape
            errs[k] = _rel(ours[k], ref[k])
    for k in ("cn_down_fp", "guess_down_fp"):
        if k in ref:
            # the three fp64 moments per residual (sum can cancel: compare against abs-sum) + the strided samples
            a, b = ours[k], ref[k]
            errs[k + ".abs"] = float(((a[:, 1] - b[:, 1]).abs() / b[:, 1]).max())
            errs[k + ".sq"] = float(((a[:, 2] - b[:, 2]).abs() / b[:, 2]).max())
            errs[k + ".sum"] = float(((a[:, 0] - b[:, 0]).abs() / b[:, 1]).max())
            errs[k + ".samples"] = _rel(a[:, 3:], b[:, 3:])
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, f"{name}: oracle differs from the reference's own forward: {bad} (all: {errs})"
    return errs


@pytest.fixture(scope="module")
def golden():
    assert GOLD.exists(), f"{GOLD} missing: run python tests/golden/make_reference_golden.py where /root/reference exists"
    return torch.load(GOLD, weights_only=False)


@pytest.mark.parametrize("name", ["tiny", "tiny_b1", "tiny_2layers", "svd_hd128"])
def test_oracle_matches_reference_golden(golden, name):
    _compare(oracle_case(name), golden[name], name)


@pytest.mark.slow
def test_oracle_matches_reference_golden_svd_width(golden):
    """SVD channel widths / head counts (320..1280, heads 5/10/20/20), B = 2, 14 x 16 x 24 latent."""
    errs = _compare(oracle_case("svd"), golden["svd"], "svd")
    assert golden["svd"]["unet_params"] == 1_524_623_082 and golden["svd"]["cn_params"] == 680_946_577, \
        "the reference's own modules (under the shim) must have the published SVD / GestureNet parameter counts"
    print(errs)


@pytest.mark.skipif(not Path("/root/reference/svd/unet_spatio_temporal_condition.py").exists(),
                    reason="reference tree not present (GPU box): the committed golden is used instead")
def test_golden_is_reproducible_from_reference_source(golden, tmp_path):
    """Re-run the reference's files in a subprocess (its own `svd` package, this repo's is never imported there)."""
    out = tmp_path / "pin.pt"
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "make_reference_golden.py"), "--out", str(out),
                        "--cases", "tiny"], cwd=str(ROOT), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    fresh = torch.load(out, weights_only=False)["tiny"]
    for k in ("unet", "cn_mid", "vgl", "guess_mid", "cn_down_fp"):
        assert torch.equal(fresh[k], golden["tiny"][k]) or _rel(fresh[k], golden["tiny"][k]) < 1e-6, k


def test_shim_is_not_importable_from_the_product():
    """The stand-in for diffusers lives under tests/ and is only ever put on sys.path by the golden generator."""
    import importlib.util
    assert importlib.util.find_spec("diffusers") is None or "diffusers_shim" not in (
        importlib.util.find_spec("diffusers").origin or "")
    for pkg in ("svd", "this_and_that_vdm_b200", "data_loader"):
        for f in (ROOT / pkg).rglob("*.py"):
            src = f.read_text()
            assert "diffusers_shim" not in src and "refpin" not in src, f
