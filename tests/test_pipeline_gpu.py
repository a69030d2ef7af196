"""The whole drop-in on the B200, called like test_code/inference.py:246-270 calls the reference: PIL first frame, token
ids, numpy gesture condition -> CLIP towers -> conditioning assembly -> VAE encode -> 2 Euler steps of UNet + GestureNet
with CFG -> chunked VAE decode -> frames, every stage on the sm_100a kernels, against the same computation composed
from the three CPU oracles (tests/pipeline_case.py)."""
import pytest
import torch

from tests import pipeline_case as PC
from tests.common import rel_l2

pytestmark = pytest.mark.gpu


def test_vgl_pipeline_end_to_end_vs_oracles_on_gpu():
    from this_and_that_vdm_b200 import lib
    mods, sds = PC.build("cuda")
    with torch.no_grad():
        n0 = lib.launch_count()
        frames = PC.run_pipeline(mods, "cuda", output_type="pt")
        launches = lib.launch_count() - n0
        pil = PC.run_pipeline(mods, "cuda", output_type="pil")
        ref, _ = PC.run_oracle(sds)
    assert len(frames) == 1 and frames[0].shape == (PC.FRAMES, 3, PC.H, PC.W) and frames[0].is_cuda
    assert launches > 2500  # CLIP + VAE + denoiser kernels of libttvdm_sm100.so (measured: 2856); no library dispatch
    want = (ref[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert rel_l2(frames[0], want) < 5e-2
    assert len(pil) == 1 and len(pil[0]) == PC.FRAMES and pil[0][0].size == (PC.W, PC.H)


def test_vl_pipeline_end_to_end_without_text_on_gpu():
    mods, sds = PC.build("cuda")
    with torch.no_grad():
        frames = PC.run_vl_pipeline(mods, "cuda", output_type="pt")
        ref = PC.run_vl_oracle(sds)
    want = (ref[0].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)
    assert frames[0].shape == want.shape and rel_l2(frames[0], want) < 5e-2
