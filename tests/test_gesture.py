"""Gesture rasteriser: oracle pinned against cv2 and against golden vectors produced by the
reference's own get_thisthat_sam (tests/golden/make_gesture_golden.py); CUDA path checked against both.
Tolerance (floating point, values in [0, 1]): 5e-6 absolute — cv2's float32 DFT filter and float32 separable resize
round differently from any other evaluation order by ~1e-6."""
import os
import tempfile

import numpy as np
import pytest

from oracle import gesture_oracle as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gesture_golden.npz")
TOL = 5e-6


def _cases():
    z = np.load(GOLDEN)
    for name in sorted({k.split("/")[0] for k in z.files}):
        oh, ow, h, w, flip, dil = (int(v) for v in z[name + "/meta"])
        yield name, (oh, ow), (h, w), bool(flip), bool(dil), [str(s) for s in z[name + "/lines"]], z[name + "/cond"]


CASES = list(_cases())
SMALL = [c for c in CASES if c[1][0] * c[1][1] <= 160 * 120]  # the direct 99x99 oracle filter is O(HW * 9801)


def test_golden_holds_every_case():
    assert {c[0] for c in CASES} == {"two_points_256x384", "corner_clipped", "same_size_flip", "no_dilate_upscale",
                                     "same_frame_overwrite"}
    for name, _, (h, w), _, _, lines, cond in CASES:
        assert cond.shape == (14, 3, h, w) and cond.dtype == np.float32
        frames = {int(l.split(" ")[0]) for l in lines}
        for f in range(14):  # frames without a gesture point are exactly zero
            assert (f in frames) or not cond[f].any(), name


@pytest.mark.parametrize("case", SMALL, ids=[c[0] for c in SMALL])
def test_oracle_matches_reference_golden(case):
    name, org, out, flip, dilate, lines, cond = case
    got = G.rasterise(G.parse_data_txt(lines), org, out, dilate=dilate, flip=flip)
    assert np.abs(got - cond).max() < TOL


def test_oracle_primitives_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 255, (50, 66, 3)).astype(np.float32)
    k = G.bivariate_gaussian_kernel()
    assert abs(k.sum() - 1.0) < 1e-12 and k.shape == (99, 99)
    assert np.abs(G.filter2d_reflect101(img, k) - cv2.filter2D(img, -1, k)).max() < 1e-3  # on the 0..255 scale
    for (w, h) in [(48, 32), (66, 50), (131, 77), (20, 90)]:
        ref = cv2.resize(img, (w, h), interpolation=cv2.INTER_CUBIC)
        assert np.abs(G.resize_cubic(img, w, h) - ref).max() < 5e-3, (w, h)


def test_parse_data_txt():
    assert G.parse_data_txt(["0 320.7 240.2\n", "13 100 400\n", "\n"]) == [(0, 240, 320), (13, 400, 100)]


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_reference_golden(case):
    from data_loader.video_this_that_dataset import rasterise
    from this_and_that_vdm_b200 import lib
    name, org, out, flip, dilate, lines, cond = case
    n0 = lib.launch_count() if lib._lib is not None else 0
    got = rasterise(G.parse_data_txt(lines), org, out, dilate=dilate, flip=flip).cpu().numpy()
    assert lib.launch_count() > n0, "no sm_100a kernel was launched"
    assert np.abs(got - cond).max() < TOL, name


@pytest.mark.gpu
def test_cuda_matches_oracle_on_random_points():
    from data_loader.video_this_that_dataset import rasterise
    rng = np.random.default_rng(3)
    for trial in range(4):
        oh, ow = int(rng.integers(40, 110)), int(rng.integers(40, 130))
        h, w = int(rng.integers(16, 90)), int(rng.integers(4, 30)) * 4 + int(trial % 2)  # odd widths: scalar stores
        pts = [(int(rng.integers(0, 14)), int(rng.integers(-15, oh + 15)), int(rng.integers(-15, ow + 15)))
               for _ in range(int(rng.integers(1, 4)))]
        flip, dilate = bool(trial & 1), trial != 2
        ref = G.rasterise(pts, (oh, ow), (h, w), dilate=dilate, flip=flip)
        got = rasterise(pts, (oh, ow), (h, w), dilate=dilate, flip=flip).cpu().numpy()
        assert np.abs(got - ref).max() < TOL, (trial, pts, oh, ow, h, w)


@pytest.mark.gpu
def test_dropin_get_thisthat_sam_and_full_size_properties():
    """The reference's entry point on a 1080p frame at the bench resolution (576 x 1024): size-independent properties."""
    from PIL import Image
    from data_loader.video_this_that_dataset import get_thisthat_sam
    cfg = {"video_seq_length": 14, "conditioning_channels": 3, "height": 576, "width": 1024, "dilate": True,
           "motion_bucket_id": None}
    with tempfile.TemporaryDirectory() as d:
        Image.new("RGB", (1920, 1080)).save(os.path.join(d, "im_0.jpg"))
        open(os.path.join(d, "data.txt"), "w").write("0 960.4 540.9\n13 100 1000")
        cond, bucket, idxs, coords = get_thisthat_sam(cfg, d)
        flipped = get_thisthat_sam(cfg, d, flip=True)[0]
    assert isinstance(cond, np.ndarray) and cond.dtype == np.float32 and cond.shape == (14, 3, 576, 1024)
    assert bucket == 200 and idxs == [0, 13] and coords == [(540, 960), (1000, 100)]
    assert not cond[1:13].any()                                   # 12 of the 14 frames carry no gesture
    assert np.all(cond[0, 2] == 1.0) and np.all(cond[13, 1] == 1.0)  # the colour channel of each point stays at 255
    assert np.array_equal(cond[0, 0], cond[0, 1]) and np.array_equal(cond[13, 0], cond[13, 2])
    assert cond.min() > -1e-3 and cond.max() <= 1.0 + 1e-6
    # 21-wide box blurred by sigma 10: (P(|z| < 1.05))^2 = 0.50 of the colour survives at the centre; white far away
    assert 0.45 < cond[0, 0, 288, 512] < 0.55 and cond[0, 0, 10, 10] == 1.0
    assert np.array_equal(flipped, cond[..., ::-1])               # np.fliplr of the reference
