"""Golden vectors for the gesture rasteriser, produced by EXECUTING the reference's own `get_thisthat_sam`
(/root/reference/data_loader/video_this_that_dataset.py:28-130) with the reference's own Gaussian kernel
(utils/optical_flow_utils.py:bivariate_Gaussian) and the real cv2 — run in the authoring container only:

    python tests/golden/make_gesture_golden.py          # writes tests/golden/gesture_golden.npz

The reference module itself cannot be imported here (moviepy is not installed), so the function's source text is
pulled out of the reference file with `ast` and executed unmodified in a namespace that holds exactly the names it
uses (os, np, cv2, blur_kernel). Nothing of the reference is copied into this repository: only inputs and outputs.
"""
import ast
import os
import sys
import tempfile

import cv2
import numpy as np

REF = "/root/reference"
SRC = os.path.join(REF, "data_loader", "video_this_that_dataset.py")

CASES = {
    # name: (org_h, org_w, out_h, out_w, data.txt lines "frame horizontal vertical", flip, dilate)
    "two_points_256x384": (480, 640, 256, 384, ["0 320.0 240.0", "13 100.5 400.9"], False, True),
    "corner_clipped": (120, 160, 64, 96, ["0 3 2", "13 158 119"], False, True),
    "same_size_flip": (64, 96, 64, 96, ["0 40 30", "5 70 10"], True, True),
    "no_dilate_upscale": (50, 60, 80, 112, ["0 30 25"], False, False),
    "same_frame_overwrite": (90, 100, 45, 50, ["2 20 20", "2 70 60"], False, True),
}


def load_reference_function():
    sys.path.insert(0, REF)
    from utils.optical_flow_utils import bivariate_Gaussian  # the reference's own kernel builder
    tree = ast.parse(open(SRC).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "get_thisthat_sam"][0]
    ns = {"os": os, "np": np, "cv2": cv2,
          "blur_kernel": bivariate_Gaussian(99, 10, 10, 0, grid=None, isotropic=True)}  # reference :26
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    return ns["get_thisthat_sam"]


def main():
    fn = load_reference_function()
    out = {}
    for name, (oh, ow, h, w, lines, flip, dilate) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            cv2.imwrite(os.path.join(d, "im_0.jpg"), np.zeros((oh, ow, 3), np.uint8))
            open(os.path.join(d, "data.txt"), "w").write("\n".join(lines))
            cfg = {"video_seq_length": 14, "conditioning_channels": 3, "height": h, "width": w, "dilate": dilate,
                   "motion_bucket_id": None}
            cond, bucket, idxs, coords = fn(cfg, d, store_dir=None, flip=flip, verbose=False)
        assert bucket == 200
        out[name + "/cond"] = cond.astype(np.float32)
        out[name + "/meta"] = np.array([oh, ow, h, w, int(flip), int(dilate)], np.int64)
        out[name + "/lines"] = np.array(lines)
        out[name + "/frames"] = np.array(idxs, np.int64)
        out[name + "/coords"] = np.array(coords, np.int64)
        print(name, cond.shape, float(cond.min()), float(cond.max()), idxs, coords)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gesture_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
