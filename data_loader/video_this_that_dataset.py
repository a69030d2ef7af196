"""`get_thisthat_sam` with the reference's signature and return value
(data_loader/video_this_that_dataset.py:28-130 of the reference; the same code is inlined in app.py:282-328), computed
on the B200 by the closed-form gesture rasteriser (this_and_that_vdm_b200/csrc/gesture.cu) instead of cv2 on the host.

Returns `(thisthat_condition [F,3,H,W] float32 numpy, motion_bucket_id, controlnet_image_index, coordinate_values)`
exactly like the reference, so `torch.from_numpy(...)` in the callers keeps working. Pass `as_tensor=True` (an
addition) to keep the condition on the device and skip the round trip.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from this_and_that_vdm_b200 import lib


def _frame_size(path: str):
    """(height, width) of im_0.jpg — the reference reads the whole image with cv2 just for its shape (:45-46)."""
    try:
        from PIL import Image
        with Image.open(path) as im:
            w, h = im.size
        return h, w
    except ImportError:
        import cv2
        img = cv2.imread(path)
        return img.shape[0], img.shape[1]


def rasterise(points, org_hw, out_hw, n_frames=14, dilate=True, flip=False, device=None) -> torch.Tensor:
    """points: (frame_idx, vertical, horizontal) in data.txt order -> fp32 [n_frames, 3, H, W] on the device."""
    lib.init()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    H, W = out_hw
    out = torch.empty(n_frames, 3, H, W, dtype=torch.float32, device=device)
    scratch = torch.empty(max(1, len(points)) * (H + W), dtype=torch.float32, device=device)
    lib.gesture_raster(points, out, scratch, org_h=org_hw[0], org_w=org_hw[1], dilate=dilate, flip=flip)
    return out


def get_thisthat_sam(config, intput_dir, store_dir=None, flip=False, verbose=False, as_tensor=False):
    with open(os.path.join(intput_dir, "data.txt"), "r") as f:
        lines = f.readlines()
    org_h, org_w = _frame_size(os.path.join(intput_dir, "im_0.jpg"))
    if config["conditioning_channels"] != 3:
        raise NotImplementedError()  # reference :101-102

    controlnet_image_index, coordinate_values, points = [], [], []
    for line in lines:
        frame_idx, horizontal, vertical = line.split(" ")  # reference :49-50
        frame_idx, vertical, horizontal = int(frame_idx), int(float(vertical)), int(float(horizontal))
        controlnet_image_index.append(frame_idx)
        coordinate_values.append((vertical, horizontal))
        points.append((frame_idx, vertical, horizontal))

    cond = rasterise(points, (org_h, org_w), (config["height"], config["width"]), n_frames=config["video_seq_length"],
                     dilate=bool(config["dilate"]), flip=flip)

    if store_dir is not None and verbose:  # reference :96-97: the un-normalised frames as condition_TT<idx>.png
        import cv2
        # (when two points share a frame the reference would have written each point's own image; here the stored
        # frame is the final content of that frame)
        for idx, frame_idx in enumerate(controlnet_image_index):
            img = (cond[frame_idx] * 255.0).permute(1, 2, 0).cpu().numpy()
            cv2.imwrite(os.path.join(store_dir, "condition_TT" + str(idx) + ".png"), img)

    motion_bucket_id = 200 if config["motion_bucket_id"] is None else config["motion_bucket_id"]  # reference :118-122
    thisthat_condition = cond if as_tensor else cond.cpu().numpy()
    return (thisthat_condition, motion_bucket_id, controlnet_image_index, coordinate_values)
