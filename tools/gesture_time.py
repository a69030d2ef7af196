"""Timing of the gesture rasteriser at the bench resolution: CUDA closed form vs the reference's cv2 path on the host."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from data_loader.video_this_that_dataset import rasterise
pts = [(0, 540, 960), (13, 1000, 100)]
for _ in range(3): out = rasterise(pts, (1080, 1920), (576, 1024))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = rasterise(pts, (1080, 1920), (576, 1024))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = out.numel() * 4
print("cuda: %.3f ms per call (2 points, 1080p -> 14x3x576x1024 fp32 = %.1f MB written: %.0f GB/s)" % (ms, nbytes / 1e6, nbytes / ms / 1e6))
try:
    import cv2
    from oracle.gesture_oracle import bivariate_gaussian_kernel
    k = bivariate_gaussian_kernel()
    t = time.time()
    cond = np.zeros((14, 3, 576, 1024), np.float32)
    for idx, (f, v, h) in enumerate(pts):
        base = np.full((1080, 1920, 3), 255.0, np.float32)
        base[max(v - 10, 0):v + 11, max(h - 10, 0):h + 11] = [0, 0, 255] if idx == 0 else [0, 255, 0]
        base = cv2.resize(cv2.filter2D(base, -1, k), (1024, 576), interpolation=cv2.INTER_CUBIC)
        cond[f] = (base / 255.0).transpose(2, 0, 1)
    print("cv2 on the host (the reference's operations, all cores cv2 uses): %.1f ms" % ((time.time() - t) * 1e3))
    print("max abs diff:", float(np.abs(out.cpu().numpy() - cond).max()))
except ImportError:
    print("cv2 not importable on this box")
