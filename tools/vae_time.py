"""Timing + full-size parity of the VAE path (scope-table next #1) on one B200.

  python tools/vae_time.py [--height 576 --width 1024 --frames 14 --iters 3] [--no-cpu] [--out gpurun_out/vae_time.json]

Random-init AutoencoderKLTemporalDecoder of the published SVD shape (128/256/512/512), synthetic latents / images.
Reports, with CUDA events on the launching stream after a warm-up pass:
  * decode_latents of one 14-frame video through the drop-in pipeline (chunk = 14, the reference's default, and 8),
  * vae.encode of the first frame + the 14 gesture frames (12 of them all-zero, as the reference's rasteriser produces),
    with and without de-duplication,
  * per-kernel-family time shares and achieved TFLOP/s of the GEMM family (lib.start_profile),
  * full-size parity: a 2-frame decode and a 1-image encode at the full resolution against the CPU fp32 oracle, and
    "videos never mix" (two 7-frame videos in one call == each alone),
  * the CPU baseline: the oracle's decode of 2 frames at 256x384 on all host threads, extrapolated by pixel*frame count.
Nothing here is a bench.py value; the numbers go to profiles/ as supporting evidence for DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import vae_oracle as VO  # noqa: E402  (checker / CPU baseline only)
from svd.pipeline_common import SVDPipelineBase  # noqa: E402
from tests.common import SVD_VAE, build_vae, rel_l2, state  # noqa: E402
from this_and_that_vdm_b200 import lib  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--out", default="gpurun_out/vae_time.json")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    lib.init(0)
    H, W, Fr = a.height, a.width, a.frames
    h, w = H // 8, W // 8
    vae = build_vae(SVD_VAE)
    sd = state(vae)
    vae.to("cuda")
    pipe = SVDPipelineBase(vae=vae, unet=None)
    g = torch.Generator().manual_seed(11)
    lat = (torch.randn(1, Fr, 4, h, w, generator=g) * 0.18215).cuda()
    res = {"config": {"workload": f"SVD VAE (random init), {Fr}x{H}x{W}", "height": H, "width": W, "frames": Fr},
           "gpu": torch.cuda.get_device_name(0)}
    with torch.no_grad():
        # ---- decode
        n0 = lib.launch_count()
        pipe.decode_latents(lat, Fr, Fr)
        res["decode_launches"] = lib.launch_count() - n0
        res["decode_ms_chunk_all"] = timed(lambda: pipe.decode_latents(lat, Fr, Fr), a.iters)
        res["decode_ms_chunk8"] = timed(lambda: pipe.decode_latents(lat, Fr, 8), a.iters)
        res["decode_peak_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
        lib.start_profile()
        pipe.decode_latents(lat, Fr, Fr)
        fam = defaultdict(lambda: [0.0, 0.0, 0])
        for name, info, ms in lib.stop_profile():
            f = fam[name]
            f[0] += ms
            f[1] += info.get("flops", 0.0)
            f[2] += 1
        tot = sum(v[0] for v in fam.values())
        res["decode_families"] = {k: {"ms": round(v[0], 3), "share": round(v[0] / tot, 4), "calls": v[2],
                                      "tflops": round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[1] else None,
                                      "flops": v[1]}
                                  for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}
        res["decode_flops"] = sum(v[1] for v in fam.values())
        # ---- encode: first frame + 14 gesture frames (2 with a point, 12 all-zero)
        img = torch.rand(1, 3, H, W, generator=g) * 2 - 1
        cond = torch.zeros(Fr, 3, H, W)
        cond[0] = torch.rand(3, H, W, generator=g)
        cond[-1] = torch.rand(3, H, W, generator=g)
        batch = torch.cat([img, cond]).cuda()
        eng = vae._get_engine()
        res["encode_ms_dedupe"] = timed(lambda: eng.encode(batch), max(1, a.iters - 1))
        res["encode_ms_all"] = timed(lambda: eng.encode(batch, dedupe=False), 1)
        res["encode_images"] = int(batch.shape[0])
        # ---- full-size properties
        two = vae.decode(lat[0, :Fr // 2 * 2].div(0.18215), num_frames=Fr // 2).sample
        one = vae.decode(lat[0, :Fr // 2].div(0.18215), num_frames=Fr // 2).sample
        res["videos_never_mix_rel"] = rel_l2(two[:Fr // 2], one)
        del two, one
        if not a.no_cpu:
            torch.set_num_threads(os.cpu_count())
            z2 = lat[0, :2].div(0.18215).cpu()
            t0 = time.time()
            ref = VO.decode(sd, z2, 2)
            res["oracle_full_res_2frames_s"] = time.time() - t0
            out = vae.decode(z2.cuda(), num_frames=2).sample
            res["decode_full_res_rel_l2_vs_oracle"] = rel_l2(out, ref)
            del out, ref
            t0 = time.time()
            ref_e = VO.encode(sd, img)
            res["oracle_full_res_encode_s"] = time.time() - t0
            enc = vae.encode(img.cuda()).latent_dist.mode()
            res["encode_full_res_rel_l2_vs_oracle"] = rel_l2(enc, ref_e)
            # CPU baseline: bounded sample, extrapolated by frames * pixels (every op of the decoder is linear in both
            # except the per-frame attention, < 5 % of the FLOPs)
            zs = torch.randn(2, 4, 32, 48, generator=g)
            VO.decode(sd, zs[:1], 1)
            t0 = time.time()
            VO.decode(sd, zs, 2)
            dt = time.time() - t0
            scale = (Fr * H * W) / (2 * 256 * 384)
            res["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "sample": "oracle decode of 2 frames at 256x384",
                                   "sample_s": dt, "extrapolated_decode_s": dt * scale}
    res["decode_tflops_overall"] = res["decode_flops"] / (res["decode_ms_chunk_all"] * 1e-3) / 1e12
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
