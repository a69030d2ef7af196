"""Timing of the conditioning builder (scope-table next #2) on one B200 at the published sizes: CLIP ViT-H/14 image
tower (32 layers, 1280 wide, 16 heads of 80, 257 tokens), SD-2.1 text tower (23 layers, 1024 wide, 16 heads, 77 causal
tokens), assembly; random-init weights, synthetic inputs. CUDA events after a warm-up pass; CPU baseline = the oracle on
all host threads on the same inputs.   python tools/cond_time.py [--out gpurun_out/cond_time.json]"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import clip_oracle as CO  # noqa: E402  (checker / CPU baseline only)
from tests.common import clip_text_sd, clip_vision_sd, rel_l2  # noqa: E402
from this_and_that_vdm_b200 import lib  # noqa: E402
from this_and_that_vdm_b200.clip_engine import ClipTowerEngine, assemble_conditioning  # noqa: E402

VIT_H = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
             patch_size=14, projection_dim=1024, hidden_act="gelu")
SD21_TEXT = dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                 num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu")


def timed(fn, iters=3):
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
    ap.add_argument("--out", default="gpurun_out/cond_time.json")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    lib.init(0)
    vsd, tsd = clip_vision_sd(VIT_H), clip_text_sd(SD21_TEXT)
    vis = ClipTowerEngine(vsd, VIT_H, "vision", "cuda:0")
    txt = ClipTowerEngine(tsd, SD21_TEXT, "text", "cuda:0")
    g = torch.Generator().manual_seed(2)
    px = torch.randn(1, 3, 224, 224, generator=g)
    ids = torch.randint(0, 49408, (1, 77), generator=g)
    pxd, idd = px.cuda(), ids.cuda()
    res = {"config": {"workload": "encode_clip: ViT-H/14 image tower + SD-2.1 text tower + assembly, 1 video"}}
    with torch.no_grad():
        n0 = lib.launch_count()
        emb = vis.image_embeds(pxd)
        n1 = lib.launch_count()
        hs = txt.last_hidden_state(idd)
        n2 = lib.launch_count()
        res["launches"] = {"vision": n1 - n0, "text": n2 - n1}
        res["vision_ms"] = timed(lambda: vis.image_embeds(pxd))
        res["text_ms"] = timed(lambda: txt.last_hidden_state(idd))
        res["assemble_ms"] = timed(lambda: assemble_conditioning(emb, hs, True))
        res["total_ms"] = res["vision_ms"] + res["text_ms"] + res["assemble_ms"]
        # the same towers replayed as one CUDA graph each (what the drop-in modules do by default)
        eg = vis.image_embeds(pxd, use_graph=True)
        hg = txt.last_hidden_state(idd, use_graph=True)
        res["graph_equals_eager"] = bool(torch.equal(eg, emb) and torch.equal(hg, hs))
        res["vision_graph_ms"] = timed(lambda: vis.image_embeds(pxd, use_graph=True))
        res["text_graph_ms"] = timed(lambda: txt.last_hidden_state(idd, use_graph=True))
        res["total_graph_ms"] = res["vision_graph_ms"] + res["text_graph_ms"] + res["assemble_ms"]
        if not a.no_cpu:
            torch.set_num_threads(os.cpu_count())
            t0 = time.time()
            ref_e = CO.vision_image_embeds(vsd, px, 16, "gelu")
            ref_t = CO.text_last_hidden_state(tsd, ids, 16, "gelu")
            ref = CO.assemble(ref_e, ref_t, True)
            res["cpu_baseline"] = {"kind": "port", "cores": os.cpu_count(), "value_ms": (time.time() - t0) * 1e3,
                                   "sample": "the same call (full size) through oracle/clip_oracle.py"}
            res["rel_l2_vs_oracle"] = {"image_embeds": rel_l2(emb, ref_e), "text": rel_l2(hs, ref_t),
                                       "conditioning": rel_l2(assemble_conditioning(emb, hs, True), ref)}
    Path(a.out).parent.mkdir(parents=True, exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
