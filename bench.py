#!/usr/bin/env python
"""bench.py — frames/s of 25-step VGL (UNet + GestureNet, CFG) generation on B200, the metric of BASELINE.json.

A "step" of this bench is ONE VIDEO: a full 25-Euler-step VGL denoising of one 14-frame clip (CFG pair, B = 2)
from synthetic latents with random-init weights — the hot path of
svd/pipeline_stable_video_diffusion_controlnet.py:623-720. frames/s = n_gpus * 14 / seconds_per_video (weak
scaling: every rank denoises its own video; one conditioning broadcast + one latent gather per round of videos).

  value    : device-resident inputs, FusedDenoiser driven directly (hoists included), CUDA events, max over ranks
  e2e      : the same videos through the public drop-in API (StableVideoDiffusionControlNetPipeline.__call__ in
             latent mode) with pinned HOST inputs and a host read-back of the latents inside the timed region
  roofline : tensor roofline of the dominant kernel family (gemm_kernel: all linear / conv3x3 / temporal conv),
             from a CUDA-event pass over one Euler step; per-family shares of the step are reported beside it
  cpu_baseline / --impl reference : the CPU fp32 oracle (oracle/svd_oracle.py — the reference's torch path
             restated; diffusers is not installable here) timed on the host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FRAMES = 14
NUM_STEPS = 25
# The CPU baseline always uses the same number of host threads (BENCH and SCALE boxes expose different core counts; a
# fixed figure keeps the driver's value / reference ratio comparable between the two files).
CPU_BASELINE_THREADS = min(16, os.cpu_count() or 1)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {"bf16_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))),
                    "bf16_tflops_burst": float(d.get("bf16_tflops", 0.0)),
                    "hbm_gbs": float(d.get("hbm_gbs", 0.0)), "source": "measured (MEASURED_PEAKS.json, sustained)"}
        except Exception:  # noqa: BLE001
            pass
    # fallback stated in /opt/skills/guides/B200_PROFILING.md (sustained ~1.4 PF inside a long step; burst 1.59)
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md: ~1.4 PF sustained, 1.59 PF burst, 6.65 TB/s)"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_median": sorted(power)[len(power) // 2],
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_gemm_traffic() -> dict:
    """roofline.traffic: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per gemm_kernel launch, averaged over the
    launches sampled by the committed `ncu --set full` capture of the same step (profiles/r02_ncu_full_gemm.csv, written by
    tools/run_profile.sh + tools/ncu_summarize.py). The family has 537 launches of different shapes, so this is a sample
    mean next to the algorithmic bytes, not a per-shape figure (those: profiles/r02_ncu_full_gemm_level0_shapes.csv)."""
    import csv
    f = ROOT / "profiles" / "r02_ncu_full_gemm.csv"
    try:
        rows = list(csv.reader(open(f)))
        h, units = rows[0], rows[1]
        ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir]) * mult[units[ir]] + float(r[iw]) * mult[units[iw]] for r in rows[2:]]
        return {"traffic": round(sum(vals) / len(vals)), "traffic_unit": "bytes / launch",
                "traffic_note": f"mean DRAM read + write of the {len(vals)} gemm_kernel launches sampled by ncu --set full "
                                f"(profiles/r02_ncu_full_gemm.csv; per-shape captures of the level-0 linears in "
                                f"profiles/r02_ncu_full_gemm_level0_shapes.csv show DRAM bytes = algorithmic A + residual + C bytes)"}
    except Exception as e:  # noqa: BLE001
        return {"traffic": None, "traffic_note": f"profiles/r02_ncu_full_gemm.csv unreadable: {e}"[:200]}



def synth_inputs(h: int, w: int, n_videos: int, seed: int = 0):
    """BASELINE.md §4 synthetic inputs, generated on CPU in fp32 so every arm sees identical bits."""
    from oracle_free_inputs import make  # local helper below (kept import-free of oracle/)
    return make(h, w, n_videos, seed)


def build_models(device):
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    torch.manual_seed(1234)
    unet = UNetSpatioTemporalConditionModel(num_attention_heads=(5, 10, 20, 20), num_frames=FRAMES).eval()
    cn = ControlNetModel().eval()
    g = torch.Generator().manual_seed(1236)
    with torch.no_grad():  # GestureNet zero-inits -> non-zero (otherwise VGL == VL), mix factors as initialised
        for name, p in cn.named_parameters():
            if name.startswith("controlnet_") or name.startswith("conv_in_concat"):
                fan_in = p[0].numel() if p.ndim > 1 else p.numel()
                p.copy_(torch.randn(p.shape, generator=g) * (fan_in ** -0.5 if p.ndim > 1 else 0.1))
    return unet.to(device), cn.to(device)


# ====================================================================================================== our arm
def run_ours(args):
    import torch.distributed as dist
    from this_and_that_vdm_b200 import lib
    from this_and_that_vdm_b200.sampler import FusedDenoiser
    from svd.pipeline_stable_video_diffusion_controlnet import StableVideoDiffusionControlNetPipeline
    from svd.scheduler import EulerDiscreteScheduler
    from tools.flop_census import step_flops, step_flops_split

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.init(local_rank)
    h, w = args.height // 8, args.width // 8
    unet, cn = build_models(dev)
    sched = EulerDiscreteScheduler()
    sched.set_timesteps(NUM_STEPS)
    sigmas, timesteps = sched.sigmas, sched.timesteps
    guidance = torch.linspace(1.0, 3.0, FRAMES)
    from this_and_that_vdm_b200.sharding import broadcast_conditioning, gather_latents

    # K videos per round (default: one per rank = weak scaling); rank 0 owns the conditioning of all of them. The batch is
    # partitioned by sharding.plan(): whole CFG pairs per rank when K >= N (no per-step communication), split pairs
    # (uncond / cond halves on two ranks, one 2-rank exchange of the noise prediction per step) when K < N.
    from this_and_that_vdm_b200.sharding import plan, run_sharded
    K = args.videos if args.videos > 0 else world
    cond_all = synth_inputs(h, w, K, seed=0) if rank == 0 else None
    vgl = not args.vl
    if args.vl:
        args.no_eager = args.no_cpu_baseline = args.no_full_pipeline = True
        if world > 1:
            raise SystemExit("--vl is a single-GPU line")
    den = FusedDenoiser(unet._get_engine(), cn._get_engine() if vgl else None)
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    my_plan = plan(K, world)[rank]
    split_pairs = sum(1 for a in plan(K, world)[0] if a.b_local == 1) > 0 or any(
        a.b_local == 1 for r in plan(K, world) for a in r)

    def one_video_resident(c, v) -> torch.Tensor:
        idx = [v, K + v]
        state = c["latents"][v].clone().contiguous()
        den.prepare(c["encoder_hidden_states"][idx], c["image_latents"][idx], c["added_time_ids"][idx], sigmas,
                    timesteps, guidance, num_frames=FRAMES, height=h, width=w,
                    controlnet_cond=c["controlnet_cond"][v] if vgl else None)
        for i in range(NUM_STEPS):
            den.step(i, state)
        return state

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def round_resident():
        if world > 1:
            return run_sharded(K, cond_dev if rank == 0 else None, dev, lambda: den, sigmas, timesteps, guidance)
        st = None
        for v in range(K):
            st = one_video_resident(cond_dev, v)
        return st

    cond_dev = {k: v.to(dev) for k, v in cond_all.items()} if rank == 0 else None
    if args.profile_only:
        # ncu helper: `ncu --profile-from-start off ... python bench.py --profile-only` captures exactly ONE Euler step
        c = cond_dev
        state = c["latents"][0].clone().contiguous()
        den.prepare(c["encoder_hidden_states"][[0, K]], c["image_latents"][[0, K]], c["added_time_ids"][[0, K]],
                    sigmas, timesteps, guidance, num_frames=FRAMES, height=h, width=w,
                    controlnet_cond=c["controlnet_cond"][0] if vgl else None)
        den.step(0, state)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        den.step(1, state)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_only": True, "launches_in_range": "one VGL Euler step"}))
        return
    for _ in range(args.warmup):
        round_resident()
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between videos (the working set is >> L2 anyway)
        round_resident()
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_video = ms / args.steps          # one "step" of this bench = one round of K videos
    value = K * FRAMES / (ms_per_video / 1e3)

    # ---------------- e2e through the public API with pinned HOST buffers (H2D of the inputs and D2H of the result inside
    # the timed region). N = 1: StableVideoDiffusionControlNetPipeline.__call__ (latent mode), K videos one after the
    # other. N > 1: sharding.run_sharded — rank 0 uploads the pinned conditioning pack, ONE broadcast, every rank
    # denoises its share, ONE gather, rank 0 reads the latents back.
    n_e2e = max(5, min(args.steps, 8)) if not args.quick_e2e else 1
    if world == 1:
        if vgl:
            pipe = StableVideoDiffusionControlNetPipeline.from_pretrained(None, unet=unet).to(dev)
        else:
            from svd.pipeline_stable_video_diffusion import StableVideoDiffusionPipeline
            pipe = StableVideoDiffusionPipeline.from_pretrained(None, unet=unet).to(dev)
        host = synth_inputs(h, w, 1, seed=rank)
        if not vgl:
            host.pop("controlnet_cond")
        host = {k: v.pin_memory() for k, v in host.items()}
        h2d = sum(v.numel() * v.element_size() for v in host.values()) * K
        out_host = torch.empty(1, FRAMES, 4, h, w, dtype=torch.float32).pin_memory()
        d2h = out_host.numel() * 4 * K
        e2e_api = ("StableVideoDiffusionControlNetPipeline" if vgl else "StableVideoDiffusionPipeline") + ".__call__ (latent mode)"

        def one_round_e2e():
            for _ in range(K):
                extra = dict(controlnet=cn, guess_mode=False,
                             controlnet_cond_latents=host["controlnet_cond"][0].to(dev, non_blocking=True)) if vgl else {}
                res = pipe(height=args.height, width=args.width, num_frames=FRAMES,
                           num_inference_steps=NUM_STEPS, min_guidance_scale=1.0, max_guidance_scale=3.0, fps=7,
                           motion_bucket_id=200, noise_aug_strength=0.1, output_type="latent",
                           latents=host["latents"].to(dev, non_blocking=True) / sched.init_noise_sigma,
                           encoder_hidden_states=host["encoder_hidden_states"].to(dev, non_blocking=True),
                           image_latents=host["image_latents"].to(dev, non_blocking=True), **extra)
                out_host.copy_(res.frames, non_blocking=True)
            torch.cuda.synchronize()
    else:
        host = {k: v.pin_memory() for k, v in cond_all.items()} if rank == 0 else None
        h2d = sum(v.numel() * v.element_size() for v in host.values()) if rank == 0 else 0
        out_host = torch.empty(K, FRAMES, 4, h, w, dtype=torch.float32).pin_memory() if rank == 0 else None
        d2h = out_host.numel() * 4 if rank == 0 else 0
        e2e_api = "this_and_that_vdm_b200.sharding.run_sharded (pinned host conditioning -> latents on the host)"

        def one_round_e2e():
            c = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if rank == 0 else None
            final = run_sharded(K, c, dev, lambda: den, sigmas, timesteps, guidance)
            if rank == 0:
                out_host.copy_(final, non_blocking=True)
            torch.cuda.synchronize()

    one_round_e2e()  # warm
    sync()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        one_round_e2e()
    sync()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = K * FRAMES / e2e_s

    # ---------------- roofline pass: CUDA events around every C-ABI call of ONE Euler step (rank 0)
    roof = None
    shares = None
    if rank == 0:
        peaks = load_peaks()
        c = cond_dev
        idx = [0, K]
        state = c["latents"][0].clone().contiguous()
        den.prepare(c["encoder_hidden_states"][idx], c["image_latents"][idx], c["added_time_ids"][idx], sigmas,
                    timesteps, guidance, num_frames=FRAMES, height=h, width=w, controlnet_cond=c["controlnet_cond"][0])
        den.step(0, state)
        torch.cuda.synchronize()
        lib.start_profile()
        den.step(1, state)
        rec = lib.stop_profile()
        fam = {}
        for name, info, ms_k in rec:
            f = fam.setdefault(name, {"ms": 0.0, "flops": 0.0, "launches": 0})
            f["ms"] += ms_k; f["flops"] += info.get("flops", 0.0); f["launches"] += 1
        shp = {}
        for name, info, ms_k in rec:
            if name == "ttvdm_gemm":
                k = (info["mode"], info["M"], info["N"], info["K"])
                d = shp.setdefault(k, [0.0, 0.0, 0])
                d[0] += ms_k; d[1] += info["flops"]; d[2] += 1
        gemm_shapes = [{"mode": k[0], "M": k[1], "N": k[2], "K": k[3], "n": v[2], "ms": round(v[0], 3),
                        "tflops": round(v[1] / v[0] / 1e9, 1)} for k, v in sorted(shp.items(), key=lambda kv: -kv[1][0])[:14]]
        tot_ms = sum(f["ms"] for f in fam.values())
        shares = {k.replace("ttvdm_", ""): {"share": round(v["ms"] / tot_ms, 4), "ms": round(v["ms"], 3),
                                             "launches": v["launches"],
                                             "tflops": round(v["flops"] / v["ms"] / 1e9, 1) if v["flops"] else None}
                  for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        gemm_alg, attn_alg, _ = step_flops_split(h, w, 2, vgl)
        g = fam.get("ttvdm_gemm", {"ms": 1.0, "launches": 1})
        achieved = gemm_alg / g["launches"] / (g["ms"] / g["launches"]) / 1e9  # TFLOP/s: alg flops per launch / avg ms
        roof = {"kernel": "gemm_kernel (linear + conv3x3 + temporal-conv family)", "bound": "tensor",
                "achieved": round(achieved, 1), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(achieved / peaks["bf16_tflops"], 4), **ncu_gemm_traffic(),
                "peak_source": peaks["source"],
                "launches_per_step": g["launches"], "avg_launch_ms": round(g["ms"] / g["launches"], 4),
                "algorithmic_tflop_per_step": round(gemm_alg / 1e12, 3),
                "algorithmic_note": "FLOPs of the reference's operators (tools/flop_census.py). The three Upsample2D convolutions "
                                    "run as 2x2-tap parity convolutions on the low-resolution input: 16/36 of their algorithmic "
                                    "FLOPs are executed (-2.4 TF per VGL step at 576x1024, 2 % of the total)",
                "whole_step": {"algorithmic_tflop": round(step_flops(h, w, 2, vgl) / 1e12, 3),
                               "achieved_tflops": round(step_flops(h, w, 2, vgl) * NUM_STEPS / (ms_per_video / 1e3) / 1e12, 1),
                               "frac": round(step_flops(h, w, 2, vgl) * NUM_STEPS / (ms_per_video / 1e3) / 1e12 / peaks["bf16_tflops"], 4)},
                "attention": {"algorithmic_tflop_per_step": round(attn_alg / 1e12, 3),
                              "achieved_tflops": round(attn_alg / (fam.get("ttvdm_attn_spatial", {"ms": 1})["ms"] + fam.get("ttvdm_attn_cross", {"ms": 0})["ms"]) / 1e9, 1)}}

    # ---------------- CPU baseline (rank 0, N = 1 only)
    full = None
    if rank == 0 and world == 1 and not args.no_full_pipeline and not args.profile_only:
        try:
            full = full_pipeline_extra(args, dev, unet, cn)
        except Exception as e:  # noqa: BLE001  (an extra must never cost the bench line)
            full = {"error": f"{type(e).__name__}: {e}"[:300]}
    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        try:
            eager = eager_bf16_gpu(dev, unet, cn, sizes=((32, 48), (h, w)) if (h, w) != (32, 48) else ((32, 48),))
        except Exception as e:  # noqa: BLE001
            eager = {"error": f"{type(e).__name__}: {e}"[:300]}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference(args, sample_only=True)

    if rank == 0:
        line = {
            "metric": "frames/sec for 14-frame 576x1024 VGL, 25 Euler steps" if (args.height, args.width) == (576, 1024) and vgl
            else f"frames/sec for 14-frame {args.height}x{args.width} {'VGL' if vgl else 'VL'}, 25 Euler steps",
            "value": round(value, 4), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_video, 2), "higher_is_better": True,
            "scaling": "weak" if args.videos <= 0 else "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init weights, seeded N(0,1) latents/conditioning)",
            "config": {"workload": f"{'VGL (UNet+GestureNet)' if vgl else 'VL (UNet only)'} 25-step Euler, 14x{args.height}x{args.width}, CFG pair (B=2) per video, "
                                   f"{K} video(s) per round on {world} GPU(s)", "videos": K,
                       "step": f"one round = {K} video(s) x 25 Euler steps", "latent": [FRAMES, 4, h, w],
                       "parallelism": f"dp{world}: " + ("split CFG pairs (uncond / cond halves on two ranks, one 2-rank exchange of "
                                                        "the noise prediction per Euler step)" if split_pairs else
                                                        "whole CFG pairs per rank, no per-step communication") +
                                      "; 1 conditioning broadcast + 1 latent gather per round",
                       "l2": "192 MB flush write between videos; per-step working set >> 126 MB L2"},
            "e2e": {"value": round(e2e_value, 4), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "api": e2e_api, "rounds_timed": n_e2e, "videos_timed": n_e2e * K},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernel_shares": shares,
            "gemm_shapes": gemm_shapes if rank == 0 else None,
            "cpu_baseline": cpu,
            "eager_bf16_gpu": eager,
            "full_pipeline": full,
        }
        print(json.dumps(line), flush=True)
        try:
            (ROOT / "gpurun_out").mkdir(exist_ok=True)
            (ROOT / "gpurun_out" / f"bench_last_{args.height}x{args.width}_n{world}_k{K}.json").write_text(json.dumps(line, indent=1))
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        dist.destroy_process_group()


# ============================================================================================ image -> frames (extra)
def full_pipeline_extra(args, dev, unet, cn):
    """NOT the headline metric (that is the denoising loop): one video through the whole drop-in the way
    test_code/inference.py calls the reference — PIL first frame + token ids + numpy gesture frames -> CLIP towers
    (ViT-H/14 + SD-2.1 text shapes) -> conditioning -> VAE encode -> 25 Euler steps -> VAE decode -> frames on the host —
    every stage on libttvdm_sm100.so, random-init weights. Reported under "full_pipeline" next to the stage times."""
    import numpy as np
    import PIL.Image
    from svd.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from svd.clip_towers import CLIPTextModel, CLIPVisionModelWithProjection
    from svd.pipeline_stable_video_diffusion_controlnet import StableVideoDiffusionControlNetPipeline
    torch.manual_seed(4321)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                                       down_block_types=("DownEncoderBlock2D",) * 4).eval().to(dev)
    vis = CLIPVisionModelWithProjection(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                                        num_attention_heads=16, image_size=224, patch_size=14, projection_dim=1024,
                                        hidden_act="gelu").to(dev)
    txt = CLIPTextModel(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                        num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu").to(dev)
    g = torch.Generator().manual_seed(77)
    image = PIL.Image.fromarray((torch.rand(args.height, args.width, 3, generator=g) * 255).to(torch.uint8).numpy())
    ids = torch.randint(0, 49408, (1, 77), generator=g).to(dev)
    cond = np.zeros((FRAMES, 3, args.height, args.width), dtype=np.float32)
    cond[0] = torch.rand(3, args.height, args.width, generator=g).numpy()
    cond[-1] = torch.rand(3, args.height, args.width, generator=g).numpy()
    pipe = StableVideoDiffusionControlNetPipeline.from_pretrained(None, vae=vae, image_encoder=vis, unet=unet).to(dev)

    def one():
        out = pipe(image, cond, controlnet=cn, prompt=ids, use_text=True, text_encoder=txt, height=args.height,
                   width=args.width, num_frames=FRAMES, num_inference_steps=NUM_STEPS, decode_chunk_size=8, fps=7,
                   motion_bucket_id=200, noise_aug_strength=0.02, generator=torch.Generator().manual_seed(0),
                   guess_mode=False, output_type="np")
        torch.cuda.synchronize()
        return out.frames

    one()  # warm-up (weight packing of the VAE / towers, CUDA-graph capture of the towers)
    t0 = time.perf_counter()
    frames = one()
    dt = time.perf_counter() - t0
    return {"value": round(FRAMES / dt, 4), "unit": "frames/s", "s_per_video": round(dt, 3),
            "api": "StableVideoDiffusionControlNetPipeline.__call__(PIL image, numpy condition, token ids) -> np frames",
            "frames_shape": list(frames[0].shape), "stages": "CLIP ViT-H + SD-2.1 text towers, VAE encode (1 + 14 "
            "images), 25 Euler steps UNet + GestureNet, VAE decode (chunks of 8), host read-back",
            "note": "extra; the headline metric and `e2e` cover the denoising loop only (latent mode)"}


# ====================================================================================================== eager bf16 on the same GPU
def eager_bf16_gpu(dev, unet, cn, sizes=((32, 48), (72, 128)), warm=2, iters=3):
    """The competitor the reference actually runs on a GPU (SURVEY.md §8d last bullet): its torch-eager
    path — cuDNN convolutions, cuBLAS linears, F.scaled_dot_product_attention — in bf16 on the SAME B200. diffusers is
    not installable, so the graph is the oracle's restatement of it (pinned to the reference's own forward code by
    tests/test_reference_pin.py) moved to CUDA bf16: one VGL Euler step (GestureNet + UNet, CFG pair, all the
    reference's per-step recomputation included). A reported baseline; nothing of it is on the product path."""
    from oracle import svd_oracle as O
    torch.backends.cudnn.benchmark = True
    usd = {k: v.detach().to(dev, torch.bfloat16) for k, v in unet.state_dict().items()}
    csd = {k: v.detach().to(dev, torch.bfloat16) for k, v in cn.state_dict().items()}
    cfg = dict(O.SVD_CONFIG)
    sig = O.karras_sigmas(NUM_STEPS)
    ts = O.euler_timesteps(sig).to(dev)
    out = {}
    for (h, w) in sizes:
        c = synth_inputs(h, w, 1, seed=0)
        lat = c["latents"].to(dev)
        img = c["image_latents"][:, None].repeat(1, FRAMES, 1, 1, 1).to(dev, torch.bfloat16)
        ehs = c["encoder_hidden_states"].to(dev, torch.bfloat16)
        ati = c["added_time_ids"].to(dev, torch.bfloat16)
        cc = torch.cat([c["controlnet_cond"][0]] * 2).to(dev, torch.bfloat16)
        gd = torch.linspace(1.0, 3.0, FRAMES, device=dev)[None, :, None, None, None]

        def step(i):
            s, sn = float(sig[i]), float(sig[i + 1])
            x = (torch.cat([lat] * 2) / (s * s + 1) ** 0.5).to(torch.bfloat16)
            x = torch.cat([x, img], dim=2)
            d, m = O.controlnet_forward(csd, cfg, x, ts[i], ehs, ati, cc, 1.0)
            eps = O.unet_forward(usd, cfg, x, ts[i], ehs, ati, d, m).float()
            eu, ec = eps.chunk(2)
            return O.euler_step(eu + gd * (ec - eu), lat, s, sn)

        try:
            with torch.no_grad():
                for i in range(warm):
                    step(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    step(warm + i)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[f"{h * 8}x{w * 8}"] = {"ms_per_euler_step": round(ms, 2),
                                       "frames_per_s_25_steps": round(FRAMES / (NUM_STEPS * ms / 1e3), 4)}
        except Exception as e:  # noqa: BLE001
            out[f"{h * 8}x{w * 8}"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        torch.cuda.empty_cache()
    out["what"] = ("torch-eager bf16 (cuDNN / cuBLAS / SDPA, cudnn.benchmark on) of the reference graph as restated by the "
                   "oracle, one VGL Euler step, CUDA events, same GPU and process as `value`")
    del usd, csd
    torch.cuda.empty_cache()
    return out


# ====================================================================================================== reference arm
def cpu_reference(args, sample_only: bool = False, steps: int = 1, warmup: int = 0):
    """The reference's CPU path = the fp32 oracle on all host threads. Bounded sample: ONE VGL Euler step (GestureNet +
    UNet forward, CFG pair) at 14x256x384 (latent 32x48, BASELINE.json configs[0] size), extrapolated to the bench
    workload by the algorithmic FLOP ratio (tools/flop_census.py)."""
    from oracle import svd_oracle as O
    from tools.flop_census import step_flops
    torch.set_num_threads(CPU_BASELINE_THREADS)
    h, w = args.height // 8, args.width // 8
    sh, sw = min(h, 32), min(w, 48)
    unet, cn = build_models("cpu")
    usd = {k: v.detach() for k, v in unet.state_dict().items()}
    csd = {k: v.detach() for k, v in cn.state_dict().items()}
    c = synth_inputs(sh, sw, 1, seed=0)
    cfg = dict(O.SVD_CONFIG)
    sig = O.karras_sigmas(NUM_STEPS)
    img = c["image_latents"][:, None].repeat(1, FRAMES, 1, 1, 1)
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            O.denoise_loop(usd, cfg, c["latents"], img, c["encoder_hidden_states"], c["added_time_ids"], NUM_STEPS, 1.0,
                           3.0, csd, cfg, c["controlnet_cond"][0], 1.0, max_steps=1)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    t_step = sum(times) / len(times)
    scale = step_flops(h, w, 2, True) / step_flops(sh, sw, 2, True)
    video_s = t_step * scale * NUM_STEPS
    return {"value": round(FRAMES / video_s, 6), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"1 VGL Euler step (GestureNet+UNet fp32, CFG pair) at 14x{sh * 8}x{sw * 8} = {t_step:.2f} s on "
                      f"{torch.get_num_threads()} threads; extrapolated x{scale:.2f} (algorithmic FLOPs) x25 steps to "
                      f"14x{args.height}x{args.width}",
            "sample_seconds": round(t_step, 3), "sample_tflops": round(step_flops(sh, sw, 2, True) / t_step / 1e12, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = cpu_reference(args, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1))
    line = {"impl": "reference",
            "metric": "frames/sec for 14-frame 576x1024 VGL, 25 Euler steps" if (args.height, args.width) == (576, 1024)
            else f"frames/sec for 14-frame {args.height}x{args.width} VGL, 25 Euler steps",
            "value": cpu["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(FRAMES / cpu["value"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same seeds as the sm_100a arm)",
            "config": {"workload": f"VGL (UNet+GestureNet) 25-step Euler, 14x{args.height}x{args.width}, CFG pair (B=2) per video",
                       "note": "reference CPU PyTorch path restated (oracle, diffusers not installable); bounded sample, extrapolated"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--videos", type=int, default=0,
                    help="videos per round (BASELINE.json configs[4]); default = one per GPU (weak scaling)")
    ap.add_argument("--vl", action="store_true",
                    help="VL instead of VGL: UNet only through StableVideoDiffusionPipeline (BASELINE config 2); N = 1, no "
                         "eager / CPU / full-pipeline legs")
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the torch-eager bf16 leg (same GPU, N = 1 only)")
    ap.add_argument("--quick-e2e", action="store_true", help="time 1 e2e video instead of >= 5 (development runs)")
    ap.add_argument("--no-full-pipeline", action="store_true", help="skip the image -> frames extra (N = 1 only)")
    ap.add_argument("--profile-only", action="store_true", help="run prepare + 2 Euler steps; cudaProfilerStart/Stop around the 2nd")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
