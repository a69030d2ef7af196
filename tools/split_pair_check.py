"""Split CFG pair over NCCL (world size 2): 1 video on 2 GPUs — rank 0 runs the uncond half, rank 1 the cond half, one
2-rank exchange of the noise prediction per Euler step (sharding.exchange_eps) — against the same video run whole on one
GPU. Prints the equality and the latency of both modes. Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/split_pair_check.py [--height 576 --width 1024] [--tiny]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--tiny", action="store_true")
    ap.add_argument("--steps", type=int, default=25)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from oracle_free_inputs import make
    from svd.scheduler import EulerDiscreteScheduler
    from svd.temporal_controlnet import ControlNetModel
    from svd.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
    from this_and_that_vdm_b200 import lib
    from this_and_that_vdm_b200.sampler import FusedDenoiser
    from this_and_that_vdm_b200.sharding import run_sharded
    lib.init(local)
    kind = dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4)) if args.tiny else \
        dict(block_out_channels=(320, 640, 1280, 1280), num_attention_heads=(5, 10, 20, 20))
    torch.manual_seed(1234)
    unet = UNetSpatioTemporalConditionModel(num_frames=14, **kind).eval()
    cn = ControlNetModel(**kind).eval()
    g = torch.Generator().manual_seed(1236)
    with torch.no_grad():
        for name, p in cn.named_parameters():
            if name.startswith("controlnet_") or name.startswith("conv_in_concat"):
                fan_in = p[0].numel() if p.ndim > 1 else p.numel()
                p.copy_(torch.randn(p.shape, generator=g) * (fan_in ** -0.5 if p.ndim > 1 else 0.1))
    unet.to(dev)
    cn.to(dev)
    h, w = args.height // 8, args.width // 8
    sched = EulerDiscreteScheduler()
    sched.set_timesteps(25)
    guidance = torch.linspace(1.0, 3.0, 14)
    cond = {k: v.to(dev) for k, v in make(h, w, 1, seed=0).items()} if rank == 0 else None
    den = FusedDenoiser(unet._get_engine(), cn._get_engine())

    def sharded():
        return run_sharded(1, cond, dev, lambda: den, sched.sigmas, sched.timesteps, guidance, max_steps=args.steps)

    def timed(fn, n=2):
        fn()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        dist.barrier()
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) / n

    split, t_split = timed(sharded)
    whole = t_whole = None
    if rank == 0:
        def one_gpu():
            c = cond
            st = c["latents"][0].clone().contiguous()
            den.prepare(c["encoder_hidden_states"], c["image_latents"], c["added_time_ids"], sched.sigmas, sched.timesteps,
                        guidance, num_frames=14, height=h, width=w, controlnet_cond=c["controlnet_cond"][0])
            for i in range(args.steps):
                den.step(i, st)
            torch.cuda.synchronize()
            return st
        one_gpu()
        t0 = time.perf_counter()
        whole = one_gpu()
        t_whole = time.perf_counter() - t0
    dist.barrier()
    if rank == 0:
        err = float((split[0] - whole).norm() / whole.norm())
        print("SPLIT_PAIR " + json.dumps({
            "config": "tiny" if args.tiny else "svd", "latent": [14, 4, h, w], "steps": args.steps, "world": world,
            "rel_l2_split_vs_whole": err, "seconds_split_2gpu": round(t_split, 4), "seconds_whole_1gpu": round(t_whole, 4),
            "speedup": round(t_whole / t_split, 3), "exchange_bytes_per_step": 14 * h * w * 4 * 4,
            "exchange": "sharding.exchange_eps (batch_isend_irecv, NCCL)"}), flush=True)
        assert err < 2e-2, err
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
