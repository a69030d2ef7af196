"""Multi-GPU sharding of the CFG-doubled video batch: one process per GPU, torch.distributed.

The unit of independence is one *sequence* ([F, 8, h, w]); a video is a CFG pair of two sequences. No layer mixes
sequences except (a) the CFG combine once per step and (b) the temporal-context quirk, which only needs every rank
to hold both CONTEXTS of its pair (they travel in the conditioning pack). Policy for N videos on G ranks:

    N >= G          whole pairs per rank, round-robin; zero per-step communication
    G/2 <= N < G    the first (G - N) videos are split (rank 2k: uncond half, rank 2k+1: cond half) and exchange
                    their noise predictions once per step inside a 2-rank group; the rest run whole
    N <  G/2        every video is split; ranks >= 2N idle ("replicas only" beyond 2N — there is no more
                    independent work in one video: frames and pixels must never be sharded, §8e)

Collectives per sample: ONE broadcast of the conditioning pack from rank 0 and ONE gather of the final latents to
rank 0. Weights are replicated.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Assignment:
    video: int
    batch_offset: int  # 0: starts at the uncond half, 1: cond half only
    b_local: int       # 2: whole pair on this rank, 1: one half
    partner: int       # rank holding the other half (-1 if whole)


def plan(n_videos: int, world: int) -> List[List[Assignment]]:
    """Returns, per rank, the ordered list of assignments."""
    if n_videos <= 0 or world <= 0:
        raise ValueError("n_videos and world must be positive")
    per_rank: List[List[Assignment]] = [[] for _ in range(world)]
    if n_videos >= world:
        for v in range(n_videos):
            per_rank[v % world].append(Assignment(v, 0, 2, -1))
        return per_rank
    n_split = min(n_videos, world - n_videos)  # each split video uses one extra rank
    rank = 0
    for v in range(n_videos):
        if v < n_split and rank + 1 < world:
            per_rank[rank].append(Assignment(v, 0, 1, rank + 1))
            per_rank[rank + 1].append(Assignment(v, 1, 1, rank))
            rank += 2
        else:
            per_rank[rank].append(Assignment(v, 0, 2, -1))
            rank += 1
    return per_rank


PACK_KEYS = ("encoder_hidden_states", "image_latents", "added_time_ids", "controlnet_cond", "latents")


def pack_conditioning(cond: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, List[Tuple[str, Tuple[int, ...]]]]:
    """Flattens the per-sample conditioning tensors into ONE fp32 buffer (=> one broadcast)."""
    meta, flat = [], []
    for k in PACK_KEYS:
        if k in cond and cond[k] is not None:
            t = cond[k].detach().to(torch.float32).contiguous()
            meta.append((k, tuple(t.shape)))
            flat.append(t.reshape(-1))
    return torch.cat(flat), meta


def unpack_conditioning(buf: torch.Tensor, meta: Sequence[Tuple[str, Tuple[int, ...]]]) -> Dict[str, torch.Tensor]:
    out, off = {}, 0
    for k, shape in meta:
        n = 1
        for d in shape:
            n *= d
        out[k] = buf[off:off + n].view(*shape)
        off += n
    return out


def broadcast_conditioning(cond: Optional[Dict[str, torch.Tensor]], device, src: int = 0, group=None):
    """ONE broadcast of the conditioning pack (+ a tiny object broadcast of its shape table)."""
    rank = dist.get_rank()
    meta = [None]
    buf = None
    if rank == src:
        buf, m = pack_conditioning(cond)
        buf = buf.to(device)
        meta = [m]
    dist.broadcast_object_list(meta, src=src, group=group)
    if rank != src:
        n = sum(int(torch.tensor(s).prod()) if len(s) else 1 for _, s in meta[0])
        buf = torch.empty(n, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src, group=group)
    return unpack_conditioning(buf, meta[0])


def gather_latents(local: torch.Tensor, dst: int = 0, group=None) -> Optional[List[torch.Tensor]]:
    """ONE gather of the final latents ([n_local, F, 4, h, w], padded to the same n_local on every rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank()
    outs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local.contiguous(), outs, dst=dst, group=group)
    return outs


def exchange_eps(eps_local: torch.Tensor, my_offset: int, partner: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Split-pair per-step exchange: returns (eps_uncond, eps_cond). Both ranks end with identical tensors."""
    other = torch.empty_like(eps_local)
    rank = dist.get_rank()
    ops = [dist.P2POp(dist.isend, eps_local, partner), dist.P2POp(dist.irecv, other, partner)]
    if rank > partner:
        ops.reverse()
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    return (eps_local, other) if my_offset == 0 else (other, eps_local)


def run_sharded(n_videos: int, cond: Optional[Dict[str, torch.Tensor]], device,
                make_denoiser: Callable[[], "object"], sigmas: torch.Tensor, timesteps: torch.Tensor,
                guidance: torch.Tensor, *, vgl: bool = True, conditioning_scale: float = 1.0,
                max_steps: Optional[int] = None,
                decode: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> Optional[torch.Tensor]:
    """Denoises `n_videos` videos across the ranks of the default process group.

    `cond` (rank 0 only): encoder_hidden_states [2N, L, D] (uncond rows first), image_latents [2N, 4, h, w],
    added_time_ids [2N, 3], controlnet_cond [N, F, 4, h, w] (VGL), latents [N, F, 4, h, w] (x init_noise_sigma).
    Returns on rank 0 the final latents [N, F, 4, h, w]; None elsewhere.

    `decode` (optional): latents [F, 4, h, w] -> decoded video (any fixed shape, e.g. the VAE engine's
    `decode_latents` giving [3, F, H, W]). Every rank then decodes the videos it finished (the uncond-half rank of a
    split pair; its partner contributes zeros) and the ONE gather carries the decoded videos instead of the latents,
    so the VAE work is spread over the ranks as well; rank 0 gets [N, *decoded shape]."""
    rank, world = dist.get_rank(), dist.get_world_size()
    c = broadcast_conditioning(cond, device)
    N = n_videos
    Fr, h, w = c["latents"].shape[1], c["latents"].shape[-2], c["latents"].shape[-1]
    mine = plan(N, world)[rank]
    n_slots = max(len(a) for a in plan(N, world))
    den = make_denoiser()
    results = torch.zeros(max(n_slots, 1), Fr, 4, h, w, dtype=torch.float32, device=device)
    n_steps = len(timesteps) if max_steps is None else min(max_steps, len(timesteps))
    decoded: Optional[torch.Tensor] = None
    for slot, a in enumerate(mine):
        idx = [a.video, N + a.video]
        state = c["latents"][a.video].clone().contiguous()
        den.prepare(c["encoder_hidden_states"][idx], c["image_latents"][idx], c["added_time_ids"][idx], sigmas,
                    timesteps, guidance, num_frames=Fr, height=h, width=w,
                    controlnet_cond=c["controlnet_cond"][a.video] if vgl else None,
                    conditioning_scale=conditioning_scale, batch_offset=a.batch_offset, b_local=a.b_local)
        rows = Fr * h * w
        for i in range(n_steps):
            eps = den.predict(i, state)
            if a.b_local == 2:
                eu, ec = eps[:rows], eps[rows:]
            else:
                eu, ec = exchange_eps(eps, a.batch_offset, a.partner)
            den.euler_update(i, state, eu, ec)
        if decode is None:
            results[slot] = state
        elif a.batch_offset == 0:
            frames = decode(state).to(torch.float32)
            if decoded is None:
                decoded = torch.zeros(max(n_slots, 1), *frames.shape, dtype=torch.float32, device=device)
            decoded[slot] = frames
    if decode is not None:
        # ranks that decoded nothing (idle, or only cond halves) learn the decoded shape before the gather
        dims = list(decoded.shape[1:]) if decoded is not None else []
        shape = torch.tensor(dims + [0] * (8 - len(dims)), dtype=torch.int64, device=device)
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
        full = tuple(int(v) for v in shape.tolist() if v > 0)
        results = decoded if decoded is not None else torch.zeros(max(n_slots, 1), *full, dtype=torch.float32,
                                                                  device=device)
    outs = gather_latents(results)
    if rank != 0:
        return None
    final = torch.zeros(N, *results.shape[1:], dtype=torch.float32, device=device)
    pl = plan(N, world)
    for r in range(world):
        for slot, a in enumerate(pl[r]):
            if a.batch_offset == 0:  # both halves hold identical states; take the first
                final[a.video] = outs[r][slot]
    return final
