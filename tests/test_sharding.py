"""Multi-GPU host logic on CPU: the sharding plan, the one-broadcast / one-gather protocol and the split-pair
exchange, exercised with world_size 2 over gloo and a fake denoiser (no CUDA)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from this_and_that_vdm_b200.sharding import pack_conditioning, plan, unpack_conditioning


def test_plan_policies():
    for n in range(1, 12):
        for g in (1, 2, 4, 8):
            p = plan(n, g)
            assert len(p) == g
            seen = {}
            for r, assigns in enumerate(p):
                for a in assigns:
                    seen.setdefault(a.video, []).append((r, a))
            assert sorted(seen) == list(range(n))
            for v, lst in seen.items():
                if len(lst) == 1:
                    assert lst[0][1].b_local == 2 and lst[0][1].partner == -1
                else:
                    (r0, a0), (r1, a1) = lst
                    assert {a0.batch_offset, a1.batch_offset} == {0, 1} and a0.b_local == a1.b_local == 1
                    assert a0.partner == r1 and a1.partner == r0
            if n >= g:
                assert max(len(a) for a in p) - min(len(a) for a in p) <= 1  # balanced whole pairs
            if n * 2 <= g:
                assert all(len(lst) == 2 for lst in seen.values())            # every video split
    with pytest.raises(ValueError):
        plan(0, 2)


def test_pack_roundtrip():
    c = {"encoder_hidden_states": torch.randn(4, 78, 16), "image_latents": torch.randn(4, 4, 8, 8),
         "added_time_ids": torch.randn(4, 3), "controlnet_cond": torch.randn(2, 14, 4, 8, 8),
         "latents": torch.randn(2, 14, 4, 8, 8)}
    buf, meta = pack_conditioning(c)
    out = unpack_conditioning(buf, meta)
    assert all(torch.equal(out[k], c[k]) for k in c)


class FakeDenoiser:
    """Stands in for FusedDenoiser: eps of a half = (half_id + 1) * mean(latents) * step-dependent factor."""

    def prepare(self, ehs, img, ids, sigmas, timesteps, guidance, *, num_frames, height, width, controlnet_cond,
                conditioning_scale, batch_offset, b_local):
        self.off, self.bl, self.F, self.h, self.w = batch_offset, b_local, num_frames, height, width
        self.guidance = guidance
        self.sig = [float(s) for s in sigmas]
        self.bias = float(ehs.sum() * 0 + img[1].mean()) + float(controlnet_cond.mean())

    def predict(self, i, state):
        rows = self.F * self.h * self.w
        halves = [self.off] if self.bl == 1 else [0, 1]
        x = state.permute(0, 2, 3, 1).reshape(rows, 4)
        return torch.cat([(hf + 1) * 0.1 * x + self.bias * (i + 1) for hf in halves])

    def euler_update(self, i, state, eu, ec):
        g = self.guidance.repeat_interleave(self.h * self.w)[:, None]
        eps = (eu + g * (ec - eu)).reshape(self.F, self.h, self.w, 4).permute(0, 3, 1, 2)
        s, sn = self.sig[i], self.sig[i + 1]
        x0 = eps * (-s / (s * s + 1) ** 0.5) + state / (s * s + 1)
        state += (state - x0) / s * (sn - s)


def _cond(n):
    g = torch.Generator().manual_seed(3)
    return {"encoder_hidden_states": torch.randn(2 * n, 5, 8, generator=g), "image_latents": torch.randn(2 * n, 4, 4, 4, generator=g),
            "added_time_ids": torch.randn(2 * n, 3, generator=g), "controlnet_cond": torch.randn(n, 3, 4, 4, 4, generator=g),
            "latents": torch.randn(n, 3, 4, 4, 4, generator=g) * 10}


def _fake_decode(state):
    """Stands in for the VAE: [F, 4, h, w] -> [3, F, 2h, 2w]."""
    up = state[:, :3].repeat_interleave(2, -1).repeat_interleave(2, -2) * 0.5 + 1.0
    return up.permute(1, 0, 2, 3).contiguous()


def _worker(rank, world, port, n_videos, out_path, with_decode=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from this_and_that_vdm_b200.sharding import run_sharded
    sig = torch.tensor([5.0, 3.0, 1.0, 0.0])
    ts = torch.tensor([0.4, 0.2, 0.0])
    res = run_sharded(n_videos, _cond(n_videos) if rank == 0 else None, torch.device("cpu"), FakeDenoiser, sig, ts,
                      torch.linspace(1, 3, 3), vgl=True, decode=_fake_decode if with_decode else None)
    if rank == 0:
        torch.save(res, out_path)
    else:
        assert res is None
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_videos", [1, 2, 3])
def test_sharded_equals_single_rank(tmp_path, n_videos):
    """world 2 (split pair for N=1, whole pairs for N>=2) must reproduce the single-rank result bit for bit."""
    from this_and_that_vdm_b200 import sharding
    out = tmp_path / "r.pt"
    mp.spawn(_worker, args=(2, _free_port(), n_videos, str(out)), nprocs=2, join=True)
    got = torch.load(out)
    # single-rank reference with the same fake denoiser
    c = _cond(n_videos)
    sig = torch.tensor([5.0, 3.0, 1.0, 0.0])
    ref = []
    for v in range(n_videos):
        d = FakeDenoiser()
        idx = [v, n_videos + v]
        d.prepare(c["encoder_hidden_states"][idx], c["image_latents"][idx], c["added_time_ids"][idx], sig, None,
                  torch.linspace(1, 3, 3), num_frames=3, height=4, width=4, controlnet_cond=c["controlnet_cond"][v],
                  conditioning_scale=1.0, batch_offset=0, b_local=2)
        st = c["latents"][v].clone()
        for i in range(3):
            e = d.predict(i, st)
            d.euler_update(i, st, e[:48], e[48:])
        ref.append(st)
    assert torch.equal(got, torch.stack(ref))


@pytest.mark.parametrize("n_videos", [1, 3])
def test_sharded_decode_on_every_rank(tmp_path, n_videos):
    """With `decode=` each rank decodes the videos it finished and the one gather carries the decoded videos: N = 1 on 2
    ranks is a split pair (the cond-half rank decodes nothing and must still join the gather with the right shape)."""
    lat_path, dec_path = tmp_path / "lat.pt", tmp_path / "dec.pt"
    mp.spawn(_worker, args=(2, _free_port(), n_videos, str(lat_path)), nprocs=2, join=True)
    mp.spawn(_worker, args=(2, _free_port(), n_videos, str(dec_path), True), nprocs=2, join=True)
    lat, dec = torch.load(lat_path), torch.load(dec_path)
    assert dec.shape == (n_videos, 3, 3, 8, 8)
    assert torch.equal(dec, torch.stack([_fake_decode(x) for x in lat]))
