"""Data-parallel sharding of encode()/decode() over the GPUs of one node.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU tests).  Every rank sees the same
utterance list, plans the same global window list (windows.py), computes only its shard, and the outputs are
gathered — there is no collective inside the forward.  The decode pad length T' is part of the global plan, so
shards reproduce the single-GPU (= reference) result exactly.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import windows


def _gather_rows(local: torch.Tensor, counts: List[int], group) -> torch.Tensor:
    """all_gather of tensors (n_r, ...) with different n_r -> concatenation in rank order."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    nmax = max(counts) if counts else 0
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, counts)], dim=0)


class ShardedCodec:
    """encode()/decode() of an `AudioCodec`, window-sharded over a process group."""

    def __init__(self, model, group=None, encode_jobs: Optional[Callable] = None, decode_jobs: Optional[Callable] = None):
        self.model = model
        self.group = group
        self.encode_jobs = encode_jobs or model.encode_jobs
        self.decode_jobs = decode_jobs or model.decode_jobs

    @torch.inference_mode()
    def encode(self, wav_list, overlap_seconds=10, device=torch.device("cuda")):
        m = self.model
        device = torch.device(device)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lens = [int(len(w)) for w in wav_list]
        jobs = windows.plan_encode(lens, overlap_seconds, m.input_sample_rate, m.max_audio_seconds)
        if not jobs:
            return {"codes_list": [torch.zeros(m.num_groups, 0, device=device, dtype=torch.long) for _ in lens]}
        shards = windows.shard_round_robin(len(jobs), [j.n_valid for j in jobs], world)
        mine = self.encode_jobs(wav_list, [jobs[j] for j in shards[rank]], device)          # (8, n_r, 375)
        rows = _gather_rows(mine.permute(1, 0, 2).contiguous(), [len(s) for s in shards], self.group)   # (N, 8, 375)
        order = torch.tensor([j for s in shards for j in s], dtype=torch.int64, device=device)
        codes = torch.empty_like(rows)
        codes[order] = rows                                                                    # back to job order
        return {"codes_list": m.stitch_codes(codes.permute(1, 0, 2).contiguous(), lens, jobs, overlap_seconds)}

    @torch.inference_mode()
    def decode(self, codes_list, overlap_seconds=10, device=torch.device("cuda")):
        m = self.model
        device = torch.device(device)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lens = [int(c.shape[-1]) for c in codes_list]
        up = m.decoder_upsample_rate
        outs = [torch.zeros(L * up, dtype=torch.float32, device=device) for L in lens]
        groups = windows.plan_decode(lens, overlap_seconds, m.input_sample_rate, m.max_audio_seconds, m.encoder_downsample_rate)
        for pad_len, jobs in sorted(groups.items()):
            shards = windows.shard_round_robin(len(jobs), [j.n_valid for j in jobs], world)
            my_jobs = [jobs[j] for j in shards[rank]]
            if my_jobs:
                wav = self.decode_jobs(codes_list, my_jobs, device)                         # (n_r, 1280 T')
            else:
                wav = torch.zeros((0, up * pad_len), dtype=torch.float32, device=device)
            rows = _gather_rows(wav, [len(s) for s in shards], self.group)
            k = 0
            for s in shards:
                for j in s:
                    off, n = windows.decode_keep(jobs[j], overlap_seconds, m.input_sample_rate, m.max_audio_seconds,
                                                 m.encoder_downsample_rate, up)
                    outs[jobs[j].item][off:off + n] = rows[k, :n]
                    k += 1
        return {"syn_wav_list": outs}
