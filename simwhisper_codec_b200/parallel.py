"""Data-parallel sharding of encode()/decode() over the GPUs of one node.

One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU tests).  Every rank sees the same
utterance list, plans the same global window list (windows.py), computes only its shard, and the outputs are
gathered — there is no collective inside the forward.  The decode pad length T' is part of the global plan, so
shards reproduce the single-GPU (= reference) result exactly.

What travels:
  * encode: the codes of every window, 12 KB each — all-gathered, so every rank holds `codes_list` (decode() needs it
    on every rank, and it is four orders of magnitude less data than the audio);
  * decode: the KEPT samples of every window (at most 20 s = 1.3 MB, packed into one flat buffer per rank) go to ONE
    rank (`dst`, the caller's) with `dist.gather`, or nowhere (`gather_wav=False`: every rank keeps the windows it
    decoded, e.g. to write them to disk itself).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import windows


def _pad_rows(local: torch.Tensor, nmax: int) -> torch.Tensor:
    if local.shape[0] == nmax:
        return local.contiguous()
    pad = torch.empty((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    return pad


def _all_gather_rows(local: torch.Tensor, counts: List[int], group) -> List[torch.Tensor]:
    """tensors (n_r, ...) with different n_r -> per-rank list [(n_r, ...)] on every rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return [local]
    nmax = max(counts) if counts else 0
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)   # concatenated along dim 0
    dist.all_gather_into_tensor(out, _pad_rows(local, nmax), group=group)
    return [out[r * nmax: r * nmax + n] for r, n in enumerate(counts)]


class ShardedCodec:
    """encode()/decode() of an `AudioCodec`, window-sharded over a process group."""

    def __init__(self, model, group=None, encode_jobs: Optional[Callable] = None, decode_jobs: Optional[Callable] = None,
                 dst: int = 0):
        self.model = model
        self.group = group
        self.dst = dst
        self.encode_jobs = encode_jobs or model.encode_jobs
        self.decode_jobs = decode_jobs or model.decode_jobs

    @torch.inference_mode()
    def encode(self, wav_list, overlap_seconds=10, device=torch.device("cuda")):
        """-> {"codes_list": ...} on every rank (codes are small; decode() needs them everywhere)."""
        m = self.model
        device = torch.device(device)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lens = [int(len(w)) for w in wav_list]
        jobs = windows.plan_encode(lens, overlap_seconds, m.input_sample_rate, m.max_audio_seconds)
        if not jobs:
            return {"codes_list": [torch.zeros(m.num_groups, 0, device=device, dtype=torch.long) for _ in lens]}
        shards = windows.shard_round_robin(len(jobs), [j.n_valid for j in jobs], world)
        mine = self.encode_jobs(wav_list, [jobs[j] for j in shards[rank]], device)          # (8, n_r, 375)
        parts = _all_gather_rows(mine.permute(1, 0, 2), [len(s) for s in shards], self.group)   # [(n_r, 8, 375)]
        order = torch.tensor([j for s in shards for j in s], dtype=torch.int64).to(device, non_blocking=True)
        rows = torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]
        codes = torch.empty_like(rows)
        codes.index_copy_(0, order, rows)                                                      # back to job order
        return {"codes_list": m.stitch_codes(codes.permute(1, 0, 2).contiguous(), lens, jobs, overlap_seconds)}

    @torch.inference_mode()
    def decode(self, codes_list, overlap_seconds=10, device=torch.device("cuda"), gather_wav: bool = True):
        """-> {"syn_wav_list": ...}: the stitched waveforms on rank `dst` (None on the other ranks); with
        gather_wav=False nothing is exchanged and every rank gets {"local": [(item, offset, samples tensor), ...]}."""
        m = self.model
        device = torch.device(device)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lens = [int(c.shape[-1]) for c in codes_list]
        up = m.decoder_upsample_rate
        groups = windows.plan_decode(lens, overlap_seconds, m.input_sample_rate, m.max_audio_seconds, m.encoder_downsample_rate)
        keep_args = (overlap_seconds, m.input_sample_rate, m.max_audio_seconds, m.encoder_downsample_rate, up)
        on_dst = rank == self.dst
        outs = None
        if gather_wav and on_dst:            # one flat buffer, one view per item: every sample is written by exactly one window
            flat = torch.empty(sum(lens) * up, dtype=torch.float32, device=device)
            outs = list(torch.split(flat, [L * up for L in lens]))
        local, dst_views, src_views = [], [], []
        order = sorted(groups.items())
        # the groups (one launch set per pad length) are sharded jointly: a small group goes to few ranks (windows.shard_groups)
        if os.environ.get("SWC_SHARD_JOINT", "1") != "0":
            all_shards = windows.shard_groups([[j.n_valid for j in jobs] for _, jobs in order], world)
        else:                                # every group dealt to every rank (A/B switch for measurements)
            all_shards = [windows.shard_round_robin(len(jobs), [j.n_valid for j in jobs], world) for _, jobs in order]
        # every group's launch set first, ONE gather at the end: a collective between the groups would make the ranks that
        # have no window in a group wait (on the device) for the ranks that do
        wavs = []
        for (pad_len, jobs), shards in zip(order, all_shards):
            my_jobs = [jobs[j] for j in shards[rank]]
            wav = self.decode_jobs(codes_list, my_jobs, device) if my_jobs else None        # (n_r, 1280 T')
            wavs.append(wav)
            if not gather_wav:
                for k, j in enumerate(my_jobs):
                    off, n = windows.decode_keep(j, *keep_args)
                    local.append((j.item, off, wav[k, :n]))
        if gather_wav:
            # only the samples that are kept travel: every rank packs the kept part of all its windows into one flat buffer
            # (all ranks know all jobs, hence every rank's total), rank dst receives world buffers of the largest total
            keeps = [[[windows.decode_keep(jobs[j], *keep_args) for j in s] for s in shards]
                     for (_, jobs), shards in zip(order, all_shards)]                       # [group][rank][window] -> (offset, n)
            totals = [sum(n for g in keeps for _, n in g[r]) for r in range(world)]
            send = torch.empty(max(totals + [1]), dtype=torch.float32, device=device)
            mine_n = [n for g in keeps for _, n in g[rank]]
            if mine_n:
                torch._foreach_copy_(list(torch.split(send[: totals[rank]], mine_n)),
                                     [wav[k, :n] for wav, g in zip(wavs, keeps) for k, (_, n) in enumerate(g[rank])])
            if world == 1:
                bufs = [send]
            else:
                bufs = [torch.empty_like(send) for _ in range(world)] if on_dst else None
                dist.gather(send, bufs, dst=dist.get_global_rank(self.group, self.dst) if self.group is not None else self.dst,
                            group=self.group)
            if on_dst:
                for r in range(world):
                    pos = 0
                    for ((_, jobs), shards), g in zip(zip(order, all_shards), keeps):
                        for j, (off, n) in zip(shards[r], g[r]):
                            dst_views.append(outs[jobs[j].item][off:off + n])
                            src_views.append(bufs[r][pos:pos + n])
                            pos += n
        if not gather_wav:
            return {"local": local}
        if on_dst and dst_views:
            torch._foreach_copy_(dst_views, src_views)                                          # one fused launch for all windows
        return {"syn_wav_list": outs}
