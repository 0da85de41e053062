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
        all_shards = windows.shard_groups([[j.n_valid for j in jobs] for _, jobs in order], world)
        for (pad_len, jobs), shards in zip(order, all_shards):
            my_jobs = [jobs[j] for j in shards[rank]]
            if my_jobs:
                wav = self.decode_jobs(codes_list, my_jobs, device)                         # (n_r, 1280 T')
            else:
                wav = torch.empty((0, up * pad_len), dtype=torch.float32, device=device)
            if not gather_wav:
                for k, j in enumerate(my_jobs):
                    off, n = windows.decode_keep(j, *keep_args)
                    local.append((j.item, off, wav[k, :n]))
                continue
            # only the samples that are kept travel: every rank packs the kept part of its windows into one flat buffer
            # (all ranks know all jobs, hence every rank's total), rank dst receives world buffers of the largest total
            keeps = [[windows.decode_keep(jobs[j], *keep_args) for j in s] for s in shards]
            totals = [sum(n for _, n in ks) for ks in keeps]
            send = torch.empty(max(totals), dtype=torch.float32, device=device)
            if my_jobs:
                torch._foreach_copy_(list(torch.split(send[: totals[rank]], [n for _, n in keeps[rank]])),
                                     [wav[k, :n] for k, (_, n) in enumerate(keeps[rank])])
            if world == 1:
                bufs = [send]
            else:
                bufs = [torch.empty_like(send) for _ in range(world)] if on_dst else None
                dist.gather(send, bufs, dst=dist.get_global_rank(self.group, self.dst) if self.group is not None else self.dst,
                            group=self.group)
            if not on_dst:
                continue
            for r, s in enumerate(shards):
                pos = 0
                for j, (off, n) in zip(s, keeps[r]):
                    dst_views.append(outs[jobs[j].item][off:off + n])
                    src_views.append(bufs[r][pos:pos + n])
                    pos += n
        if not gather_wav:
            return {"local": local}
        if on_dst and dst_views:
            torch._foreach_copy_(dst_views, src_views)                                          # one fused launch for all windows
        return {"syn_wav_list": outs}
