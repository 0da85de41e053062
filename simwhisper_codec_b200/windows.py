"""Host-side window planning and stitching for `AudioCodec.encode()/decode()`.

Restates the integer index math of the reference's chunk loops (audiocodec/model.py:254-302, 320-367) as a flat
job list, so that all (item, window) pairs can run as ONE batch (and be sharded across GPUs) instead of a
Python loop with host synchronisations:

* encode: window c of an item starts at c*hop samples (hop = (30-overlap) s), holds <= 30 s, and contributes its
  first `keep` = hop/1280 codes ("keep-first / overlap-discard").
* decode: window c starts at c*keep codes; its pad length T'_c = min(c*keep + 375, max_len) - c*keep is a property
  of the BATCH (the longest item), because the reference decodes un-padded batches whose un-masked convolutions
  see the zero padding.  T' must therefore be computed before any sharding.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class EncodeJob:
    item: int
    start: int        # first sample of the window
    n_valid: int      # valid samples in the window (<= 480000)


@dataclass(frozen=True)
class DecodeJob:
    item: int
    chunk: int        # window index c
    start: int        # first code frame of the window
    n_valid: int      # valid code frames in the window
    pad_len: int      # T' of this window index (batch-global)


def plan_encode(lens: Sequence[int], overlap_seconds: int = 10, sr: int = 16000, max_seconds: int = 30) -> List[EncodeJob]:
    win = max_seconds * sr
    hop = (max_seconds - overlap_seconds) * sr
    jobs: List[EncodeJob] = []
    if hop <= 0:
        return jobs
    for i, L in enumerate(lens):
        for c in range((L + hop - 1) // hop):
            jobs.append(EncodeJob(i, c * hop, min(L - c * hop, win)))
    return jobs


def code_frames(n_samples: int) -> int:
    """Valid code frames of a window with n_samples valid samples: mel ceil(n/160) -> encoder //2 -> down-sampler ceil(/4)
    (reference feature_extractor.py:221-226, modules.py:322, 549)."""
    return (((n_samples + 159) // 160) // 2 + 3) // 4


def encode_gather_index(lens: Sequence[int], jobs: Sequence[EncodeJob], overlap_seconds: int = 10, sr: int = 16000,
                        max_seconds: int = 30, rate: int = 1280, frames_per_window: int = 375) -> Tuple[np.ndarray, List[int]]:
    """Flat source index (job * 375 + frame; int64 array) of every output code position, and the per-item output lengths.

    The reference concatenates `keep` code columns per window index of the BATCH (model.py:271-302): output position p of
    item i is frame p % keep of its window p // keep if the item owns that window and the frame is below the window's valid
    code length, else zero (source index -1); an item keeps min(len // rate, max_chunks * keep) positions.  With the default
    overlap (hop a multiple of `rate`) every position has a source; with e.g. overlap 5 s (hop 400 000 samples, keep 312) the
    reference's time axis drifts and the trailing positions are zeros or cut - reproduced here as it is."""
    hop = (max_seconds - overlap_seconds) * sr
    keep = hop // rate
    if keep <= 0 or not lens:
        return np.zeros(0, np.int64), [0 for _ in lens]
    max_chunks = (max(lens) + hop - 1) // hop
    owned: Dict[Tuple[int, int], Tuple[int, int]] = {}     # (item, window index) -> (job index, valid code frames kept)
    for j, job in enumerate(jobs):
        owned[(job.item, job.start // hop)] = (j, min(code_frames(job.n_valid), keep))
    splits: List[int] = []
    seg_len, seg_job, seg_valid = [], [], []                # one segment per (item, window index) of the output
    for i, L in enumerate(lens):
        n = min(L // rate, max_chunks * keep)
        splits.append(n)
        for c in range((n + keep - 1) // keep):
            j, valid = owned.get((i, c), (0, 0))
            seg_len.append(min(keep, n - c * keep))
            seg_job.append(j)
            seg_valid.append(valid)
    seg_len = np.asarray(seg_len, np.int64)
    total = int(seg_len.sum())
    frame = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(seg_len) - seg_len, seg_len)
    src = np.repeat(np.asarray(seg_job, np.int64), seg_len) * frames_per_window + frame
    src[frame >= np.repeat(np.asarray(seg_valid, np.int64), seg_len)] = -1
    return src, splits


def plan_decode(code_lens: Sequence[int], overlap_seconds: int = 10, sr: int = 16000, max_seconds: int = 30,
                rate: int = 1280) -> Dict[int, List[DecodeJob]]:
    """Jobs grouped by pad length T' (one detokenize batch per distinct T')."""
    win = max_seconds * sr // rate
    keep = (max_seconds - overlap_seconds) * sr // rate
    groups: Dict[int, List[DecodeJob]] = {}
    if keep <= 0 or not code_lens:
        return groups
    maxlen = max(code_lens)
    for c in range((maxlen + keep - 1) // keep):
        pad_len = min(c * keep + win, maxlen) - c * keep
        for i, L in enumerate(code_lens):
            n = min(max(L - c * keep, 0), pad_len)
            if n > 0:
                groups.setdefault(pad_len, []).append(DecodeJob(i, c, c * keep, n, pad_len))
    return groups


def decode_keep(job: DecodeJob, overlap_seconds: int = 10, sr: int = 16000, max_seconds: int = 30, rate: int = 1280,
                up: int = 1280) -> Tuple[int, int]:
    """(output offset in samples, number of samples kept) of one decode window."""
    keep = (max_seconds - overlap_seconds) * sr // rate
    return job.start * up, min(job.n_valid, keep) * up


def shard_round_robin(n_jobs: int, cost: Sequence[int], world: int, load: Sequence[int] = None,
                      ranks: Sequence[int] = None) -> List[List[int]]:
    """Deal jobs to ranks longest-first (greedy by remaining load).  Deterministic on every rank.
    load: work the ranks already carry (updated in place when a list is passed); ranks: deal only to these ranks."""
    order = sorted(range(n_jobs), key=lambda j: (-cost[j], j))
    load = [0] * world if load is None else load
    ranks = list(range(world)) if ranks is None else list(ranks)
    out: List[List[int]] = [[] for _ in range(world)]
    for j in order:
        r = min(ranks, key=lambda k: (load[k], k))
        out[r].append(j)
        load[r] += max(cost[j], 1)
    for r in range(world):
        out[r].sort()
    return out


def shard_groups(group_costs: Sequence[Sequence[int]], world: int, per_job: int = 10, per_set: int = 400) -> List[List[List[int]]]:
    """Shard several job groups that run as separate launch sets (decode windows of different pad length T') jointly:
    a group whose whole cost is a fraction of one rank's fair share is not scattered over all ranks (every rank would pay
    a launch set of a few windows at poor GPU efficiency) but dealt to the k least-loaded ranks, k = its cost in fair
    shares rounded up; the larger groups then fill the ranks up.  Cost of a job = its valid code frames + per_job (the
    Vocos halo, 80 frames = 10 code frames); a rank that takes part in a group pays per_set on top (the ~440 launches of
    a set, measured ~1.5 ms = 400 code frames of work).  Returns, per group, the per-rank job index lists.
    Deterministic on every rank; any split gives the same results (windows are independent)."""
    costs = [[max(c, 0) + per_job for c in g] for g in group_costs]
    total = sum(sum(g) for g in costs)
    fair = max(1.0, total / max(world, 1) + per_set)
    load = [0] * world
    out: List[List[List[int]]] = [None] * len(costs)
    for g in sorted(range(len(costs)), key=lambda g: (sum(costs[g]), g)):      # smallest group first
        k = max(1, -(-sum(costs[g]) // int(fair)))
        if 2 * k > world:                      # a group of more than half the job goes to every rank
            k = world
        ranks = sorted(range(world), key=lambda r: (load[r], r))[:k]
        if costs[g]:
            for r in ranks:
                load[r] += per_set
        out[g] = shard_round_robin(len(costs[g]), costs[g], world, load, ranks)
    return out
