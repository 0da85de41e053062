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


def encode_gather_index(lens: Sequence[int], jobs: Sequence[EncodeJob], overlap_seconds: int = 10, sr: int = 16000,
                        max_seconds: int = 30, rate: int = 1280, frames_per_window: int = 375) -> Tuple[List[int], List[int]]:
    """Flat source index (job * 375 + frame) of every output code position, and the per-item output lengths.
    Output position p of item i comes from its window p // keep, frame p % keep; items keep len // rate codes."""
    keep = (max_seconds - overlap_seconds) * sr // rate
    first: Dict[int, int] = {}
    for j, job in enumerate(jobs):
        first.setdefault(job.item, j)
    src: List[int] = []
    splits: List[int] = []
    for i, L in enumerate(lens):
        n = L // rate
        splits.append(n)
        base = first.get(i, 0)
        src.extend((base + p // keep) * frames_per_window + p % keep for p in range(n))
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


def shard_round_robin(n_jobs: int, cost: Sequence[int], world: int) -> List[List[int]]:
    """Deal jobs to ranks longest-first (greedy by remaining load).  Deterministic on every rank."""
    order = sorted(range(n_jobs), key=lambda j: (-cost[j], j))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for j in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(j)
        load[r] += max(cost[j], 1)
    for r in range(world):
        out[r].sort()
    return out
