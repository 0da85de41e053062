"""Host logic: window planning/stitching equals the reference's chunk loops (via the oracle port), and the
window-sharded multi-process path (gloo, world_size 2) reproduces the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simwhisper_codec_b200 import windows


class FakeCodec:
    """Stand-in compute with the AudioCodec window interface: codes/wavs are cheap deterministic functions of the
    window content, so stitching and sharding mistakes show up as value mismatches."""
    num_groups, input_sample_rate, max_audio_seconds = 8, 16000, 30
    encoder_downsample_rate = decoder_upsample_rate = 1280

    def encode_jobs(self, wav_list, jobs, device):
        out = torch.zeros(8, len(jobs), 375, dtype=torch.int32)
        for k, j in enumerate(jobs):
            w = wav_list[j.item][j.start:j.start + j.n_valid]
            n = (-(-(-(-j.n_valid // 160)) // 2) + 3) // 4 if False else (((j.n_valid + 159) // 160) // 2 + 3) // 4
            fr = w[: n * 1280].reshape(n, 1280) if n * 1280 <= w.numel() else torch.nn.functional.pad(w, (0, n * 1280 - w.numel())).reshape(n, 1280)
            val = (fr.abs().sum(1) * 1000).long() % 2016
            for g in range(8):
                out[g, k, :n] = ((val + 37 * g + 11 * j.n_valid) % 2016).int()
        return out

    def stitch_codes(self, codes, lens, jobs, overlap_seconds):
        src, splits = windows.encode_gather_index(lens, jobs, overlap_seconds)
        flat = torch.cat([codes.reshape(8, -1), torch.zeros(8, 1, dtype=codes.dtype)], dim=1)      # last column: the zero source
        src = torch.from_numpy(src)
        src = torch.where(src < 0, torch.full_like(src, flat.shape[1] - 1), src)
        return list(torch.split(flat.index_select(1, src), splits, dim=1))

    def decode_jobs(self, codes_list, jobs, device):
        Tp = jobs[0].pad_len
        out = torch.zeros(len(jobs), 1280 * Tp)
        for k, j in enumerate(jobs):
            c = codes_list[j.item][:, j.start:j.start + j.n_valid].float().sum(0)          # (n,)
            out[k, : j.n_valid * 1280] = (c[:, None] + torch.arange(1280)[None, :] * 1e-3 + Tp).reshape(-1)
        return out


def _reference_style_encode(fake, wavs, overlap=10):
    """the reference's chunk loop (model.py:254-302) around the same fake compute"""
    hop, win, keep = (30 - overlap) * 16000, 480000, (30 - overlap) * 16000 // 1280
    L = [len(w) for w in wavs]
    maxlen = max(L)
    pieces = []
    for c in range((maxlen + hop - 1) // hop):
        s, e = c * hop, min(c * hop + win, maxlen)
        jobs = [windows.EncodeJob(i, s, max(0, min(L[i] - s, e - s))) for i in range(len(wavs))]
        codes = fake.encode_jobs(wavs, jobs, "cpu")
        piece = torch.zeros(8, len(wavs), keep, dtype=torch.int32)
        for b, j in enumerate(jobs):
            cl = (((j.n_valid + 159) // 160) // 2 + 3) // 4
            v = min(cl, keep)
            piece[:, b, :v] = codes[:, b, :v]
        pieces.append(piece)
    allc = torch.cat(pieces, -1)
    return [allc[:, i, : L[i] // 1280] for i in range(len(wavs))]


def test_encode_plan_matches_reference_loop():
    g = torch.Generator().manual_seed(0)
    lens = [48123, 800000, 365000, 1280 * 250, 479999, 1279]
    wavs = [torch.randn(n, generator=g) for n in lens]
    fake = FakeCodec()
    jobs = windows.plan_encode(lens)
    assert [sum(1 for j in jobs if j.item == i) for i in range(len(lens))] == [1, 3, 2, 1, 2, 1]
    codes = fake.encode_jobs(wavs, jobs, "cpu")
    ours = fake.stitch_codes(codes, lens, jobs, 10)
    ref = _reference_style_encode(fake, wavs)
    assert [tuple(c.shape) for c in ours] == [(8, n // 1280) for n in lens]
    for a, b in zip(ours, ref):
        assert torch.equal(a, b)


@pytest.mark.parametrize("overlap", [0, 5, 10, 25])
def test_encode_plan_matches_reference_loop_any_overlap(overlap):
    """hop not a multiple of 1280 (overlap 5 -> keep 312, overlap 25 -> keep 62): the reference's code axis drifts against
    the sample axis, trailing positions are zero or cut (model.py:271-302) - the flat plan reproduces exactly that."""
    g = torch.Generator().manual_seed(overlap)
    for lens in ([1200000], [48123, 800000, 365000, 1280 * 250, 479999, 1279], [400000, 400001, 399999, 1600000]):
        wavs = [torch.randn(n, generator=g) for n in lens]
        fake = FakeCodec()
        jobs = windows.plan_encode(lens, overlap)
        ours = fake.stitch_codes(fake.encode_jobs(wavs, jobs, "cpu"), lens, jobs, overlap)
        ref = _reference_style_encode(fake, wavs, overlap)
        assert [tuple(c.shape) for c in ours] == [tuple(c.shape) for c in ref]
        for a, b in zip(ours, ref):
            assert torch.equal(a, b)


def _reference_style_decode(fake, codes_list, overlap):
    """the reference's decode chunk loop (model.py:320-367) around the same fake compute"""
    keep, win = (30 - overlap) * 16000 // 1280, 375
    L = [c.shape[-1] for c in codes_list]
    maxlen = max(L)
    pieces = []
    for c in range((maxlen + keep - 1) // keep):
        s, e = c * keep, min(c * keep + win, maxlen)
        jobs = [windows.DecodeJob(i, c, s, max(0, min(L[i] - s, e - s)), e - s) for i in range(len(L))]
        wav = fake.decode_jobs(codes_list, jobs, "cpu")
        piece = torch.zeros(len(L), keep * 1280)
        for b, j in enumerate(jobs):
            v = min(j.n_valid, keep) * 1280
            piece[b, :v] = wav[b, :v]
        pieces.append(piece)
    allw = torch.cat(pieces, -1)
    return [allw[i, : L[i] * 1280] for i in range(len(L))]


@pytest.mark.parametrize("overlap", [0, 5, 10, 25])
def test_decode_plan_matches_reference_loop_any_overlap(overlap):
    g = torch.Generator().manual_seed(100 + overlap)
    fake = FakeCodec()
    codes_list = [torch.randint(0, 2016, (8, n), generator=g) for n in (37, 625, 285, 936, 1)]
    L = [c.shape[-1] for c in codes_list]
    outs = [torch.zeros(n * 1280) for n in L]
    for _, dj in windows.plan_decode(L, overlap).items():
        w = fake.decode_jobs(codes_list, dj, "cpu")
        for k, j in enumerate(dj):
            off, n = windows.decode_keep(j, overlap)
            outs[j.item][off:off + n] = w[k, :n]
    for a, b in zip(outs, _reference_style_decode(fake, codes_list, overlap)):
        assert torch.equal(a, b)


def test_decode_plan_pad_lengths():
    groups = windows.plan_decode([37, 625, 285])
    assert sorted(groups) == [125, 375]                    # windows 0,1 are full; window 2 of the batch has T' = 125
    assert sorted((j.item, j.chunk, j.n_valid) for j in groups[375]) == [(0, 0, 37), (1, 0, 375), (1, 1, 375), (2, 0, 285), (2, 1, 35)]
    assert [(j.item, j.chunk, j.n_valid) for j in groups[125]] == [(1, 2, 125)]
    assert windows.decode_keep(groups[125][0]) == (500 * 1280, 125 * 1280)
    assert windows.plan_decode([]) == {} and windows.plan_encode([]) == []
    assert windows.plan_encode([0, 5]) == [windows.EncodeJob(1, 0, 5)]


def test_shard_is_a_partition():
    cost = [5, 1, 9, 3, 3, 7, 2]
    for world in (1, 2, 3, 8):
        sh = windows.shard_round_robin(len(cost), cost, world)
        assert sorted(j for s in sh for j in s) == list(range(len(cost)))
        assert max(sum(cost[j] for j in s) for s in sh) <= max(sum(cost) / world + max(cost), max(cost))


def test_joint_sharding_of_decode_groups():
    """Decode windows of different pad lengths run as separate launch sets: a small group goes to few ranks instead of
    leaving every rank a launch set of a handful of windows, and the total load stays balanced."""
    import random
    rnd = random.Random(5)
    big = [rnd.randint(25, 375) for _ in range(256)]
    small = [rnd.randint(1, 125) for _ in range(90)]
    for world in (1, 2, 4, 8):
        sh = windows.shard_groups([small, big], world)
        for costs, g in zip((small, big), sh):
            assert len(g) == world and sorted(j for s in g for j in s) == list(range(len(costs)))
        load = [sum(small[j] for j in sh[0][r]) + sum(big[j] for j in sh[1][r]) for r in range(world)]
        fair = (sum(small) + sum(big)) / world
        # (the rank that takes the small group is under-loaded on purpose: 10 frames of halo per window + one more launch set)
        assert max(load) <= fair + 375 + 10 * len(small) + 400 and min(load) >= fair - 4 * 375
        assert sum(1 for s in sh[0] if s) <= max(1, world // 4)                            # the small group is not scattered
    assert windows.shard_groups([], 4) == [] and windows.shard_groups([[]], 2) == [[[], []]]


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simwhisper_codec_b200.parallel import ShardedCodec
    g = torch.Generator().manual_seed(0)
    wavs = [torch.randn(n, generator=g) for n in lens]
    fake = FakeCodec()
    sc = ShardedCodec(fake, encode_jobs=fake.encode_jobs, decode_jobs=fake.decode_jobs)
    codes = sc.encode(wavs, device="cpu")["codes_list"]
    wav = sc.decode(codes, device="cpu")["syn_wav_list"]
    if rank == 0:
        q.put(([c.numpy().copy() for c in codes], [w.numpy().copy() for w in wav]))  # numpy: pickled by value
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,lens", [(2, [48123, 800000, 365000, 1280 * 250, 479999]),
                                        (3, [48123, 70000]),                       # a rank without any window
                                        (3, [1000000, 2000, 330000, 481000])])     # three decode groups over three ranks
def test_sharded_equals_single_process_gloo(world, lens):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    codes2, wav2 = q.get(timeout=120)
    codes2 = [torch.from_numpy(c) for c in codes2]
    wav2 = [torch.from_numpy(w) for w in wav2]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process
    g = torch.Generator().manual_seed(0)
    wavs = [torch.randn(n, generator=g) for n in lens]
    fake = FakeCodec()
    jobs = windows.plan_encode(lens)
    codes1 = fake.stitch_codes(fake.encode_jobs(wavs, jobs, "cpu"), lens, jobs, 10)
    for a, b in zip(codes1, codes2):
        assert torch.equal(a, b)
    outs = [torch.zeros(c.shape[-1] * 1280) for c in codes1]
    for _, dj in windows.plan_decode([c.shape[-1] for c in codes1]).items():
        w = fake.decode_jobs(codes1, dj, "cpu")
        for k, j in enumerate(dj):
            off, n = windows.decode_keep(j)
            outs[j.item][off:off + n] = w[k, :n]
    for a, b in zip(outs, wav2):
        assert torch.equal(a, b)


def test_plans_match_reference_loops_randomised():
    """Seeded sweep over batch compositions and overlaps (including hops that are not multiples of 1280 samples and items
    shorter than one code frame): the flat encode / decode plans must reproduce the reference's chunk loops exactly."""
    g = torch.Generator().manual_seed(2024)
    fake = FakeCodec()
    for trial in range(40):
        overlap = int(torch.randint(0, 26, (), generator=g))
        n_items = int(torch.randint(1, 6, (), generator=g))
        lens = [int(torch.randint(0, 1_300_000, (), generator=g)) if float(torch.rand((), generator=g)) > 0.2
                else int(torch.randint(0, 3000, (), generator=g)) for _ in range(n_items)]
        if max(lens) == 0:
            continue
        wavs = [torch.randn(n, generator=g) for n in lens]
        jobs = windows.plan_encode(lens, overlap)
        ours = fake.stitch_codes(fake.encode_jobs(wavs, jobs, "cpu"), lens, jobs, overlap)
        ref = _reference_style_encode(fake, wavs, overlap)
        assert [tuple(c.shape) for c in ours] == [tuple(c.shape) for c in ref], (trial, overlap, lens)
        assert all(torch.equal(a, b) for a, b in zip(ours, ref)), (trial, overlap, lens)
        codes_list = [torch.randint(0, 2016, (8, max(1, n // 1280)), generator=g) for n in lens]
        L = [c.shape[-1] for c in codes_list]
        outs = [torch.zeros(n * 1280) for n in L]
        for _, dj in windows.plan_decode(L, overlap).items():
            w = fake.decode_jobs(codes_list, dj, "cpu")
            for k, j in enumerate(dj):
                off, n = windows.decode_keep(j, overlap)
                outs[j.item][off:off + n] = w[k, :n]
        for a, b in zip(outs, _reference_style_decode(fake, codes_list, overlap)):
            assert torch.equal(a, b), (trial, overlap, L)
