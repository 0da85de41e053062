"""Multi-GPU paths on real devices (skipped below two GPUs): the window-sharded encode()/decode() of BASELINE configs[3]
(variable-length batch, one process per GPU over NCCL) must reproduce the single-GPU API bit for bit, and a model living on
cuda:1 must work whatever the process's current device is (every C entry point switches to the model's device)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import synthetic_wave

pytestmark = pytest.mark.gpu

LENS = [48123, 800000, 365000, 1280 * 250, 479999, 32000, 161 * 3, 480000, 250000, 90000]


def _need_two():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def _worker(rank, world, port, precision, q):
    import yaml
    from simwhisper_codec_b200 import AudioCodec
    from simwhisper_codec_b200.parallel import ShardedCodec
    from simwhisper_codec_b200.weights import random_state_dict
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gp = yaml.safe_load(open(os.path.join(root, "simwhisper_codec_b200", "config", "SimWhisperCodec.yaml")))["generator_params"]
    model = AudioCodec(gp, precision=precision, max_batch=8)
    model.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    wavs = [synthetic_wave(6000 + i, n) for i, n in enumerate(LENS)]
    sc = ShardedCodec(model)
    codes = sc.encode(wavs, device=dev)["codes_list"]
    wav = sc.decode(codes, device=dev)["syn_wav_list"]
    local = sc.decode(codes, device=dev, gather_wav=False)["local"]
    if rank == 0:
        c1 = model.encode(wavs, device=dev)["codes_list"]
        w1 = model.decode(c1, device=dev)["syn_wav_list"]
        ok_codes = all(torch.equal(a, b) for a, b in zip(codes, c1))
        ok_wav = all(torch.equal(a, b) for a, b in zip(wav, w1))
        ok_local = all(torch.equal(w1[i][off:off + len(x)], x) for i, off, x in local)
        q.put((ok_codes, ok_wav, ok_local, [c.cpu().numpy().copy() for c in c1], len(local)))
    else:
        assert wav is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_sharded_codec_two_gpus_equals_single_gpu(precision, sd_ex):
    _need_two()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, precision, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_codes, ok_wav, ok_local, c1, n_local = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok_codes and ok_wav and ok_local and n_local > 0
    assert [c.shape for c in c1] == [(8, n // 1280) for n in LENS]
    if precision == "bf16x3":       # and the sharded result is the reference's: < 0.1 % flips against the CPU oracle
        from oracle import port as oracle_port
        sub = [0, 5, 8]
        with torch.inference_mode():
            ref = oracle_port.encode(sd_ex, [synthetic_wave(6000 + i, LENS[i]) for i in sub])
        total = sum(r.numel() for r in ref)
        flips = sum((torch.from_numpy(c1[i]).long() != r.long()).sum().item() for i, r in zip(sub, ref))
        assert flips / total < 1e-3, (flips, total)


def test_model_on_second_device_with_other_current_device(gen_params, sd_ex):
    """The library makes the model's device current inside every entry point and restores the caller's afterwards."""
    _need_two()
    from simwhisper_codec_b200 import AudioCodec
    m0 = AudioCodec(gen_params, precision="bf16")
    m0.load_state_dict(sd_ex)
    m1 = AudioCodec(gen_params, precision="bf16")
    m1.load_state_dict(sd_ex)
    w = [synthetic_wave(7100, 200000), synthetic_wave(7101, 48123)]
    torch.cuda.set_device(0)
    c0 = m0.encode(w, device="cuda:0")["codes_list"]
    c1 = m1.encode(w, device="cuda:1")["codes_list"]            # current device is 0, the model lives on 1
    assert torch.cuda.current_device() == 0
    with torch.cuda.device(0):
        y1 = m1.decode(c1, device="cuda:1")["syn_wav_list"]
    y0 = m0.decode(c0, device="cuda:0")["syn_wav_list"]
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(c0, c1))
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(y0, y1))
    assert c1[0].device == torch.device("cuda", 1)
    # a tensor on the wrong device is refused, not silently mis-executed (raw C ABI: the Python wrapper would re-home the model)
    import ctypes as C
    from simwhisper_codec_b200 import _lib
    nat = m1._native_for(torch.device("cuda", 1))
    lat = torch.zeros(1, 32, 8, device="cuda:0")
    ll = torch.tensor([8], device="cuda:0")
    zq = torch.empty_like(lat)
    codes = torch.empty(8, 1, 8, dtype=torch.int32, device="cuda:0")
    rc = nat.lib.swc_quantize(nat.handle, C.c_void_p(lat.data_ptr()), C.c_void_p(ll.data_ptr()), 1, 8, C.c_void_p(zq.data_ptr()),
                              C.c_void_p(codes.data_ptr()), C.c_void_p(torch.cuda.current_stream(1).cuda_stream))
    assert rc != 0 and "lives on device" in _lib.last_error()
