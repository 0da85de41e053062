"""Command-line counterpart of the reference's inference.py:9-66: directory of audio -> batches -> encode -> decode ->
16-bit PCM wav, on the B200 kernels.  Same flags as the reference plus --precision, --codes_dir (dump the 11-byte-per-
frame code streams of bitstream.py) and --random_init (no checkpoint: deterministic random weights, for smoke runs).

    python -m simwhisper_codec_b200.inference --checkpoint_path weights/SimWhisperCodec.pt --input_dir in --output_dir out
"""
from __future__ import annotations

import argparse
import logging
import os

import torch
import yaml

from .audiocodec.model import AudioCodec
from .bitstream import pack_codes
from .utils.helpers import can_decode, find_audio_files, load_audio, save_audio, set_logging

_DEFAULT_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config", "SimWhisperCodec.yaml")


def main(argv=None) -> int:
    set_logging()
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", type=str, default=_DEFAULT_CFG)
    ap.add_argument("--checkpoint_path", type=str, default="./weights/SimWhisperCodec.pt")
    ap.add_argument("--device", type=str, default="cuda")
    ap.add_argument("--batch_size", type=int, default=8)
    ap.add_argument("--input_dir", type=str, default="input_wavs")
    ap.add_argument("--output_dir", type=str, default="output_wavs")
    ap.add_argument("--precision", type=str, default="bf16x3", choices=["bf16x3", "bf16", "fp32"],
                    help="bf16x3: parity-grade tensor-core mode (default); bf16: throughput mode; fp32: CUDA-core reference arithmetic")
    ap.add_argument("--codes_dir", type=str, default=None, help="also write <name>.swc code streams here")
    ap.add_argument("--random_init", action="store_true", help="deterministic random weights instead of a checkpoint")
    args = ap.parse_args(argv)

    device = torch.device(args.device)
    if args.random_init:
        from .weights import random_state_dict
        gp = yaml.safe_load(open(args.config_path))["generator_params"]
        generator = AudioCodec(gp, precision=args.precision)
        generator.load_state_dict(random_state_dict(gp, seed=0, exercise=True))
    else:
        generator = AudioCodec.load_from_checkpoint(config_path=args.config_path, ckpt_path=args.checkpoint_path,
                                                    precision=args.precision)
    generator.eval()

    audio_paths = find_audio_files(input_dir=args.input_dir)
    skipped = [p for p in audio_paths if not can_decode(p)]
    for p in skipped:
        logging.warning(f"Skipping {p}: no decoder for this format here (RIFF/WAVE is built in; flac / mp3 need soundfile or torchaudio)")
    audio_paths = [p for p in audio_paths if p not in skipped]
    os.makedirs(args.output_dir, exist_ok=True)
    if args.codes_dir:
        os.makedirs(args.codes_dir, exist_ok=True)
    logging.info(f"Processing {len(audio_paths)} audio files, output will be saved to {args.output_dir}")

    bs = max(1, args.batch_size)
    for i in range(0, len(audio_paths), bs):
        batch_paths = audio_paths[i:i + bs]
        logging.info(f"Processing batch {i // bs + 1}/{(len(audio_paths) + bs - 1) // bs}, files: {batch_paths}")
        # decoded on the host, moved to the device as it is, resampled to 16 kHz there (utils.helpers.resample is device-agnostic)
        wav_list = [load_audio(p, target_sample_rate=generator.input_sample_rate, device=device).squeeze() for p in batch_paths]
        logging.info(f"Successfully loaded {len(wav_list)} audio files with lengths {[len(w) for w in wav_list]} samples")
        codes_list = generator.encode(wav_list, overlap_seconds=10, device=device)["codes_list"]
        logging.info(f"Encoding completed, code lengths: {[c.shape[-1] for c in codes_list]}")
        syn_wav_list = generator.decode(codes_list, overlap_seconds=10, device=device)["syn_wav_list"]
        logging.info(f"Decoding completed, generated waveform lengths: {[len(w) for w in syn_wav_list]} samples")
        for path, codes, syn in zip(batch_paths, codes_list, syn_wav_list):
            stem = os.path.splitext(os.path.basename(path))[0]
            save_audio(os.path.join(args.output_dir, stem + ".wav"), syn.cpu().reshape(1, -1),
                       sample_rate=generator.output_sample_rate)
            if args.codes_dir:
                with open(os.path.join(args.codes_dir, stem + ".swc"), "wb") as f:
                    f.write(pack_codes(codes))
    logging.info("All audio processing completed")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
