"""Batched 24 kHz log-mel extraction to disk: `<wav>.mel.pth`, the second on-disk format either side of the hot path.

Replaces `ttts/prepare/mel_extract.py:11-34` + `ttts/prepare/save_mel_to_disk.py:18-22`, which load, resample and transform ONE file per
call in 8 spawned worker processes.  Here equal-length clips are stacked and pushed through the sm_100a STFT / mel kernel
(`ttts_b200.vqvae.mel.MelSpectrogramFeatures`, csrc/stft.cu) `batch_size` at a time; every output file is what the reference writes:
`torch.save(mel.cpu())` with `mel` of shape [1, 100, 1 + L // 256] (mel_extract.py:33-34; read back by `ttts/prepare/extract_vq.py:11-13`,
`ttts/diffusion/dataset.py:46`, `ttts/classifier/dataset.py:25`).

Waveform conditioning follows mel_extract.py:18-22: multi-channel audio is mixed down to mono by the channel MEAN (note: the VQ path
takes the FIRST channel instead, vqvae/dataset.py:63).  Resampling to 24 kHz is the loader's business (`load_fn` returns 24 kHz float32
[C, L] or [L]); the default loader reads `<path>.wav24k.pth` tensors, `torchaudio` is not required.

Only clips of the same length share a batch: the front end pads each clip by reflection (center=True), so the frames at a clip's end
depend on where it ends.  Equal-length batching keeps every file bit-identical to a batch-of-one run.
"""
import os

import torch

from .extract_vq import plan_batches          # equal-length clips together, at most batch_size per batch, longest first


def condition_wav(wav):
    """mel_extract.py:18-22.  [C, L] or [L] -> [L] float32 mono (channel mean); None for clips the reflect padding cannot handle."""
    if wav.dim() == 2:
        wav = wav.mean(dim=0) if wav.shape[0] > 1 else wav[0]
    wav = wav.float()
    if wav.shape[-1] <= 512:          # reflect padding of n_fft / 2 = 512 samples needs a longer clip (torch raises for these too)
        return None
    return wav


def save_mel(path, mel):
    """mel_extract.py:13,34: `<wav_file>.mel.pth` holds the CPU tensor [1, n_mels, frames]."""
    outp = path + ".mel.pth"
    d = os.path.dirname(outp)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save(mel.detach().cpu().unsqueeze(0).contiguous(), outp)
    return outp


def default_load(path):
    return torch.load(path + ".wav24k.pth")


@torch.no_grad()
def extract_mel(paths, extractor=None, load_fn=default_load, batch_size=64, device="cuda", on_error=None):
    """Transform every clip in `paths` and write `<path>.mel.pth`.  Returns {path: n_frames} for the clips written.
    `extractor(wav[B, L]) -> [B, n_mels, frames]`; default: MelSpectrogramFeatures() (24 kHz, n_fft 1024, hop 256, 100 mels)."""
    if extractor is None:
        from ..vqvae.mel import MelSpectrogramFeatures
        extractor = MelSpectrogramFeatures()
    wavs, lengths = [], []
    for p in paths:
        try:
            w = condition_wav(load_fn(p))
        except Exception as e:          # the reference prints and skips unreadable files (mel_extract.py:29-32)
            if on_error is not None:
                on_error(p, e)
            w = None
        wavs.append(w)
        lengths.append(None if w is None else int(w.shape[-1]))
    done = {}
    for idx in plan_batches(lengths, batch_size):
        batch = torch.stack([wavs[i] for i in idx]).to(device, non_blocking=True)
        mel = extractor(batch)
        for row, i in enumerate(idx):
            save_mel(paths[i], mel[row])
            done[paths[i]] = int(mel.shape[-1])
    return done
