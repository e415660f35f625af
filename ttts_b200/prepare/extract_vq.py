"""Batched VQ-code extraction to disk: the data edge either side of the encode hot path.

Replaces `ttts/prepare/2_save_vq_to_disk.py:13-17` + `ttts/prepare/extract_vq.py:10-25`, which push ONE clip at a time through the
codec in 4 spawned worker processes.  Here clips are grouped by code length and pushed through the sm_100a encode path
(`ttts_b200.vqvae.encoder.VQEncoder`, CUDA-graph replay per batch shape) 64 at a time.  The on-disk format is the reference's:
`<path>.vq.pth` = `torch.save(list[int])` (extract_vq.py:22-24), which `ttts/gpt/dataset.py` reads back with `torch.load`.

Waveform conditioning follows `ttts/vqvae/dataset.py:63-71`: first channel, clamp to [-1, 1], trim to a multiple of 2*hop samples
(one code = 2 hops = 1 280 samples at 32 kHz), clips shorter than 16 hops are skipped.  Resampling is the loader's business
(`load_fn` returns 32 kHz mono float32); the default loader reads `<path>.wav32k.pth` tensors, `torchaudio` is not required.

Only clips with the SAME trimmed length share a batch: the convolution stack has no per-layer length masking (neither has the
reference's), so padding a short clip into a longer batch would change its last codes.  Equal-length batching keeps every clip's codes
bit-identical to a batch-of-one run.
"""
import os
from collections import defaultdict

import torch

HOP = 640
MIN_HOPS = 16


def condition_wav(wav, hop=HOP):
    """dataset.py:63-71.  wav: [C, L] or [L] float tensor at 32 kHz -> [L'] clamped, L' = 2*hop*floor(L / hop / 2); None if too short."""
    if wav.dim() == 2:
        wav = wav[0]
    if wav.shape[-1] < MIN_HOPS * hop:
        return None
    n = int(hop * 2 * (wav.shape[-1] // hop // 2))
    return torch.clamp(wav[:n].float(), min=-1.0, max=1.0)


def plan_batches(lengths, batch_size=64):
    """lengths: list of conditioned clip lengths in samples (None = skipped).  Returns a list of index lists: equal-length clips together,
    at most `batch_size` per batch, longest first (big graphs are captured first, the tail of odd lengths runs last)."""
    groups = defaultdict(list)
    for i, n in enumerate(lengths):
        if n is not None:
            groups[int(n)].append(i)
    batches = []
    for n in sorted(groups, reverse=True):
        idx = groups[n]
        for s in range(0, len(idx), batch_size):
            batches.append(idx[s:s + batch_size])
    return batches


def save_codes(path, codes):
    """extract_vq.py:22-24: `<path>.vq.pth` holds a plain python list of ints."""
    outp = path + ".vq.pth"
    d = os.path.dirname(outp)
    if d:
        os.makedirs(d, exist_ok=True)
    torch.save([int(c) for c in codes], outp)
    return outp


def default_load(path):
    return torch.load(path + ".wav32k.pth")


@torch.no_grad()
def extract_vq(paths, encoder, load_fn=default_load, batch_size=64, device="cuda", graphed=True, on_error=None):
    """Encode every clip in `paths` and write `<path>.vq.pth`.  Returns {path: n_codes} for the clips written.
    `encoder(wav[B, L]) -> {"codes": [1, B, L / 1280]}` (VQEncoder); `encoder.encode_graphed` is used when present and `graphed`."""
    wavs, lengths = [], []
    for p in paths:
        try:
            w = condition_wav(load_fn(p))
        except Exception as e:          # the reference prints and skips unreadable files (extract_vq.py:13-18)
            if on_error is not None:
                on_error(p, e)
            w = None
        wavs.append(w)
        lengths.append(None if w is None else int(w.shape[-1]))
    done = {}
    for idx in plan_batches(lengths, batch_size):
        batch = torch.stack([wavs[i] for i in idx]).to(device, non_blocking=True)
        if graphed and hasattr(encoder, "encode_graphed"):
            codes = encoder.encode_graphed(batch)
        else:
            codes = encoder(batch)["codes"]
        codes = codes.reshape(codes.shape[-2], codes.shape[-1]).cpu()       # [1, B, N] -> [B, N]
        for row, i in enumerate(idx):
            save_codes(paths[i], codes[row].tolist())
            done[paths[i]] = codes.shape[-1]
    return done
