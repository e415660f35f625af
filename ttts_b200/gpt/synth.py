"""Synthetic benchmark inputs and the algorithmic FLOP model of the GPT step."""
import torch


def synthetic_batch(B, TL, CL, seed=1234):
    """text ~ U{1..254} (B,TL), codes ~ U{0..1023} (B,CL), full lengths (nothing is clipped or stop-padded)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(1, 255, (B, TL), generator=g, dtype=torch.int64)
    codes = torch.randint(0, 1024, (B, CL), generator=g, dtype=torch.int64)
    return text, torch.full((B,), TL, dtype=torch.int64), codes, torch.full((B,), CL * 1024, dtype=torch.int64)


def flops_per_step(layers, model_dim, B, TL, CL, n_text_vocab=257, n_mel_vocab=1026):
    """F_step = 3 * B * [L (24 d^2 T + 2 T^2 d) + 2 d V_m (CL+2) + 2 d V_t (TL+2)]  (causal attention counted half, no
    recompute credit)."""
    d, L = model_dim, layers
    T = TL + CL + 4
    fwd = L * (24 * d * d * T + 2 * T * T * d) + 2 * d * n_mel_vocab * (CL + 2) + 2 * d * n_text_vocab * (TL + 2)
    return 3 * B * fwd
