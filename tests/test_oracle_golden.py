"""CPU: the oracle restatement (oracle/gpt_oracle.py) against golden vectors minted from the REAL reference
(tests/golden/make_golden.py ran /root/reference/ttts/gpt/model.py in fp32 eval mode).  This pins the oracle."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import gpt_oracle as O


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    cfg = ast.literal_eval(str(z["cfg_json"]))
    return z, cfg


@pytest.mark.parametrize("name", ["gpt_tiny", "gpt_ragged"])
def test_oracle_matches_reference_golden(golden_dir, name):
    z, cfg = _load(golden_dir, name)
    params = O.init_params(cfg, seed=int(z["seed"]))
    text = torch.tensor(z["text"]); tl = torch.tensor(z["text_lengths"])
    codes = torch.tensor(z["codes"]); wl = torch.tensor(z["wav_lengths"])
    codes_in = codes.clone()
    lt, lm, logits = O.forward(params, cfg, text, tl, codes_in, wl)
    assert abs(float(lt) - float(z["loss_text"])) < 2e-6
    assert abs(float(lm) - float(z["loss_mel"])) < 2e-6
    np.testing.assert_allclose(logits.numpy(), z["mel_logits"], rtol=0, atol=2e-5)
    # set_mel_padding mutates the caller's tensor exactly like the reference (Appendix E #4)
    assert np.array_equal(codes_in.numpy(), z["codes_after"])
    lat = O.forward(params, cfg, text, tl, codes.clone(), wl, return_latent=True)
    np.testing.assert_allclose(lat.numpy(), z["latent"], rtol=0, atol=2e-5)
    _, _, _, grads = O.loss_and_grads(params, cfg, text, tl, codes, wl)
    for k, g in grads.items():
        ref = z["grad/" + k]
        denom = np.linalg.norm(ref) + 1e-12
        assert np.linalg.norm(g.numpy() - ref) / denom < 2e-5, k


def test_masked_oracle_matches_reference_in_training_mode(golden_dir):
    """TRAINING mode: the oracle with keep masks at the four GPT-2 dropout sites == the REAL reference in train() with exactly those masks
    injected in place of its dropout draws (tests/golden/make_golden.py::gpt_train_masked_case: eager attention, no gradient checkpointing).
    This pins the mask-taking oracle that tests/test_gpu_gpt.py::test_training_mode_step_matches_oracle_with_exported_masks checks the
    CUDA step against (there with the masks the kernels drew themselves)."""
    z, cfg = _load(golden_dir, "gpt_train_masked")
    params = O.init_params(cfg, seed=int(z["seed"]))
    masks = {k[5:]: torch.tensor(np.unpackbits(z[k])[:int(np.prod(z["mshape/" + k[5:]]))].reshape(z["mshape/" + k[5:]])) for k in z.files if k.startswith("mask/")}
    args = (torch.tensor(z["text"]), torch.tensor(z["text_lengths"]), torch.tensor(z["codes"]), torch.tensor(z["wav_lengths"]))
    lt, lm, logits, grads = O.loss_and_grads(params, cfg, *args, masks=masks, drop_scale=1.0 / 0.9)
    assert abs(float(lt) - float(z["loss_text"])) < 2e-6 and abs(float(lm) - float(z["loss_mel"])) < 2e-6
    np.testing.assert_allclose(logits.numpy(), z["mel_logits"], rtol=0, atol=2e-5)
    for k, g in grads.items():
        ref = z["grad/" + k]
        assert np.linalg.norm(g.numpy() - ref) / (np.linalg.norm(ref) + 1e-12) < 2e-5, k
    lt0, lm0, _ = O.forward(params, cfg, *[a.clone() for a in args])
    assert abs(float(lm0) - float(z["loss_mel"])) > 1e-3                 # the masks matter


def test_bf16_emulation_is_close_to_fp32(golden_dir):
    z, cfg = _load(golden_dir, "gpt_tiny")
    params = O.init_params(cfg, seed=int(z["seed"]))
    args = (torch.tensor(z["text"]), torch.tensor(z["text_lengths"]), torch.tensor(z["codes"]), torch.tensor(z["wav_lengths"]))
    lt, lm, logits = O.forward(params, cfg, *[a.clone() for a in args], emulate_bf16=True)
    assert abs(float(lm) - float(z["loss_mel"])) < 2e-3
    rel = np.linalg.norm(logits.numpy() - z["mel_logits"]) / np.linalg.norm(z["mel_logits"])
    assert rel < 2e-2


def test_flop_model_matches_survey():
    cfg2 = O.default_config(layers=12, model_dim=512, heads=8)
    cfg3 = O.default_config(layers=24, model_dim=1024, heads=16)
    assert abs(O.flops_per_step(cfg2, 8, 128, 512) / 1.3030e12 - 1) < 1e-3
    assert abs(O.flops_per_step(cfg3, 32, 128, 1024) / 7.3546e13 - 1) < 1e-3


def test_adamw_matches_torch():
    torch.manual_seed(0)
    p = {"a": torch.randn(7, 5), "b": torch.randn(11)}
    g = {k: torch.randn_like(v) * 3 for k, v in p.items()}
    tp = [torch.nn.Parameter(v.clone()) for v in p.values()]
    opt = torch.optim.AdamW(tp, lr=1e-3, betas=(0.9, 0.96), weight_decay=0.01)
    state = {}
    for step in range(1, 4):
        for t, gg in zip(tp, g.values()):
            t.grad = gg.clone()
        torch.nn.utils.clip_grad_norm_(tp, 1.0)
        opt.step()
        O.clip_and_adamw(p, g, state, 1e-3, step)
    for t, v in zip(tp, p.values()):
        assert torch.allclose(t.detach(), v, atol=1e-6)


def test_attn_dropout_hash_statistics():
    """The counter-based keep function used for attention-probability dropout (csrc/common.cuh attn_drop_*, numpy restatement in the
    oracle): exact keep rate, no visible correlation along keys / queries / diagonals / seeds, flat 2-D spectrum."""
    rows = np.arange(2048)
    T = 1156
    for seed in (0, 7, 2 ** 63 + 12345):
        m = O.attn_dropout_keep_mask(seed, rows, T, 0.1).astype(np.float64)
        n = m.size
        assert abs(m.mean() - (1 - 6554 / 65536)) < 4 * np.sqrt(0.09 / n)
        c = m - m.mean(); v = (c * c).mean()
        for (dr, dk) in ((0, 1), (0, 2), (0, 3), (0, 4), (0, 8), (1, 0), (1, 1), (2, 0), (16, 0)):
            r = (c[dr:, dk:] * c[:c.shape[0] - dr, :c.shape[1] - dk]).mean() / v
            assert abs(r) < 5 / np.sqrt(n), (seed, dr, dk, r)
        assert abs(m.mean(1).std() - np.sqrt(0.09 / T)) < 0.1 * np.sqrt(0.09 / T)            # per-row keep rates: binomial spread, no more
        assert abs(m.mean(0).std() - np.sqrt(0.09 / len(rows))) < 0.15 * np.sqrt(0.09 / len(rows))
    a = O.attn_dropout_keep_mask(5, rows, T, 0.1).astype(np.float64); b = O.attn_dropout_keep_mask(6, rows, T, 0.1).astype(np.float64)
    assert abs(((a - a.mean()) * (b - b.mean())).mean() / 0.09) < 5 / np.sqrt(a.size)
    h = O.attn_dropout_keep_mask(9, np.arange(1024), 1024, 0.5).astype(np.float64) - 0.5
    spec = np.abs(np.fft.fft2(h)) ** 2 / (1024 * 1024 * 0.25)
    spec[0, 0] = 0
    assert spec.max() < 25 and abs(spec.mean() - 1) < 0.01          # exponential(1) bins: max over 1M bins ~ 14
    # p = 0 keeps everything; the mask depends on the row id only through (seed, row)
    assert O.attn_dropout_keep_mask(1, rows[:4], 9, 0.0).all()
    assert np.array_equal(O.attn_dropout_keep_mask(3, [77], 64, 0.3)[0], O.attn_dropout_keep_mask(3, [5, 77], 64, 0.3)[1])
