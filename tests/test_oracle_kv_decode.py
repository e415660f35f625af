"""KV-cache decode (last item): the oracle's incremental restatement (oracle/gpt_oracle.py: kv_prefill / kv_decode_step)
pinned three ways on CPU:
  * against the REAL reference's cached branch stepped by hand (tests/golden/gpt_kvstep.npz, minted by make_golden.py kv): pos_shift = 1;
  * against the oracle's own full forward at every step: pos_shift = 0 (the rule the uncached path -- the one ttts/api_zh.py:51 uses -- implies);
  * through the generation loop against the tokens the REAL reference's uncached `inference_speech` produced (gpt_generate.npz).
The CUDA decode step (csrc/gpt_decode.cu) is compared with these same functions in tests/test_gpu_gpt.py."""
import ast
import os

import numpy as np
import torch

from oracle import gpt_oracle as O
from ttts_b200.gpt import sampling as S


def _prompt(cfg, text, cond):
    B = text.shape[0]
    text_in = torch.cat([torch.full((B, 1), cfg["start_text_token"]), text, torch.zeros(B, 1, dtype=torch.int64)], 1)      # [start, text, stop]
    mel_in = torch.cat([torch.full((B, 1), cfg["start_mel_token"]), cond], 1)                                               # [start_mel, codes]
    return text_in, mel_in


def test_cached_branch_matches_the_real_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "gpt_kvstep.npz"))
    cfg = ast.literal_eval(str(z["cfg_json"]))
    params = O.init_params(cfg, seed=int(z["seed"]))
    text, cond = torch.tensor(z["text"]), torch.tensor(z["cond"])
    text_in, mel_in = _prompt(cfg, text, cond)
    with torch.no_grad():
        cache, slot, lg = O.kv_prefill(params, cfg, text_in, mel_in, T_max=40)
        np.testing.assert_allclose(lg.numpy(), z["logits"][0], atol=2e-5, rtol=0)
        for s, tok in enumerate(z["tokens"]):
            lg = O.kv_decode_step(params, cfg, cache, slot, torch.tensor(tok), text_in.shape[1], pos_shift=1)
            slot += 1
            np.testing.assert_allclose(lg.numpy(), z["logits"][s + 1], atol=2e-5, rtol=0)
        # the uncached position rule gives different logits for the same tokens: the two branches of the reference disagree
        cache0, slot0, _ = O.kv_prefill(params, cfg, text_in, mel_in, T_max=40)
        lg0 = O.kv_decode_step(params, cfg, cache0, slot0, torch.tensor(z["tokens"][0]), text_in.shape[1], pos_shift=0)
        assert float((lg0 - torch.tensor(z["logits"][1])).abs().max()) > 1e-3


def test_cached_steps_equal_full_forward():
    cfg = O.default_config(layers=3, model_dim=128, heads=2, max_text_tokens=20, max_mel_tokens=30)
    params = O.init_params(cfg, seed=2)
    g = torch.Generator().manual_seed(11)
    B, TL, m, steps = 3, 7, 2, 9
    text = torch.randint(1, 255, (B, TL), generator=g)
    codes = torch.randint(0, 1024, (B, m + steps), generator=g)
    tl = torch.full((B,), TL, dtype=torch.int64)
    text_in, mel_in = _prompt(cfg, text, codes[:, :m])
    with torch.no_grad():
        cache, slot, lg = O.kv_prefill(params, cfg, text_in, mel_in, T_max=TL + 2 + m + 1 + steps)
        for n in range(m, m + steps):
            _, _, full = O.forward(params, cfg, text, tl, codes[:, :n].clone(), torch.full((B,), (n + 1) * 1024, dtype=torch.int64))
            np.testing.assert_allclose(lg.numpy(), full[:, :, n].numpy(), atol=2e-5, rtol=0)
            lg = O.kv_decode_step(params, cfg, cache, slot, codes[:, n], TL + 2)
            slot += 1


def test_cached_generation_reproduces_the_uncached_reference_tokens(golden_dir):
    z = np.load(os.path.join(golden_dir, "gpt_generate.npz"))
    cfg = ast.literal_eval(str(z["cfg_json"]))
    params = O.init_params(cfg, seed=int(z["seed"]))
    text, cond = torch.tensor(z["text"]), torch.tensor(z["cond"])
    B, TL, m = text.shape[0], text.shape[1], cond.shape[1]
    text_in, mel_in = _prompt(cfg, text, cond)

    def run(**kw):
        n_max = m + 12
        codes = torch.full((B, n_max + 1), cfg["stop_mel_token"], dtype=torch.int64)
        codes[:, :m] = cond
        st = {}

        def step_logits(n):
            with torch.no_grad():
                if n == m:
                    st["cache"], st["slot"], lg = O.kv_prefill(params, cfg, text_in, mel_in, T_max=TL + 3 + n_max)
                    return lg
                lg = O.kv_decode_step(params, cfg, st["cache"], st["slot"], codes[:, n - 1], TL + 2)
                st["slot"] += 1
                return lg
        n = S.generate_codes(step_logits, codes, m, n_max, TL + 2, cfg["start_mel_token"], cfg["stop_mel_token"], **kw)
        return codes[:, m:n]
    assert torch.equal(run(), torch.tensor(z["greedy"]))
    assert torch.equal(run(repetition_penalty=2.0), torch.tensor(z["greedy_rep2"]))
