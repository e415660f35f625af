"""Logit processors and the token-selection step of `UnifiedVoice.inference_speech` (ttts/gpt/model.py:533-562).

The reference delegates to HuggingFace `GenerationMixin.generate`; what that does for the arguments `ttts/api_zh.py:78-86` passes
(do_sample, top_p, temperature, repetition_penalty, num_return_sequences, max length, eos = pad = stop_mel_token) is restated here as
plain tensor functions in HF's order -- repetition penalty -> [typical] -> temperature -> top-k -> top-p -> softmax -> multinomial --
so the behaviour is checked against `transformers.generation.logits_process` on CPU (tests/test_sampling_cpu.py).  HF's default
`top_k = 50` applies whenever sampling is on and the caller does not override it (GenerationConfig), as in the reference's call.
`length_penalty` only affects beam search, which the reference's call does not use (num_beams = 1): accepted and ignored.
"""
import torch


def repetition_penalty_(scores, input_ids, penalty):
    """RepetitionPenaltyLogitsProcessor: every id already in `input_ids` has its score divided (if > 0) or multiplied (if < 0) by `penalty`."""
    if penalty == 1.0:
        return scores
    s = torch.gather(scores, 1, input_ids)
    s = torch.where(s < 0, s * penalty, s / penalty)
    return scores.scatter(1, input_ids, s)


def typical_(scores, mass=0.9, filter_value=-float("inf"), min_tokens_to_keep=1):
    """TypicalLogitsWarper (ttts/utils/typical_sampling.py; HF TypicalLogitsWarper): keep the tokens whose surprise is closest to the entropy."""
    normalized = torch.log_softmax(scores, dim=-1)
    p = torch.exp(normalized)
    ent = -(normalized * p).nansum(-1, keepdim=True)
    shifted = torch.abs((-normalized) - ent)
    sorted_scores, sorted_indices = torch.sort(shifted, descending=False)
    sorted_logits = scores.gather(-1, sorted_indices)
    cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    last_ind = (cumulative < mass).sum(dim=1)
    last_ind.clamp_(max=sorted_scores.shape[-1] - 1)
    remove_sorted = sorted_scores > sorted_scores.gather(1, last_ind.view(-1, 1))
    remove_sorted[..., :min_tokens_to_keep] = False
    remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
    return scores.masked_fill(remove, filter_value)


def temperature_(scores, temperature):
    return scores if temperature == 1.0 else scores / temperature


def top_k_(scores, top_k, filter_value=-float("inf"), min_tokens_to_keep=1):
    """TopKLogitsWarper."""
    if top_k is None or top_k <= 0:
        return scores
    k = min(max(int(top_k), min_tokens_to_keep), scores.shape[-1])
    remove = scores < torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(remove, filter_value)


def top_p_(scores, top_p, filter_value=-float("inf"), min_tokens_to_keep=1):
    """TopPLogitsWarper: ascending sort, drop the tail whose cumulative probability is <= 1 - top_p."""
    if top_p is None or top_p >= 1.0:
        return scores
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove_sorted = cumulative <= (1 - top_p)
    remove_sorted[..., -min_tokens_to_keep:] = False
    remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
    return scores.masked_fill(remove, filter_value)


def process_logits(scores, input_ids, do_sample=False, temperature=1.0, top_k=None, top_p=None, repetition_penalty=1.0,
                   typical_sampling=False, typical_mass=0.9):
    """scores [B, V] fp32 (last position), input_ids [B, n] int64 (everything generate() would have in `input_ids`, fake prompt ids included)."""
    scores = repetition_penalty_(scores, input_ids, float(repetition_penalty))
    if typical_sampling:
        scores = typical_(scores, typical_mass)
    if do_sample:
        scores = temperature_(scores, float(temperature))
        scores = top_k_(scores, 50 if top_k is None else top_k)
        scores = top_p_(scores, top_p)
    return scores


def select_tokens(scores, do_sample, generator=None):
    """greedy argmax, or one multinomial draw per row from softmax(scores)"""
    if not do_sample:
        return torch.argmax(scores, dim=-1)
    probs = torch.softmax(scores, dim=-1)
    return torch.multinomial(probs, num_samples=1, generator=generator).squeeze(1)


def generate_codes(step_logits, codes, n_prompt, n_max, text_positions, start_mel_token, stop_mel_token, do_sample=False, temperature=1.0,
                   top_k=None, top_p=None, repetition_penalty=1.0, typical_sampling=False, typical_mass=0.9, generator=None):
    """The token loop of `GPT2InferenceModel.generate` as the reference drives it (ttts/gpt/model.py:533-562), independent of how logits
    are produced: `step_logits(n) -> [B, V] float32` returns the mel logits at the last of the n codes currently in `codes[:, :n]`
    (`codes` [B, >= n_max + 1] int64 is updated in place, pre-filled with the conditioning codes in its first `n_prompt` columns).
    What HF's processors see as `input_ids` is rebuilt here: the id 1 for each of the `text_positions` text slots (model.py:541), then
    start_mel_token, then the codes so far.  Finished rows keep emitting stop_mel_token (pad_token_id = eos_token_id, model.py:555-556).
    Returns the number of code columns filled."""
    B, dev = codes.shape[0], codes.device
    ids = torch.cat([torch.ones(B, text_positions, dtype=torch.int64, device=dev),
                     torch.full((B, 1), start_mel_token, dtype=torch.int64, device=dev), codes], dim=1)
    unfinished = torch.ones(B, dtype=torch.bool, device=dev)
    n = n_prompt
    while n < n_max:
        scores = step_logits(n)
        scores = process_logits(scores, ids[:, :text_positions + 1 + n], do_sample, temperature, top_k, top_p, repetition_penalty,
                                typical_sampling, typical_mass)
        nxt = select_tokens(scores, do_sample, generator)
        nxt = torch.where(unfinished, nxt, torch.full_like(nxt, stop_mel_token))
        codes[:, n] = nxt
        ids[:, text_positions + 1 + n] = nxt
        unfinished = unfinished & (nxt != stop_mel_token)
        n += 1
        if not bool(unfinished.any()):                               # the same host sync HF's stopping criteria make every step
            break
    return n
