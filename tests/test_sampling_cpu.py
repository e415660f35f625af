"""CPU: the token-selection rules of inference_speech (ttts_b200/gpt/sampling.py) against HuggingFace's own logits processors -- the
third-party code the reference delegates to (ttts/gpt/model.py:552-559 -> transformers GenerationMixin)."""
import pytest
import torch

from ttts_b200.gpt import sampling as S

lp = pytest.importorskip("transformers.generation.logits_process")


def _scores(seed, B=5, V=1026):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, V, generator=g) * 3.0


def _ids(seed, B=5, n=40, V=1026):
    g = torch.Generator().manual_seed(100 + seed)
    ids = torch.randint(0, V, (B, n), generator=g)
    ids[:, :10] = 1                         # the fake text ids of the reference prompt
    return ids


@pytest.mark.parametrize("penalty", [1.0, 2.0, 1.3])
def test_repetition_penalty(penalty):
    sc, ids = _scores(1), _ids(1)
    want = lp.RepetitionPenaltyLogitsProcessor(penalty=penalty)(ids, sc.clone()) if penalty != 1.0 else sc
    assert torch.equal(S.repetition_penalty_(sc.clone(), ids, penalty), want)


@pytest.mark.parametrize("t", [0.8, 1.0, 0.2])
def test_temperature(t):
    sc = _scores(2)
    assert torch.equal(S.temperature_(sc.clone(), t), lp.TemperatureLogitsWarper(t)(None, sc.clone()))


@pytest.mark.parametrize("k", [1, 50, 5000])
def test_top_k(k):
    sc = _scores(3)
    assert torch.equal(S.top_k_(sc.clone(), k), lp.TopKLogitsWarper(top_k=k)(None, sc.clone()))


@pytest.mark.parametrize("p", [0.8, 0.3, 0.999])
def test_top_p(p):
    sc = _scores(4)
    assert torch.equal(S.top_p_(sc.clone(), p), lp.TopPLogitsWarper(top_p=p)(None, sc.clone()))


@pytest.mark.parametrize("mass", [0.9, 0.5])
def test_typical(mass):
    sc = _scores(5)
    assert torch.equal(S.typical_(sc.clone(), mass), lp.TypicalLogitsWarper(mass=mass)(None, sc.clone()))


def test_pipeline_order_and_defaults():
    """repetition penalty -> temperature -> top-k (HF default 50) -> top-p, as GenerationMixin assembles them for the reference's call"""
    sc, ids = _scores(6), _ids(6)
    chain = lp.LogitsProcessorList([lp.RepetitionPenaltyLogitsProcessor(2.0), lp.TemperatureLogitsWarper(0.8), lp.TopKLogitsWarper(50),
                                    lp.TopPLogitsWarper(0.8)])
    got = S.process_logits(sc.clone(), ids, do_sample=True, temperature=0.8, top_p=0.8, repetition_penalty=2.0)
    assert torch.equal(got, chain(ids, sc.clone()))
    assert int(torch.isfinite(got).sum(-1).max()) <= 50
    greedy = S.process_logits(sc.clone(), ids, do_sample=False, temperature=0.8, top_p=0.8, repetition_penalty=2.0)
    assert torch.equal(greedy, lp.RepetitionPenaltyLogitsProcessor(2.0)(ids, sc.clone()))          # warpers are sampling-only
    assert torch.equal(S.select_tokens(greedy, False), greedy.argmax(-1))
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    a, b = S.select_tokens(got, True, g1), torch.multinomial(torch.softmax(got, -1), 1, generator=g2).squeeze(1)
    assert torch.equal(a, b) and bool(torch.isfinite(got.gather(1, a[:, None])).all())
