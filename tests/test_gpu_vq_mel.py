"""GPU (-m gpu): VQ lookup / EMA / backward and the STFT-mel kernels against the reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vq_mel_oracle as V


@pytest.fixture(scope="module")
def vq(golden_dir):
    return np.load(os.path.join(golden_dir, "vq.npz"))


@pytest.fixture(scope="module")
def mel(golden_dir):
    return np.load(os.path.join(golden_dir, "mel.npz"))


def make_q(vq, mode):
    from ttts_b200.vqvae.quantize import ResidualVectorQuantizer
    q = ResidualVectorQuantizer(dimension=192, n_q=1, bins=1024).cuda()
    cb = q.vq.layers[0]._codebook
    cb.embed.copy_(torch.tensor(vq["E"])); cb.embed_avg.copy_(torch.tensor(vq["E"]) * 3.0)
    cb.cluster_size.copy_(torch.tensor(vq[mode + "/cluster_size_in"])); cb.inited.fill_(1)
    q.train(mode == "train")
    return q, cb


def test_state_dict_keys_match_reference():
    from ttts_b200.vqvae.quantize import ResidualVectorQuantizer
    q = ResidualVectorQuantizer(dimension=192, n_q=1, bins=1024)
    assert set(q.state_dict().keys()) == {"vq.layers.0._codebook." + k for k in ("inited", "cluster_size", "embed", "embed_avg")}


def test_vq_eval_bit_exact_indices(vq):
    q, cb = make_q(vq, "eval")
    x = torch.tensor(vq["x"]).cuda()
    quantized, codes, commit, qlist = q(x, layers=[0])
    assert codes.shape == (1,) + x.shape[:1] + x.shape[2:] and codes.dtype == torch.int64
    assert np.array_equal(codes.cpu().numpy(), vq["eval/codes"])              # bit-exact, incl. duplicate-row tie rule
    assert np.array_equal(quantized.cpu().numpy(), vq["eval/quantized"])      # a pure row gather
    assert float(commit) == 0.0 and len(qlist) == 1
    assert np.array_equal(q.encode(x).cpu().numpy(), vq["eval/encode"])
    np.testing.assert_array_equal(q.decode(q.encode(x)).cpu().numpy(), vq["eval/decode"])
    assert np.array_equal(cb.embed.cpu().numpy(), vq["E"])                    # eval never touches the buffers


def test_vq_train_forward_backward_ema(vq):
    q, cb = make_q(vq, "train")
    x = torch.tensor(vq["x"]).cuda().requires_grad_(True)
    quantized, codes, commit, _ = q(x, layers=[0])
    assert np.array_equal(codes.cpu().numpy(), vq["train/codes"])
    np.testing.assert_allclose(quantized.detach().cpu().numpy(), vq["train/quantized"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(float(commit), float(vq["train/commit"]), rtol=1e-5)
    ((quantized * torch.tensor(vq["train/dquantized"]).cuda()).sum() + commit * 3.0).backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), vq["train/dx"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(cb.cluster_size.cpu().numpy(), vq["train/cluster_size_out"], rtol=1e-6)
    np.testing.assert_allclose(cb.embed_avg.cpu().numpy(), vq["train/embed_avg_out"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(cb.embed.cpu().numpy(), vq["train/embed_out"], rtol=1e-5, atol=1e-6)


def test_vq_large_random_vs_oracle_with_tie_margin():
    from ttts_b200.vqvae.quantize import vq_lookup
    rs = np.random.RandomState(3)
    E = rs.standard_normal((1024, 192)).astype(np.float32)
    x = rs.standard_normal((20000, 192)).astype(np.float32)
    codes, q, _ = vq_lookup(torch.tensor(x).cuda(), torch.tensor(E).cuda(), False)
    got = codes.cpu().numpy()
    want = V.vq_quantize(x, E)
    margin = V.vq_margin(x, E, want)
    flips = got != want
    assert not np.any(flips & (margin > 1e-6)), "index mismatch away from an fp32 near-tie"
    assert flips.sum() <= 2
    assert np.array_equal(q.cpu().numpy(), E[got])


@pytest.mark.parametrize("K,D,N,bdn", [(1000, 64, 777, False), (130, 16, 65, False), (1024, 192, 18 * 7, True), (300, 24, 100, False)])
def test_vq_ragged_shapes_vs_oracle(K, D, N, bdn):
    """codebook sizes that are not a multiple of the 128-code tile, short x tiles, both input layouts; D = 24 takes the non-pipelined kernel"""
    from ttts_b200.vqvae.quantize import vq_lookup
    rs = np.random.RandomState(K + D)
    E = rs.standard_normal((K, D)).astype(np.float32)
    if bdn:
        xb = rs.standard_normal((7, D, N // 7)).astype(np.float32)
        x = np.ascontiguousarray(xb.transpose(0, 2, 1)).reshape(-1, D)
        codes, q, _ = vq_lookup(torch.tensor(xb).cuda(), torch.tensor(E).cuda(), True)
        qn = np.ascontiguousarray(q.cpu().numpy().transpose(0, 2, 1)).reshape(-1, D)
    else:
        x = rs.standard_normal((N, D)).astype(np.float32)
        codes, q, _ = vq_lookup(torch.tensor(x).cuda(), torch.tensor(E).cuda(), False)
        qn = q.cpu().numpy()
    got = codes.cpu().numpy()
    want = V.vq_quantize(x, E)
    flips = got != want
    assert got.min() >= 0 and got.max() < K
    assert not np.any(flips & (V.vq_margin(x, E, want) > 1e-6)) and flips.sum() <= 1
    assert np.array_equal(qn, E[got])


def test_vq_full_size_properties():
    """N = 2^20 vectors (dataset-extraction regime): codebook rows map to themselves; encode(decode(c)) == c."""
    from ttts_b200.vqvae.quantize import vq_lookup
    g = torch.Generator(device="cuda").manual_seed(0)
    E = torch.randn(1024, 192, device="cuda", generator=g)
    c0 = torch.randint(0, 1024, (1 << 20,), device="cuda", generator=g)
    x = E[c0].contiguous()
    codes, q, _ = vq_lookup(x, E, False)
    assert torch.equal(codes, c0) and torch.equal(q, x)
    noisy = x + 1e-3 * torch.randn(x.shape, device="cuda", generator=g)
    codes2, _, _ = vq_lookup(noisy, E, False)
    assert torch.equal(codes2, c0)


def test_vq_kmeans_init_and_empty_codebook():
    from ttts_b200.vqvae.quantize import ResidualVectorQuantizer
    q = ResidualVectorQuantizer(dimension=192, n_q=1, bins=64, kmeans_iters=5).cuda()
    x = torch.randn(8, 192, 18, device="cuda")
    q.eval()
    assert int(q.encode(x).abs().max()) == 0              # fresh (all-zero) codebook -> all codes 0
    q.train()
    torch.manual_seed(0)
    q(x)
    cb = q.vq.layers[0]._codebook
    assert float(cb.inited) == 1.0 and float(cb.embed.abs().sum()) > 0


def test_spectrogram_and_mel_vs_reference_golden(mel):
    from ttts_b200.vqvae.mel import spectrogram_torch, spec_to_mel_torch, mel_spectrogram_torch
    wav = torch.tensor(mel["wav"][:, :23040]).cuda()
    spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
    assert spec.shape == (3, 1025, 36) and spec.dtype == torch.float32
    np.testing.assert_allclose(spec.cpu().numpy(), mel["spec"], rtol=3e-4, atol=6e-5)
    m = spec_to_mel_torch(torch.tensor(mel["spec"]).cuda(), 2048, 128, 32000, 0, None).cpu().numpy()
    ok = mel["mel"] > np.log(1e-5) + 1e-3
    assert np.abs(m - mel["mel"])[ok].max() < 1e-4
    m2 = mel_spectrogram_torch(wav, 2048, 128, 32000, 640, 2048, 0, None, center=False).cpu().numpy()
    loud = mel["mel2"] > -6.0
    assert np.abs(m2 - mel["mel2"])[ok & loud].max() < 1e-4          # stated tolerance, away from the clamp floor
    assert np.abs(m2 - mel["mel2"])[ok].max() < 5e-3                  # pure-tone bands inside the reference's fp32 FFT noise
    # against the fp64 oracle the kernel itself is accurate everywhere that is audible
    ref = V.mel_spectrogram(mel["wav"][:, :23040])
    assert np.abs(m2 - ref)[loud].max() < 1e-4


def test_mel_features_24k_vs_reference_golden(mel):
    from ttts_b200.vqvae.mel import MelSpectrogramFeatures
    f = MelSpectrogramFeatures()(torch.tensor(mel["wav"]).cuda()).cpu().numpy()
    assert f.shape == (3, 100, 94)
    ok = mel["feats24"] > np.log(1e-7) + 1e-3
    loud = mel["feats24"] > -4.0
    assert np.abs(f - mel["feats24"])[ok & loud].max() < 2e-4
    assert np.abs(f - mel["feats24"])[ok].max() < 0.15
    one = MelSpectrogramFeatures()(torch.tensor(mel["wav"][0]).cuda())
    assert one.shape == (100, 94)


def test_batched_mel_extractor_equals_per_clip(mel, tmp_path):
    """`<wav>.mel.pth` files of the batched extractor (prepare/extract_mel.py) == the front end run one clip at a time, reference format."""
    from ttts_b200.prepare.extract_mel import extract_mel
    from ttts_b200.vqvae.mel import MelSpectrogramFeatures
    g = torch.Generator().manual_seed(3)
    clips = {"a.wav": torch.tensor(mel["wav"][0]).unsqueeze(0), "b.wav": torch.tensor(mel["wav"][1]), "c.wav": 0.1 * torch.randn(2, 31111, generator=g),
             "d.wav": 0.1 * torch.randn(1, 24000, generator=g), "odd.wav": 0.1 * torch.randn(1, 5000, generator=g), "tiny.wav": torch.zeros(1, 400)}
    paths = [str(tmp_path / k) for k in clips]
    done = extract_mel(paths, load_fn=lambda p: clips[os.path.basename(p)], batch_size=4)
    assert set(done) == set(paths[:5]) and not os.path.exists(paths[5] + ".mel.pth")
    fe = MelSpectrogramFeatures()
    for p in paths[:5]:
        w = clips[os.path.basename(p)]
        w = w.mean(0) if w.dim() == 2 and w.shape[0] > 1 else w.reshape(-1)
        got = torch.load(p + ".mel.pth")
        assert got.shape == (1, 100, 1 + w.shape[0] // 256) and got.dtype == torch.float32 and got.device.type == "cpu"
        assert torch.equal(got[0], fe(w.cuda()).cpu())
    ref = torch.tensor(mel["feats24"][0])
    got = torch.load(paths[0] + ".mel.pth")[0]
    ok = ref > -4.0
    assert (got - ref)[ok].abs().max() < 2e-4


@pytest.mark.parametrize("L,B", [(23041, 3), (4099, 2), (2048 + 5 * 640 + 7, 1)])
def test_spectrogram_odd_lengths_vs_oracle(L, B):
    """odd clip lengths: clips 1, 2, ... start at odd sample offsets (the kernel's scalar-load path), frame counts not a multiple of 4
    (the scalar spectrogram write-out), every frame of a short clip touches the reflect padding"""
    from ttts_b200.vqvae.mel import spectrogram_torch, mel_spectrogram_torch
    rs = np.random.RandomState(L)
    wav = np.clip(0.1 * rs.standard_normal((B, L)), -1, 1).astype(np.float32)
    s = spectrogram_torch(torch.tensor(wav).cuda(), 2048, 640, 2048).cpu().numpy()
    ref = V.spectrogram(wav)
    assert s.shape == ref.shape == (B, 1025, 1 + (L + 1408 - 2048) // 640)
    assert np.abs(s - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())
    m = mel_spectrogram_torch(torch.tensor(wav).cuda(), 2048, 128, 32000, 640, 2048, 0, None).cpu().numpy()
    assert np.abs(m - V.mel_spectrogram(wav)).max() < 1e-3


def test_stft_linearity_and_batch_independence():
    from ttts_b200.vqvae.mel import spectrogram_torch
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(64, 23040, device="cuda", generator=g) * 0.1
    s = spectrogram_torch(a, 2048, 640, 2048)
    s1 = spectrogram_torch(a[17:18], 2048, 640, 2048)
    assert torch.equal(s[17:18], s1)
    s2 = spectrogram_torch(2 * a, 2048, 640, 2048)
    big = s > 0.05
    assert ((s2[big] / s[big]) - 2).abs().max() < 1e-3


@pytest.mark.parametrize("N,layout_bdn", [(1 << 17, False), (4096 * 3 + 77, True), (5000, False)])
def test_vq_tensor_core_path_is_bit_identical_to_fp32(N, layout_bdn, monkeypatch):
    """Large N: distances from split-bf16 products on tcgen05 (vq_tc_scores_kernel), then an exact fp32 re-check of every candidate within
    the error bound of the best approximate score (vq_tc_finish_kernel).  Codes, dequantised rows and the commitment loss must be
    BIT-identical to the fp32 FMA kernel -- also with adversarial near-ties: duplicated codebook rows (lowest index must win), vectors
    exactly half-way between two codes, vectors equal to a code, and a cluster of codes closer together than the error bound (forces
    the exact full-scan fallback)."""
    from ttts_b200.vqvae.quantize import vq_lookup
    g = torch.Generator(device="cuda").manual_seed(N)
    E = torch.randn(1024, 192, device="cuda", generator=g)
    E[700] = E[3]; E[701] = E[3]                                  # exact duplicates
    E[800:806] = E[9] + 1e-6 * torch.randn(6, 192, device="cuda", generator=g)      # a cluster inside the error bound -> fallback scan
    x = torch.randn(N, 192, device="cuda", generator=g)
    x[0] = E[3]; x[1] = 0.5 * (E[5] + E[6]); x[2] = E[9]; x[3] = E[803]; x[4] = 0.0
    x[5:64] = E[torch.randint(0, 1024, (59,), device="cuda", generator=g)] + 1e-3 * torch.randn(59, 192, device="cuda", generator=g)
    if layout_bdn:
        Nn = 4096 * 3 + 77
        xin = x[:Nn].t().contiguous().view(1, 192, Nn)
    else:
        xin = x
    monkeypatch.setenv("TTTS_VQ_TC", "0")
    c0, q0, l0 = vq_lookup(xin, E, layout_bdn, want_quantized=True, want_commit=True)
    monkeypatch.setenv("TTTS_VQ_TC", "1")
    c1, q1, l1 = vq_lookup(xin, E, layout_bdn, want_quantized=True, want_commit=True)
    assert torch.equal(c0, c1), int((c0 != c1).sum())
    assert torch.equal(q0, q1) and float(l0) == float(l1)
    assert int(c1[0]) == 3 and int(c1[2]) in (9, 800, 801, 802, 803, 804, 805)


@pytest.mark.parametrize("K,D,N,bdn", [(1024, 192, 1152, True), (1024, 192, 65, False), (1000, 64, 777, False), (1024, 192, 2000, False), (130, 16, 300, False)])
def test_vq_split_sweep_is_bit_identical(K, D, N, bdn, monkeypatch):
    """small N (the encode metric's own 64 clips x 18 frames = 1 152 vectors): the code tiles are dealt out over grid.y and a second kernel
    merges the per-group winners.  Codes, dequantised rows and the commitment loss must be BIT-identical to the single launch, with
    duplicated codebook rows in different tile groups (lowest index wins), a vector equal to a code and a half-way vector."""
    from ttts_b200.vqvae.quantize import vq_lookup
    g = torch.Generator(device="cuda").manual_seed(N + K)
    E = torch.randn(K, D, device="cuda", generator=g)
    if K >= 1000:
        E[700] = E[3]; E[129] = E[3]; E[999] = E[520]
    x = torch.randn(N, D, device="cuda", generator=g)
    x[0] = E[3]; x[1] = 0.5 * (E[5] + E[K - 1]); x[2] = E[min(520, K - 1)]; x[3] = 0.0
    xin = x.view(N // 18, 18, D).transpose(1, 2).contiguous() if bdn else x
    monkeypatch.setenv("TTTS_VQ_SPLIT", "0")
    c0, q0, l0 = vq_lookup(xin, E, bdn, want_quantized=True, want_commit=True)
    monkeypatch.setenv("TTTS_VQ_SPLIT", "1")
    c1, q1, l1 = vq_lookup(xin, E, bdn, want_quantized=True, want_commit=True)
    assert torch.equal(c0, c1), int((c0 != c1).sum())
    assert torch.equal(q0, q1) and float(l0) == float(l1)
    if K >= 1000:
        assert int(c1.reshape(-1)[0]) == 3 and int(c1.reshape(-1)[2]) == 520
