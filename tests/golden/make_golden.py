"""Mint golden vectors by running the REAL reference (adelacvg/ttts @ /root/reference) on CPU, fp32, eval mode.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/gpt_tiny.npz, gpt_ragged.npz, gpt_train_masked.npz, gpt_generate.npz, gpt_kvstep.npz, disc.npz, flow.npz, text_encoder.npz, vqvae_step.npz, vqvae_full_step.npz, vq.npz, mel.npz, encoder.npz, diffusion.npz.  Weights are NOT stored: they are regenerated from
numpy seeds by oracle.gpt_oracle.init_params (torch-version independent), loaded into the reference module through its
state_dict, and the reference's outputs are stored.  The import shims follow SURVEY.md Appendix D; nothing under
/root/reference is modified or copied.
"""
import os
import sys
import types
import importlib.machinery as im

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

import numpy as np
import torch


def import_reference():
    s = types.ModuleType("transformers.utils.model_parallel_utils")
    s.get_device_map = s.assert_device_map = lambda *a, **k: None
    sys.modules["transformers.utils.model_parallel_utils"] = s
    t = types.ModuleType("ttts.utils.typical_sampling")
    t.TypicalLogitsWarper = object
    sys.modules["ttts.utils.typical_sampling"] = t
    import ttts.gpt.model as gm  # noqa: must precede the librosa stub
    import torchaudio
    lb = types.ModuleType("librosa"); lb.__spec__ = im.ModuleSpec("librosa", None)
    lu = types.ModuleType("librosa.util"); lu.normalize = lu.pad_center = lu.tiny = None
    lf = types.ModuleType("librosa.filters")
    lf.mel = lambda sr, n_fft, n_mels, fmin, fmax: torchaudio.functional.melscale_fbanks(
        n_fft // 2 + 1, fmin, fmax or sr / 2, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
    lb.util, lb.filters = lu, lf
    sys.modules.update({"librosa": lb, "librosa.util": lu, "librosa.filters": lf})
    e = types.ModuleType("encodec"); e.EncodecModel = object; sys.modules["encodec"] = e
    import logging
    logging.disable(logging.CRITICAL)
    return gm


def gpt_case(gm, name, cfg_over, B, TL, CL, text_lengths=None, wav_lengths=None, seed=0):
    from oracle import gpt_oracle as O
    cfg = O.default_config(**cfg_over)
    params = O.init_params(cfg, seed=seed)
    ref_kwargs = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens",
                                     "start_text_token", "number_mel_codes", "start_mel_token", "stop_mel_token")}
    model = gm.UnifiedVoice(**ref_kwargs, use_mel_codes_as_input=True, train_solo_embeddings=False).eval()
    sd = model.state_dict()
    assert set(sd.keys()) == set(params.keys()), (set(sd.keys()) ^ set(params.keys()))
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), (k, sd[k].shape, params[k].shape)
    model.load_state_dict({k: v.clone() for k, v in params.items()})
    text, tl, codes, wl = O.synthetic_batch(B, TL, CL, seed=1234)
    if text_lengths is not None:
        tl = torch.tensor(text_lengths, dtype=torch.int64)
    if wav_lengths is not None:
        wl = torch.tensor(wav_lengths, dtype=torch.int64)
    codes_in = codes.clone()
    loss_text, loss_mel, mel_logits = model(text, tl, codes_in, wl)
    loss = 0.01 * loss_text + 1.0 * loss_mel
    loss.backward()
    grads = {k: p.grad.detach().numpy() for k, p in model.named_parameters()}
    latent = model(text, tl, codes.clone(), wl, return_latent=True).detach()
    out = dict(
        cfg_json=np.array(repr(cfg)), B=B, TL=TL, CL=CL, seed=seed,
        text=text.numpy(), text_lengths=tl.numpy(), codes=codes.numpy(), wav_lengths=wl.numpy(),
        codes_after=codes_in.numpy(),          # set_mel_padding mutates the caller's tensor (Appendix E #4)
        loss_text=loss_text.detach().numpy(), loss_mel=loss_mel.detach().numpy(),
        mel_logits=mel_logits.detach().numpy(), latent=latent.numpy(),
    )
    for k, g in grads.items():
        out["grad/" + k] = g
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "loss_text %.6f loss_mel %.6f" % (float(loss_text), float(loss_mel)), "->", path, "%.1f KB" % (os.path.getsize(path) / 1e3))


def masked_train_masks(cfg, B, T, seed=77, p=0.1):
    """Seeded keep masks for the four GPT-2 dropout sites (embedding, attention probabilities, attention output, MLP output), in the
    naming oracle.gpt_oracle.gpt_hidden uses.  Any fixed masks pin the masked oracle to the reference; the CUDA path's own masks come from
    ttts_gpt_dropout_mask."""
    g = torch.Generator().manual_seed(seed)
    d, H = cfg["model_dim"], cfg["heads"]
    m = {"embd": (torch.rand(B, T, d, generator=g) >= p)}
    for l in range(cfg["layers"]):
        m["attn_p%d" % l] = torch.rand(B, H, T, T, generator=g) >= p
        m["attn_o%d" % l] = torch.rand(B, T, d, generator=g) >= p
        m["mlp_o%d" % l] = torch.rand(B, T, d, generator=g) >= p
    return m


def gpt_train_masked_case(gm):
    """The REAL reference in TRAINING mode with the dropout draws replaced by given masks: torch.nn.functional.dropout is patched (for the
    duration of the call) to multiply by the next mask of the queue -- the order HF's GPT2Model draws them: embedding `drop`, then per layer
    attention-probability dropout, attention-output `resid_dropout`, MLP `dropout` (eager attention so that the probability dropout is an
    nn.Dropout call and not inside SDPA; gradient checkpointing off so that nothing is recomputed).  Stores losses, logits and every gradient."""
    from oracle import gpt_oracle as O
    cfg = O.default_config(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=80)
    params = O.init_params(cfg, seed=0)
    kw = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens", "start_text_token",
                              "number_mel_codes", "start_mel_token", "stop_mel_token")}
    model = gm.UnifiedVoice(**kw, use_mel_codes_as_input=True, train_solo_embeddings=False, checkpointing=False).train()
    model.load_state_dict({k: v.clone() for k, v in params.items()})
    model.gpt.config._attn_implementation = "eager"
    B, TL, CL = 2, 12, 24
    T = TL + CL + 4
    text, tl, codes, wl = O.synthetic_batch(B, TL, CL, seed=1234)
    masks = masked_train_masks(cfg, B, T)
    scale = 1.0 / 0.9
    queue = [masks["embd"]]
    for l in range(cfg["layers"]):
        queue += [masks["attn_p%d" % l], masks["attn_o%d" % l], masks["mlp_o%d" % l]]
    calls = []
    real = torch.nn.functional.dropout

    def fake(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        mk = queue[len(calls)]
        calls.append(tuple(x.shape))
        assert tuple(mk.shape) == tuple(x.shape), (len(calls), mk.shape, x.shape)
        return x * mk.to(x.dtype) * scale
    torch.nn.functional.dropout = fake
    try:
        loss_text, loss_mel, mel_logits = model(text, tl, codes.clone(), wl)
        (0.01 * loss_text + loss_mel).backward()
    finally:
        torch.nn.functional.dropout = real
    assert len(calls) == len(queue), (len(calls), len(queue))
    out = dict(cfg_json=np.array(repr(cfg)), seed=0, mask_seed=77, B=B, TL=TL, CL=CL, text=text.numpy(), text_lengths=tl.numpy(), codes=codes.numpy(),
               wav_lengths=wl.numpy(), loss_text=loss_text.detach().numpy(), loss_mel=loss_mel.detach().numpy(), mel_logits=mel_logits.detach().numpy())
    for k, mk in masks.items():
        out["mask/" + k] = np.packbits(mk.numpy().reshape(-1))
        out["mshape/" + k] = np.array(mk.shape)
    for k, prm in model.named_parameters():
        out["grad/" + k] = prm.grad.detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "gpt_train_masked.npz")
    np.savez_compressed(path, **out)
    print("gpt_train_masked loss_text %.6f loss_mel %.6f (%d dropout calls) ->" % (float(loss_text), float(loss_mel), len(calls)), path,
          "%.1f KB" % (os.path.getsize(path) / 1e3))


def generate_case(gm):
    """Greedy `inference_speech` of the REAL reference (ttts/gpt/model.py:533-562, kv_cache=False as in ttts/api_zh.py:51).  transformers 5
    no longer mixes GenerationMixin into PreTrainedModel, so it is appended to the bases of the reference's GPT2InferenceModel here (a shim
    in the sense of SURVEY.md Appendix D: nothing under /root/reference is modified).  Also stores the last-position logits of the prompt."""
    from transformers import GenerationMixin
    from oracle import gpt_oracle as O
    if not hasattr(gm.GPT2InferenceModel, "generate"):
        gm.GPT2InferenceModel.__bases__ = gm.GPT2InferenceModel.__bases__ + (GenerationMixin,)
    cfg = O.default_config(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=60)
    params = O.init_params(cfg, seed=0)
    kw = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens", "start_text_token",
                              "number_mel_codes", "start_mel_token", "stop_mel_token")}
    model = gm.UnifiedVoice(**kw, use_mel_codes_as_input=True, train_solo_embeddings=False).eval()
    model.load_state_dict({k: v.clone() for k, v in params.items()})
    model.post_init_gpt2_config(use_deepspeed=False, kv_cache=False, half=False)
    g = torch.Generator().manual_seed(5)
    text = torch.randint(1, 255, (2, 9), generator=g)
    cond = torch.randint(0, 1024, (2, 6), generator=g)
    with torch.no_grad():
        greedy = model.inference_speech(text, cond, max_generate_length=12, do_sample=False)
        rep = model.inference_speech(text, cond, max_generate_length=12, do_sample=False, repetition_penalty=2.0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gpt_generate.npz"), cfg_json=np.array(repr(cfg)), seed=0, text=text.numpy(),
                        cond=cond.numpy(), greedy=greedy.numpy(), greedy_rep2=rep.numpy())
    print("gpt_generate: greedy", greedy.tolist(), "rep2", rep.tolist())


def kv_case(gm):
    """Cached decoding of the REAL reference, stepped by hand: `GPT2InferenceModel.forward` (ttts/gpt/model.py:106-171) on the full prompt with
    use_cache, then one id at a time with the returned `past_key_values` and an attention mask of the total length -- the call sequence HF 4.x
    `generate` makes when `kv_cache=True` (under transformers 5.5 `generate` hands the reference's `prepare_inputs_for_generation` a non-empty
    cache object on the first call, so the prompt is dropped; the model's own forward is unaffected, hence the manual stepping).  Pins the
    position rule of the cached branch (:144-147: index = tokens in the mel segment INCLUDING the one being fed, one more than the uncached
    branch gives the same token)."""
    import torch.nn.functional as F
    from oracle import gpt_oracle as O
    cfg = O.default_config(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=60)
    params = O.init_params(cfg, seed=0)
    kw = {k: cfg[k] for k in ("layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "number_text_tokens", "start_text_token",
                              "number_mel_codes", "start_mel_token", "stop_mel_token")}
    model = gm.UnifiedVoice(**kw, use_mel_codes_as_input=True, train_solo_embeddings=False).eval()
    model.load_state_dict({k: v.clone() for k, v in params.items()})
    model.post_init_gpt2_config(use_deepspeed=False, kv_cache=True, half=False)
    im = model.inference_model
    g = torch.Generator().manual_seed(5)
    text = torch.randint(1, 255, (2, 9), generator=g)
    cond = torch.randint(0, 1024, (2, 6), generator=g)
    steps = 6
    with torch.no_grad():
        ti = F.pad(text, (0, 1), value=model.stop_text_token)
        ti, _ = model.build_aligned_inputs_and_targets(ti, model.start_text_token, model.stop_text_token)
        emb = model.text_embedding(ti) + model.text_pos_embedding(ti)
        im.store_mel_emb(emb)
        mi, _ = model.build_aligned_inputs_and_targets(cond, model.start_mel_token, model.stop_mel_token)
        fake = torch.ones(2, emb.shape[1] + mi.shape[1], dtype=torch.long)
        fake[:, -mi.shape[1]:] = mi
        out = im(input_ids=fake, past_key_values=None, use_cache=True, attention_mask=torch.ones_like(fake), return_dict=True)
        pkv, lg, n = out.past_key_values, out.logits[:, -1], fake.shape[1]
        logits, tokens = [lg.numpy()], []
        for s in range(steps):
            tok = lg.argmax(-1)
            if s == 2:
                tok = torch.tensor([7, 1000])          # rows diverge (greedy keeps them identical on this seed)
            tokens.append(tok.numpy())
            n += 1
            out = im(input_ids=tok[:, None], past_key_values=pkv, use_cache=True, attention_mask=torch.ones(2, n, dtype=torch.long), return_dict=True)
            pkv, lg = out.past_key_values, out.logits[:, -1]
            logits.append(lg.numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gpt_kvstep.npz"), cfg_json=np.array(repr(cfg)), seed=0, text=text.numpy(),
                        cond=cond.numpy(), tokens=np.stack(tokens), logits=np.stack(logits))
    print("gpt_kvstep: tokens", np.stack(tokens).T.tolist())


def decoder_case():
    """The REAL reference Generator (ttts/vqvae/vq2.py:341-416, hyper-parameters of vqvae/config.json) on CPU: waveform for a seeded latent,
    and per parameter tensor the gradient norm / projection of L = <y, R>.  Oracle for the next scope row (decoder of the VQ-VAE-GAN step)."""
    from oracle import decoder_oracle as DO
    from ttts.vqvae.vq2 import Generator
    net = Generator(192, "1", [3, 7, 11], [[1, 3, 5]] * 3, [10, 8, 2, 2, 2], 512, [16, 16, 8, 2, 2], gin_channels=512).eval()
    P = DO.init_params(seed=9)
    sd = net.state_dict()
    assert set(sd.keys()) == set(P.keys()), sorted(set(sd.keys()) ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P)
    g0 = torch.Generator().manual_seed(31)
    z = torch.randn(2, 192, 6, generator=g0)
    gcond = torch.randn(2, 512, 1, generator=g0)
    y = net(z, g=gcond)
    R = torch.randn(y.shape, generator=torch.Generator().manual_seed(32))
    loss = (y * R).sum()
    loss.backward()
    names, norm, proj = [], [], []
    for k, prm in net.named_parameters():
        gk = prm.grad
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    path = os.path.join(ROOT, "tests", "golden", "decoder.npz")
    np.savez_compressed(path, z=z.numpy(), g=gcond.numpy(), y=y.detach().numpy(), loss=float(loss), names=np.array(names), norm=np.array(norm),
                        proj=np.array(proj))
    print("decoder ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), tuple(y.shape), "loss %.5f" % float(loss))


def disc_case():
    """The REAL reference MultiPeriodDiscriminator (ttts/vqvae/vq2.py:418-551) and GAN losses (ttts/vqvae/losses.py) on CPU, as the trainer
    uses them (ttts/vqvae/train.py:349-388): discriminator step on (y, y_hat.detach()), generator step = generator_loss + feature_loss.
    Stored: logits of every discriminator, per-feature-map means, the three losses, and per parameter tensor the gradient norm / projection
    of the discriminator loss; the gradient of the generator-side loss with respect to y_hat.  Oracle for the next scope row."""
    from oracle import disc_oracle as DO
    from ttts.vqvae.vq2 import MultiPeriodDiscriminator
    from ttts.vqvae import losses as RL
    net = MultiPeriodDiscriminator(False).eval()
    P = DO.init_params(seed=4)
    sd = net.state_dict()
    assert set(sd.keys()) == set(P.keys()), sorted(set(sd.keys()) ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P)
    g0 = torch.Generator().manual_seed(41)
    T = 2309                                         # not a multiple of 2, 3, 5, 7 or 11: every period discriminator reflect-pads
    y = torch.tanh(torch.randn(2, 1, T, generator=g0))
    y_hat = torch.tanh(y + 0.3 * torch.randn(2, 1, T, generator=g0)).requires_grad_(True)
    # discriminator step
    y_d_r, y_d_g, _, _ = net(y, y_hat.detach())
    loss_d, _, _ = RL.discriminator_loss(y_d_r, y_d_g)
    loss_d.backward()
    names, norm, proj = [], [], []
    for k, prm in net.named_parameters():
        gk = prm.grad
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    # generator step (discriminator weights are not stepped here; only dL/dy_hat matters downstream)
    y_d_r2, y_d_g2, fmap_r, fmap_g = net(y, y_hat)
    loss_fm = RL.feature_loss(fmap_r, fmap_g)
    loss_gen, _ = RL.generator_loss(y_d_g2)
    (loss_gen + loss_fm).backward()
    out = dict(y=y.numpy(), y_hat=y_hat.detach().numpy(), loss_d=float(loss_d), loss_fm=float(loss_fm), loss_gen=float(loss_gen),
               names=np.array(names), norm=np.array(norm), proj=np.array(proj), dy_hat=y_hat.grad.numpy(),
               fmap_mean=np.array([[float(f.mean()) for f in fm] + [0.0] * (7 - len(fm)) for fm in fmap_g]),
               fmap_abs=np.array([[float(f.abs().mean()) for f in fm] + [0.0] * (7 - len(fm)) for fm in fmap_g]))
    for i, (r, gg) in enumerate(zip(y_d_r, y_d_g)):
        out["d%d_real" % i] = r.detach().numpy(); out["d%d_gen" % i] = gg.detach().numpy()
    g1 = torch.Generator().manual_seed(42)
    z_p, logs_q, m_p, logs_p = [torch.randn(2, 192, 9, generator=g1) for _ in range(4)]
    z_mask = (torch.rand(2, 1, 9, generator=g1) > 0.2).float()
    out.update(kl_in=np.stack([t.numpy() for t in (z_p, logs_q, m_p, logs_p)]), kl_mask=z_mask.numpy(),
               kl=float(RL.kl_loss(z_p, logs_q, m_p, logs_p, z_mask)))
    path = os.path.join(ROOT, "tests", "golden", "disc.npz")
    np.savez_compressed(path, **out)
    print("disc ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "loss_d %.5f loss_gen %.5f loss_fm %.5f" % (float(loss_d), float(loss_gen), float(loss_fm)))


def flow_case():
    """The REAL reference flow, ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=512) (ttts/vqvae/vq2.py:209-246), forward direction as in
    SynthesizerTrn.forward (:858), followed by the KL term of the trainer (losses.py:47-61): z_p, the loss and per parameter tensor the
    gradient norm / projection, plus the gradients of z and g.  `post` is given non-zero weights (its zero init makes the flow the identity)."""
    from oracle import flow_oracle as FO
    from ttts.vqvae.vq2 import ResidualCouplingBlock
    from ttts.vqvae import losses as RL
    net = ResidualCouplingBlock(192, 192, 5, 1, 4, gin_channels=512).eval()
    P = FO.init_params(seed=6)
    sd = net.state_dict()
    assert set(sd.keys()) == set(P.keys()), sorted(set(sd.keys()) ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P)
    z, ge, mask, logs_q, m_p, logs_p = FO.golden_inputs()
    z.requires_grad_(True); ge.requires_grad_(True)
    z_p = net(z, mask, g=ge)
    loss = RL.kl_loss(z_p, logs_q, m_p, logs_p, mask)
    loss.backward()
    names, norm, proj = [], [], []
    for k, prm in net.named_parameters():
        gk = prm.grad
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    path = os.path.join(ROOT, "tests", "golden", "flow.npz")
    # inputs are NOT stored: the test regenerates them from the same seeded generator calls (FO.golden_inputs)
    np.savez_compressed(path, z_sum=float(z.sum()), z_p=z_p.detach().numpy(), loss=float(loss), names=np.array(names), norm=np.array(norm),
                        proj=np.array(proj), dz=z.grad.numpy(), dg=ge.grad.numpy())
    print("flow ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "kl %.6f" % float(loss))


def text_encoder_case():
    """The REAL reference `enc_p_2` = TextEncoder(192, 192, 768, 2, 6, 3, 0.1) (ttts/vqvae/vq2.py:101-164, built at :817-825) in eval mode:
    outputs (y, m_p, logs_p) and, for L = <m, R1> + <logs, R2>, per parameter tensor the gradient norm / projection, plus dL/dy and dL/dge."""
    from oracle import text_encoder_oracle as TO
    from ttts.vqvae.vq2 import TextEncoder
    net = TextEncoder(192, 192, 768, 2, 6, 3, 0.1).eval()
    P = TO.init_params(seed=8)
    sd = net.state_dict()
    assert set(sd.keys()) == set(P.keys()), sorted(set(sd.keys()) ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P)
    y, y_lengths, text, text_lengths, ge = TO.golden_inputs()
    y.requires_grad_(True); ge.requires_grad_(True)
    yo, m, logs = net(y, y_lengths, text.clone(), text_lengths, ge)
    gR = torch.Generator().manual_seed(62)
    R1, R2 = torch.randn(m.shape, generator=gR), torch.randn(m.shape, generator=gR)
    loss = (m * R1).sum() + (logs * R2).sum()
    loss.backward()
    names, norm, proj = [], [], []
    for k, prm in net.named_parameters():
        gk = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    path = os.path.join(ROOT, "tests", "golden", "text_encoder.npz")
    np.savez_compressed(path, y_sum=float(y.sum()), yo=yo.detach().numpy(), m=m.detach().numpy(), logs=logs.detach().numpy(), loss=float(loss),
                        names=np.array(names), norm=np.array(norm), proj=np.array(proj), dy=y.grad.numpy(), dge=ge.grad.numpy())
    print("text_encoder ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "loss %.5f" % float(loss))


def step_params():
    """name -> tensor for the whole generator (net_g) and the discriminators (net_d), assembled from the per-module oracle initialisers"""
    from oracle import decoder_oracle as DEC, disc_oracle as DIS, encoder_oracle as EO, flow_oracle as FO, text_encoder_oracle as TO
    G = dict(EO.init_params(seed=5))                                                   # enc_p.*, ref_enc.*, proj.*
    G.update({"enc_q." + k[len("enc_p."):]: v for k, v in EO.init_params(seed=7).items() if k.startswith("enc_p.")})
    G.update({"dec." + k: v for k, v in DEC.init_params(seed=9).items()})
    G.update({"flow." + k: v for k, v in FO.init_params(seed=6).items()})
    G.update({"enc_p_2." + k: v for k, v in TO.init_params(seed=8).items()})
    return G, DIS.init_params(seed=4)


def step_inputs():
    enc = np.load(os.path.join(ROOT, "tests", "golden", "encoder.npz"))
    g0 = torch.Generator().manual_seed(71)
    text = torch.randint(0, 256, (3, 15), generator=g0)
    return torch.tensor(enc["wav"]), torch.tensor(enc["lengths"]), text, torch.tensor([15, 9, 4]), torch.tensor(enc["E"])


def vqvae_step_case():
    """The generator half of ONE train step of the REAL reference (ttts/vqvae/train.py:336-395 over SynthesizerTrn.forward, vq2.py:843-871, and
    MultiPeriodDiscriminator): the five losses and, per net_g parameter tensor, the gradient norm / projection of
    loss_gen_all = loss_gen + loss_fm + loss_mel + kl_ssl + loss_kl.  Modules in eval mode (the TextEncoder's dropouts off) except the
    quantizer (training: straight-through + commitment loss); posterior noises and the segment starts come from torch.manual_seed(0) and are
    regenerated by the test in the same order; segment = 8 frames; wav_aug = wav (no augmentation)."""
    from ttts.vqvae.vq2 import SynthesizerTrn, MultiPeriodDiscriminator
    from ttts.vqvae import losses as RL
    from ttts.utils import commons
    from ttts.utils.data_utils import spectrogram_torch, spec_to_mel_torch, mel_spectrogram_torch
    import json
    cfg = json.load(open("/root/reference/ttts/vqvae/config.json"))
    SEG = 8
    net_g = SynthesizerTrn(1025, SEG, **cfg["vqvae"]).eval()
    net_d = MultiPeriodDiscriminator(False).eval()
    G, D = step_params()
    sd = net_g.state_dict()
    learn = {k for k in sd if not k.startswith("quantizer.") and not k.endswith(".filter")}
    assert learn == set(G.keys()), sorted(learn ^ set(G.keys()))[:10]
    net_g.load_state_dict(G, strict=False)
    net_d.load_state_dict(D)
    wav, lengths, text, text_lengths, E = step_inputs()
    cb = net_g.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(E); cb.embed_avg.copy_(E); cb.cluster_size.fill_(10); cb.inited.fill_(1)
    net_g.quantizer.train()
    spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
    torch.manual_seed(0)
    y_hat, kl_ssl, ids_slice, z_mask, (z, z_p, m_p, logs_p, m_q, logs_q), _ = net_g(wav, wav, lengths * 640, spec, spec, lengths, text.clone(), text_lengths)
    mel = spec_to_mel_torch(spec, 2048, 128, 32000, 0.0, None)
    y_mel = commons.slice_segments(mel, ids_slice, SEG)
    y_hat_mel = mel_spectrogram_torch(y_hat.squeeze(1), 2048, 128, 32000, 640, 2048, 0.0, None)
    y = commons.slice_segments(wav.unsqueeze(1), ids_slice * 640, SEG * 640)
    y_d_hat_r, y_d_hat_g, fmap_r, fmap_g = net_d(y, y_hat)
    loss_mel = torch.nn.functional.l1_loss(y_mel, y_hat_mel) * 45
    loss_kl = RL.kl_loss(z_p, logs_q, m_p, logs_p, z_mask) * 1.0
    loss_fm = RL.feature_loss(fmap_r, fmap_g)
    loss_gen, _ = RL.generator_loss(y_d_hat_g)
    total = loss_gen + loss_fm + loss_mel + kl_ssl * 1 + loss_kl
    total.backward()
    names, norm, proj = [], [], []
    for k, prm in net_g.named_parameters():
        gk = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    path = os.path.join(ROOT, "tests", "golden", "vqvae_step.npz")
    np.savez_compressed(path, ids_slice=ids_slice.numpy(), loss_gen=float(loss_gen), loss_fm=float(loss_fm), loss_mel=float(loss_mel),
                        kl_ssl=float(kl_ssl), loss_kl=float(loss_kl), total=float(total), y_hat_sum=float(y_hat.sum()), z_sum=float(z.sum()),
                        names=np.array(names), norm=np.array(norm), proj=np.array(proj))
    print("vqvae_step ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), len(names), "tensors; gen %.4f fm %.4f mel %.4f commit %.5f kl %.4f" %
          (float(loss_gen), float(loss_fm), float(loss_mel), float(kl_ssl), float(loss_kl)), "ids", ids_slice.tolist())


def vqvae_full_step_case():
    """ONE full optimisation step of the REAL reference in its own order (ttts/vqvae/train.py:372-395): discriminator loss on
    (y, y_hat.detach()) -> optim_d.step() -> generator + feature losses through the UPDATED discriminators -> optim_g.step(), both
    torch.optim.AdamW(lr 1e-4, betas (0.8, 0.99), eps 1e-9) (train.py:193-205).  Stored: the losses and, per parameter tensor of net_g and
    net_d, the norm of the update and its projection on a seeded direction.  Same inputs / modes / seed as vqvae_step_case."""
    from ttts.vqvae.vq2 import SynthesizerTrn, MultiPeriodDiscriminator
    from ttts.vqvae import losses as RL
    from ttts.utils import commons
    from ttts.utils.data_utils import spectrogram_torch, spec_to_mel_torch, mel_spectrogram_torch
    import json
    cfg = json.load(open("/root/reference/ttts/vqvae/config.json"))
    SEG = 8
    net_g = SynthesizerTrn(1025, SEG, **cfg["vqvae"]).eval()
    net_d = MultiPeriodDiscriminator(False).eval()
    G, D = step_params()
    net_g.load_state_dict(G, strict=False)
    net_d.load_state_dict(D)
    wav, lengths, text, text_lengths, E = step_inputs()
    cb = net_g.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(E); cb.embed_avg.copy_(E); cb.cluster_size.fill_(10); cb.inited.fill_(1)
    net_g.quantizer.train()
    optim_g = torch.optim.AdamW(net_g.parameters(), 1e-4, betas=(0.8, 0.99), eps=1e-9)
    optim_d = torch.optim.AdamW(net_d.parameters(), 1e-4, betas=(0.8, 0.99), eps=1e-9)
    g_before = {k: v.detach().clone() for k, v in net_g.named_parameters()}
    d_before = {k: v.detach().clone() for k, v in net_d.named_parameters()}
    spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
    torch.manual_seed(0)
    y_hat, kl_ssl, ids_slice, z_mask, (z, z_p, m_p, logs_p, m_q, logs_q), _ = net_g(wav, wav, lengths * 640, spec, spec, lengths, text.clone(), text_lengths)
    mel = spec_to_mel_torch(spec, 2048, 128, 32000, 0.0, None)
    y_mel = commons.slice_segments(mel, ids_slice, SEG)
    y_hat_mel = mel_spectrogram_torch(y_hat.squeeze(1), 2048, 128, 32000, 640, 2048, 0.0, None)
    y = commons.slice_segments(wav.unsqueeze(1), ids_slice * 640, SEG * 640)
    y_d_hat_r, y_d_hat_g, _, _ = net_d(y, y_hat.detach())
    loss_disc, _, _ = RL.discriminator_loss(y_d_hat_r, y_d_hat_g)
    optim_d.zero_grad(); loss_disc.backward(); optim_d.step()
    y_d_hat_r, y_d_hat_g, fmap_r, fmap_g = net_d(y, y_hat)
    loss_mel = torch.nn.functional.l1_loss(y_mel, y_hat_mel) * 45
    loss_kl = RL.kl_loss(z_p, logs_q, m_p, logs_p, z_mask) * 1.0
    loss_fm = RL.feature_loss(fmap_r, fmap_g)
    loss_gen, _ = RL.generator_loss(y_d_hat_g)
    total = loss_gen + loss_fm + loss_mel + kl_ssl * 1 + loss_kl
    optim_g.zero_grad(); total.backward(); optim_g.step()
    out = dict(loss_disc=float(loss_disc), loss_gen=float(loss_gen), loss_fm=float(loss_fm), total=float(total))
    for tag, net, before in (("g", net_g, g_before), ("d", net_d, d_before)):
        names, norm, proj = [], [], []
        for k, prm in net.named_parameters():
            dlt = prm.detach() - before[k]
            d = torch.randn(dlt.shape, generator=torch.Generator().manual_seed(len(names)))
            names.append(k); norm.append(float(dlt.norm())); proj.append(float((dlt * d).sum()))
        out[tag + "_names"], out[tag + "_norm"], out[tag + "_proj"] = np.array(names), np.array(norm), np.array(proj)
    path = os.path.join(ROOT, "tests", "golden", "vqvae_full_step.npz")
    np.savez_compressed(path, **out)
    print("vqvae_full_step ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "disc %.4f gen %.4f fm %.4f" % (float(loss_disc), float(loss_gen), float(loss_fm)))


def vq_case():
    """EuclideanCodebook / ResidualVectorQuantizer (ttts/vqvae/core_vq.py:96-382, quantize.py:28-118)."""
    from ttts.vqvae.quantize import ResidualVectorQuantizer
    rs = np.random.RandomState(7)
    D, K, B, N = 192, 1024, 6, 18
    E = rs.standard_normal((K, D)).astype(np.float32)
    x = (rs.standard_normal((B, D, N)) * 1.0).astype(np.float32)
    # a few exact duplicates of codebook rows and duplicate codebook rows (first-max-wins, core_vq.py:181)
    E[700] = E[13]
    x[0, :, 0] = E[13]
    x[1, :, 3] = E[999]
    out = {"E": E, "x": x}
    for mode in ("eval", "train"):
        q = ResidualVectorQuantizer(dimension=D, n_q=1, bins=K)
        cb = q.vq.layers[0]._codebook
        cb.embed.copy_(torch.tensor(E)); cb.embed_avg.copy_(torch.tensor(E) * 3.0)
        cs = torch.tensor(rs.uniform(2.5, 9.0, size=(K,)).astype(np.float32))
        cb.cluster_size.copy_(cs); cb.inited.fill_(1)
        out[mode + "/cluster_size_in"] = cs.numpy()
        q.train(mode == "train")
        xt = torch.tensor(x, requires_grad=True)
        quantized, codes, commit, qlist = q(xt, layers=[0])
        if mode == "train":   # eval: output is a buffer gather, nothing requires grad (core_vq.py:310-318)
            w = torch.tensor(rs.standard_normal(quantized.shape).astype(np.float32))
            out["train/dquantized"] = w.numpy()
            ((quantized * w).sum() + commit * 3.0).backward()
        out[mode + "/quantized"] = quantized.detach().numpy()
        out[mode + "/codes"] = codes.numpy()
        out[mode + "/commit"] = commit.detach().numpy()
        if mode == "train":
            out[mode + "/dx"] = xt.grad.numpy()
        out[mode + "/embed_out"] = cb.embed.numpy().copy()
        out[mode + "/embed_avg_out"] = cb.embed_avg.numpy().copy()
        out[mode + "/cluster_size_out"] = cb.cluster_size.numpy().copy()
        enc = q.encode(torch.tensor(x))
        out[mode + "/encode"] = enc.numpy()
        out[mode + "/decode"] = q.decode(enc).detach().numpy()
    path = os.path.join(ROOT, "tests", "golden", "vq.npz")
    np.savez_compressed(path, **out)
    print("vq ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3))


def mel_case():
    """spectrogram_torch / spec_to_mel_torch / mel_spectrogram_torch (ttts/utils/data_utils.py:52-156) and
    MelSpectrogramFeatures (ttts/vocoder/feature_extractors.py:28-49)."""
    from ttts.utils.data_utils import spectrogram_torch, spec_to_mel_torch, mel_spectrogram_torch
    from ttts.vocoder.feature_extractors import MelSpectrogramFeatures
    g = torch.Generator().manual_seed(1234)
    wav = torch.clamp(0.1 * torch.randn(3, 24000, generator=g), -1, 1)
    t = torch.arange(24000) / 24000.0
    wav[1] = 0.5 * torch.sin(2 * math.pi * 440.0 * t) + 0.1 * torch.sin(2 * math.pi * 3000.0 * t)
    wav32 = wav[:, :23040].contiguous()
    spec = spectrogram_torch(wav32, 2048, 640, 2048, center=False)
    mel = spec_to_mel_torch(spec, 2048, 128, 32000, 0, None)
    mel2 = mel_spectrogram_torch(wav32, 2048, 128, 32000, 640, 2048, 0, None, center=False)
    feats = MelSpectrogramFeatures()(wav)
    # librosa is absent in the build container: the Slaney basis comes from torchaudio; store it so the
    # oracle / CUDA path can be checked against the exact basis the reference multiplied with.
    import librosa
    basis = librosa.filters.mel(sr=32000, n_fft=2048, n_mels=128, fmin=0, fmax=None)
    fb24 = MelSpectrogramFeatures().mel_spec.mel_scale.fb.numpy()
    path = os.path.join(ROOT, "tests", "golden", "mel.npz")
    np.savez_compressed(path, wav=wav.numpy(), spec=spec.numpy().astype(np.float32), mel=mel.numpy(), mel2=mel2.numpy(),
                        feats24=feats.numpy(), basis32=np.asarray(basis, dtype=np.float32), fb24=fb24)
    print("mel ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), spec.shape, mel.shape, feats.shape)


def encoder_case():
    """Encode half of SynthesizerTrn with the REAL reference modules (vq2.py:826-836, 843-852): MelStyleEncoder, PosteriorAudioEncoder,
    proj, quantizer.  Weights come from oracle.encoder_oracle.init_params (numpy-seeded) through load_state_dict."""
    from oracle import encoder_oracle as EO
    from ttts.vqvae import modules
    from ttts.vqvae.vq2 import PosteriorAudioEncoder
    from ttts.vqvae.quantize import ResidualVectorQuantizer
    from ttts.utils.data_utils import spectrogram_torch

    class Half(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc_p = PosteriorAudioEncoder(1025, 192, 192, 5, 1, 16, gin_channels=512)
            self.ref_enc = modules.MelStyleEncoder(1025, style_vector_dim=512)
            self.quantizer = ResidualVectorQuantizer(dimension=192, n_q=1, bins=1024)
            self.proj = torch.nn.Conv1d(192, 192, 2, stride=2)

    net = Half().eval()
    P = EO.init_params(seed=5)
    sd = net.state_dict()
    learn = {k for k in sd if not k.startswith("quantizer.") and not k.endswith(".filter")}
    assert learn == set(P.keys()), sorted(learn ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P, strict=False)
    rs = np.random.RandomState(11)
    E = rs.standard_normal((1024, 192)).astype(np.float32)
    cb = net.quantizer.vq.layers[0]._codebook
    cb.embed.copy_(torch.tensor(E)); cb.embed_avg.copy_(torch.tensor(E)); cb.cluster_size.fill_(10); cb.inited.fill_(1)
    g = torch.Generator().manual_seed(77)
    wav = torch.clamp(0.1 * torch.randn(3, 23040, generator=g), -1, 1)
    lengths = torch.tensor([36, 30, 17])
    eps = torch.randn(3, 192, 36, generator=g)
    spec = spectrogram_torch(wav, 2048, 640, 2048, center=False)
    y_mask = (torch.arange(36)[None, :] < lengths[:, None]).float().unsqueeze(1)
    with torch.no_grad():
        ge = net.ref_enc(spec * y_mask, y_mask)
        torch.manual_seed(0)
        _, m, logs = net.enc_p(spec, wav.unsqueeze(1), y_mask, g=ge)
        z = (m + eps * torch.exp(logs)) * y_mask                      # vq2.py:744 with the noise fixed
        x = net.proj(z)
        quantized, codes, commit, _ = net.quantizer(x, layers=[0])
        # intermediate anchors for bring-up
        a = net.enc_p.down_pre(wav.unsqueeze(1))
        a1 = net.enc_p.downs[0](a)
        r0 = net.enc_p.resblocks[0](a1)
    # gradients of the encode half (the oracle for the next scope row, the encoder's backward): L = <z, R> + 0.5 mean(x^2) with a seeded R;
    # per parameter tensor the L2 norm of its gradient and its projection on a seeded direction, plus a few small tensors in full
    gR = torch.Generator().manual_seed(123)
    R = torch.randn(3, 192, 36, generator=gR)
    for prm in net.parameters():
        prm.grad = None
    ge_g = net.ref_enc(spec * y_mask, y_mask)
    _, m_g, logs_g = net.enc_p(spec, wav.unsqueeze(1), y_mask, g=ge_g)
    z_g = (m_g + eps * torch.exp(logs_g)) * y_mask
    x_g = net.proj(z_g)
    loss = (z_g * R).sum() + 0.5 * (x_g ** 2).mean()
    loss.backward()
    gnames, gnorm, gproj = [], [], []
    full = {}
    for k, prm in net.named_parameters():
        if k.startswith("quantizer."):
            continue
        gk = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(gnames)))
        gnames.append(k); gnorm.append(float(gk.norm())); gproj.append(float((gk * d).sum()))
        if gk.numel() <= 4096 and len(full) < 12:
            full["gradfull/" + k] = gk.numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "encoder_grads.npz")
    np.savez_compressed(path, names=np.array(gnames), norm=np.array(gnorm), proj=np.array(gproj), loss=float(loss), **full)
    print("encoder grads ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), len(gnames), "tensors, loss %.6f" % float(loss))
    path = os.path.join(ROOT, "tests", "golden", "encoder.npz")
    np.savez_compressed(path, wav=wav.numpy(), lengths=lengths.numpy(), eps=eps.numpy(), E=E, ge=ge.numpy(), m=m.numpy(), logs=logs.numpy(),
                        z=z.numpy(), x=x.numpy(), codes=codes.numpy(), quantized=quantized.numpy(),
                        down0=a1.numpy()[:, :, :64], res0=r0.numpy()[:, :, :64],
                        filt=net.enc_p.activation_post.upsample.filter.numpy().reshape(-1))
    print("encoder ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "codes", codes.shape, codes.flatten()[:8].tolist())


def diffusion_case():
    """The REAL `AA_diffusion` (ttts/diffusion/aa_model.py:182-287) in train() mode under the REAL `SpacedDiffusion.training_losses`
    (ttts/utils/diffusion.py:930-1014) built exactly as ttts/diffusion/train.py:90-92, one micro-step of train.py:168-180: model output,
    mse / vb terms, loss, and per parameter tensor the gradient norm / projection.  The model's random decisions are pinned by patching the two
    draws it makes (torch.rand for the unconditioned samples, random.random for the layer drop) to the values oracle.golden_inputs() names."""
    import random as pyrandom
    from oracle import diffusion_oracle as DO
    kd = types.ModuleType("k_diffusion"); ks = types.ModuleType("k_diffusion.sampling")
    ks.sample_dpmpp_2m = ks.sample_euler_ancestral = None; kd.sampling = ks
    sys.modules["k_diffusion"] = kd; sys.modules["k_diffusion.sampling"] = ks
    import ttts.diffusion.aa_model as A
    from ttts.utils.diffusion import SpacedDiffusion, space_timesteps, get_named_beta_schedule
    cfg = DO.default_config(**DO.GOLDEN_CFG)
    net = A.AA_diffusion(**cfg, dropout=0, layer_drop=0.1).train()
    P = DO.init_params(cfg, seed=12)
    sd = net.state_dict()
    assert set(sd.keys()) == set(P.keys()), sorted(set(sd.keys()) ^ set(P.keys()))[:10]
    for k in P:
        assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
    net.load_state_dict(P)
    diffuser = SpacedDiffusion(use_timesteps=space_timesteps(1000, [1000]), model_mean_type="epsilon", model_var_type="learned_range",
                               loss_type="mse", betas=get_named_beta_schedule("linear", 1000), conditioning_free=False, conditioning_free_k=2.0)
    I = DO.golden_inputs()
    B = I["x_start"].shape[0]
    n_layers = cfg["num_layers"] + 3
    draws = iter([0.0 if i in I["dropped"] else 0.5 for i in range(1, n_layers - 1)])
    real_rand, real_random = torch.rand, pyrandom.random
    captured = {}

    def fake_rand(*a, **k):
        return torch.where(I["uncond"], 0.0, 0.5).reshape(B, 1, 1).float()
    hook = net.out.register_forward_hook(lambda m, i, o: captured.__setitem__("out", o.detach().clone()))
    torch.rand, pyrandom.random = fake_rand, lambda: next(draws)
    try:
        terms = diffuser.training_losses(model=net, x_start=I["x_start"], t=torch.tensor(I["t"]), noise=I["noise"],
                                         model_kwargs={"latent": I["latent"], "refer": I["refer"]})
    finally:
        torch.rand, pyrandom.random = real_rand, real_random
        hook.remove()
    assert next(draws, None) is None
    loss = terms["loss"].mean()
    loss.backward()
    names, norm, proj = [], [], []
    for k, prm in net.named_parameters():
        gk = prm.grad if prm.grad is not None else torch.zeros_like(prm)
        d = torch.randn(gk.shape, generator=torch.Generator().manual_seed(len(names)))
        names.append(k); norm.append(float(gk.norm())); proj.append(float((gk * d).sum()))
    path = os.path.join(ROOT, "tests", "golden", "diffusion.npz")
    np.savez_compressed(path, model_out=captured["out"].numpy(), mse=terms["mse"].detach().numpy(), vb=terms["vb"].detach().numpy(), loss=float(loss),
                        names=np.array(names), norm=np.array(norm), proj=np.array(proj))
    print("diffusion ->", path, "%.1f KB" % (os.path.getsize(path) / 1e3), "loss %.6f" % float(loss), "mse", terms["mse"].detach().numpy(), "vb", terms["vb"].detach().numpy(),
          "zero-grad tensors", sum(1 for v in norm if v == 0.0), "of", len(norm))


if __name__ == "__main__":
    import math
    torch.manual_seed(0)
    torch.set_num_threads(8)
    gm = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        encoder_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "diffusion":
        diffusion_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decoder":
        decoder_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "train_masked":
        gpt_train_masked_case(gm)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "generate":
        generate_case(gm)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "full_step":
        vqvae_full_step_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "step":
        vqvae_step_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "text_encoder":
        text_encoder_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "flow":
        flow_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "disc":
        disc_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "kv":
        kv_case(gm)
        sys.exit(0)
    tiny = dict(layers=2, model_dim=128, heads=2, max_text_tokens=40, max_mel_tokens=80)
    gpt_case(gm, "gpt_tiny", tiny, B=2, TL=12, CL=24)
    # ragged: clipping + set_mel_padding paths (wav_lengths//1024+1 < CL for some rows)
    gpt_case(gm, "gpt_ragged", tiny, B=3, TL=16, CL=30, text_lengths=[9, 14, 5], wav_lengths=[20 * 1024 + 17, 27 * 1024, 6 * 1024 + 1000], seed=3)
    if len(sys.argv) > 1 and sys.argv[1] == "encoder":
        encoder_case()
        sys.exit(0)
    gpt_train_masked_case(gm)
    generate_case(gm)
    kv_case(gm)
    decoder_case()
    disc_case()
    flow_case()
    text_encoder_case()
    vqvae_step_case()
    vqvae_full_step_case()
    vq_case()
    mel_case()
    encoder_case()
    diffusion_case()
