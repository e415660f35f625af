"""`UnifiedVoice` -- drop-in for ttts/gpt/model.py:292-510 whose arithmetic runs in hand-written sm_100a CUDA.

Same constructor kwargs (ttts/gpt/model.py:293-297; called as `UnifiedVoice(**cfg['gpt'])`, ttts/gpt/train.py:45), same
`forward` signature and returns `(loss_text, loss_mel, mel_logits[B,1026,CL+2])` (or latents with `return_latent`), same
`state_dict()` keys / shapes so checkpoints interchange, ordinary `nn.Parameter`s so `AdamW(model.parameters())`,
`clip_grad_norm_` and `loss.backward()` work unchanged.  Every parameter is a view into ONE flat fp32 buffer; the gradients
autograd receives are views into one flat fp32 gradient buffer (a single NCCL all-reduce under data parallelism).

No CPU fallback: calling forward on a non-CUDA / non-sm_100 device raises.  Inference-only paths of the reference that are not on
the training hot path (`text_first=False`, `raw_mels`, `return_attentions`, `inference_speech`) raise NotImplementedError.
"""
import math
import os

import torch
import torch.nn as nn

from .. import _lib as L
from . import engine as E


class _Holder(nn.Module):
    """A module that only owns parameters (keeps the reference's state_dict nesting)."""

    def __init__(self, **params):
        super().__init__()
        for k, v in params.items():
            if isinstance(v, nn.Module):
                self.add_module(k, v)
            else:
                self.register_parameter(k, v)


class _GPTStepFn(torch.autograd.Function):
    """forward = ttts_gpt_forward, backward = ttts_gpt_backward; parameters enter as inputs so autograd routes
    their gradients (views of the flat gradient buffer) to `.grad` / DDP hooks."""

    @staticmethod
    def forward(ctx, model, text, codes, wav_lengths, TL, CL, drop_p, seed, need_grad, *params):
        eng = model._engine()
        eng.refresh_shadow(force=not eng.trust_version)
        eng.forward(text, codes, wav_lengths, TL, CL, save=need_grad, drop_p=drop_p, seed=seed)
        B = text.shape[0]
        ld = L.lib().ttts_gpt_logits_ld(model.number_mel_codes)
        logits = eng.ws_view(E.WS_MEL_LOGITS, B, TL, CL, need_grad, torch.bfloat16, (B, CL + 2, ld))
        mel_logits = logits[:, :, :model.number_mel_codes].clone()
        losses = eng.losses.clone()
        ctx.model = model
        ctx.need_grad = need_grad
        ctx.nparams = len(params)
        ctx.mark_non_differentiable(mel_logits)
        return losses[0], losses[1], mel_logits

    @staticmethod
    def backward(ctx, g_text, g_mel, _g_logits):
        model = ctx.model
        eng = model._engine()
        params = [model._params_by_name[n] for n in model._param_names]
        gviews = model._grad_views()
        # If autograd previously stole our views as .grad (zero_grad(set_to_none=True) path), they alias the flat
        # gradient buffer: accumulate in place.  Otherwise start from zero and hand the views to autograd.
        aliased = [p.grad is not None and p.grad.data_ptr() == gv.data_ptr() for p, gv in zip(params, gviews)]
        in_place = all(aliased)
        if not in_place:
            for p, a in zip(params, aliased):
                if a:
                    p.grad = p.grad.clone()
            eng.grads.zero_()
        gt = g_text.contiguous().float() if g_text is not None else torch.zeros((), device=eng.device)
        gm = g_mel.contiguous().float() if g_mel is not None else torch.zeros((), device=eng.device)
        eng.backward(gscale_text=gt, gscale_mel=gm)
        model._after_backward(in_place)
        if in_place:
            return (None,) * (9 + ctx.nparams)
        return (None,) * 9 + tuple(gviews)


class UnifiedVoice(nn.Module):
    def __init__(self, layers=8, model_dim=512, heads=8, max_text_tokens=120, max_mel_tokens=250, max_conditioning_inputs=1,
                 mel_length_compression=1024, number_text_tokens=256, start_text_token=None, number_mel_codes=8194,
                 start_mel_token=8192, stop_mel_token=8193, train_solo_embeddings=False, use_mel_codes_as_input=True,
                 checkpointing=True, types=1):
        super().__init__()
        self._ctor_kwargs = dict(layers=layers, model_dim=model_dim, heads=heads, max_text_tokens=max_text_tokens,
                                 max_mel_tokens=max_mel_tokens, max_conditioning_inputs=max_conditioning_inputs,
                                 mel_length_compression=mel_length_compression, number_text_tokens=number_text_tokens,
                                 start_text_token=start_text_token, number_mel_codes=number_mel_codes, start_mel_token=start_mel_token,
                                 stop_mel_token=stop_mel_token, train_solo_embeddings=train_solo_embeddings,
                                 use_mel_codes_as_input=use_mel_codes_as_input, checkpointing=checkpointing, types=types)
        if not use_mel_codes_as_input:
            raise NotImplementedError("use_mel_codes_as_input=False (MelEncoder input) is not on the hot path")
        if train_solo_embeddings:
            raise NotImplementedError("train_solo_embeddings=True is not on the hot path")
        self.number_text_tokens = number_text_tokens
        self.start_text_token = number_text_tokens * types if start_text_token is None else start_text_token
        self.stop_text_token = 0
        self.number_mel_codes = number_mel_codes
        self.start_mel_token = start_mel_token
        self.stop_mel_token = stop_mel_token
        self.layers = layers
        self.heads = heads
        self.max_mel_tokens = max_mel_tokens
        self.max_text_tokens = max_text_tokens
        self.model_dim = model_dim
        self.max_conditioning_inputs = max_conditioning_inputs
        self.mel_length_compression = mel_length_compression
        self.checkpointing = checkpointing      # accepted for API parity; activations fit in HBM, nothing is recomputed
        self.mel_solo_embedding = 0
        self.text_solo_embedding = 0
        # dropout probability of the four GPT-2 sites (HF GPT2Config defaults embd/attn/resid_pdrop = 0.1; the reference
        # exposes no knob).  Active only in train() mode, like nn.Dropout.
        self.dropout_p = 0.1
        self._seed_counter = 0
        self._base_seed = 0x5DEECE66D

        cfg = E.GptConfig()
        cfg.layers, cfg.model_dim, cfg.heads = layers, model_dim, heads
        cfg.max_text_tokens, cfg.max_mel_tokens = max_text_tokens, max_mel_tokens
        cfg.n_text_vocab, cfg.n_mel_vocab = number_text_tokens * types + 1, number_mel_codes
        cfg.start_text_token, cfg.stop_text_token = self.start_text_token, self.stop_text_token
        cfg.start_mel_token, cfg.stop_mel_token = start_mel_token, stop_mel_token
        cfg.mel_length_compression = mel_length_compression
        self._cfg = cfg
        self._layout = E.Layout(cfg)
        self._flat = torch.zeros(self._layout.total, dtype=torch.float32)
        self._eng = None
        self._gviews = None

        views = self._layout.views(self._flat)
        P = {name: nn.Parameter(v) for name, v in views.items()}
        self._param_names = [name for name, _, _, _ in self._layout.entries]
        # module tree with the reference's state_dict nesting / registration order
        self.text_embedding = _Holder(weight=P["text_embedding.weight"])
        self.mel_embedding = _Holder(weight=P["mel_embedding.weight"])
        blocks = []
        for i in range(layers):
            p = "gpt.h.%d." % i
            blocks.append(_Holder(
                ln_1=_Holder(weight=P[p + "ln_1.weight"], bias=P[p + "ln_1.bias"]),
                attn=_Holder(c_attn=_Holder(weight=P[p + "attn.c_attn.weight"], bias=P[p + "attn.c_attn.bias"]),
                             c_proj=_Holder(weight=P[p + "attn.c_proj.weight"], bias=P[p + "attn.c_proj.bias"])),
                ln_2=_Holder(weight=P[p + "ln_2.weight"], bias=P[p + "ln_2.bias"]),
                mlp=_Holder(c_fc=_Holder(weight=P[p + "mlp.c_fc.weight"], bias=P[p + "mlp.c_fc.bias"]),
                            c_proj=_Holder(weight=P[p + "mlp.c_proj.weight"], bias=P[p + "mlp.c_proj.bias"]))))
        self.gpt = _Holder(h=nn.ModuleList(blocks), ln_f=_Holder(weight=P["gpt.ln_f.weight"], bias=P["gpt.ln_f.bias"]))
        self.mel_pos_embedding = _Holder(emb=_Holder(weight=P["mel_pos_embedding.emb.weight"]))
        self.text_pos_embedding = _Holder(emb=_Holder(weight=P["text_pos_embedding.emb.weight"]))
        self.final_norm = _Holder(weight=P["final_norm.weight"], bias=P["final_norm.bias"])
        self.text_head = _Holder(weight=P["text_head.weight"], bias=P["text_head.bias"])
        self.mel_head = _Holder(weight=P["mel_head.weight"], bias=P["mel_head.bias"])
        self._params_by_name = P
        self.reset_parameters()

    # ------------------------------------------------------------------ init (reference statistics)
    @torch.no_grad()
    def reset_parameters(self):
        """Embeddings N(0,.02) (ttts/gpt/model.py:235,356); GPT-2 block init (HF _init_weights: N(0,.02), c_proj
        N(0,.02/sqrt(2L)), LN 1/0, biases 0); heads / final_norm torch defaults (nn.Linear kaiming-uniform, LN 1/0)."""
        L_ = self.layers
        for name, p in self._params_by_name.items():
            if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name in ("gpt.ln_f.weight", "final_norm.weight"):
                p.fill_(1.0)
            elif name in ("text_head.weight", "mel_head.weight"):
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
            elif name in ("text_head.bias", "mel_head.bias"):
                bound = 1.0 / math.sqrt(self.model_dim)
                p.uniform_(-bound, bound)
            elif name.endswith(".bias"):
                p.zero_()
            elif name.endswith("c_proj.weight"):
                p.normal_(0.0, 0.02 / math.sqrt(2 * L_))
            else:
                p.normal_(0.0, 0.02)

    # ------------------------------------------------------------------ flat-buffer plumbing
    def _apply(self, fn, recurse=True):
        """`.to()/.cuda()` move the ONE flat buffer and re-point every parameter at its slice."""
        new_flat = fn(self._flat)
        if new_flat.dtype != torch.float32:
            raise L.TTTSError("UnifiedVoice keeps fp32 master parameters; mixed precision is handled inside the kernels")
        if new_flat is not self._flat:
            self._flat = new_flat
            for name, v in self._layout.views(self._flat).items():
                p = self._params_by_name[name]
                p.data = v
                p.grad = None
            self._eng = None
            self._gviews = None
        return self

    def __deepcopy__(self, memo):
        # parameters are views of one flat buffer: rebuild rather than copying tensor by tensor (ema copy, train.py:64-69)
        new = UnifiedVoice(**self._ctor_kwargs)
        new.to(self._flat.device)
        with torch.no_grad():
            new._flat.copy_(self._flat)
        for pn, po in zip(new.parameters(), self.parameters()):
            pn.requires_grad_(po.requires_grad)
        new.train(self.training)
        new.dropout_p = self.dropout_p
        return new

    def _engine(self):
        if self._eng is None:
            if not self._flat.is_cuda:
                raise L.TTTSError("UnifiedVoice runs on sm_100a only: move the module to a CUDA device (no CPU fallback)")
            with torch.cuda.device(self._flat.device):
                self._eng = E.Engine(self._cfg, self._flat)
            plist = list(self._params_by_name.values())
            self._eng.version_fn = lambda: sum(p._version for p in plist)
        return self._eng

    def _grad_views(self):
        if self._gviews is None:
            gv = self._layout.views(self._engine().grads)
            self._gviews = [gv[n] for n in self._param_names]
        return self._gviews

    def enable_flat_allreduce(self, process_group=None, enabled=True):
        """Data parallelism WITHOUT a DistributedDataParallel wrapper: at the end of every backward the flat gradient buffer is averaged
        across the process group with ONE all-reduce (NCCL over NVLink / NVSwitch) -- what SURVEY.md 8b asks for in place of DDP's 25 MB
        buckets (K10).  Under `accelerator.prepare` / DDP leave it off: the wrapper's reducer already averages the (same) gradient views.
        Gradient accumulation: like the reference's loop (no `no_sync`), every backward reduces."""
        import torch.distributed as dist
        if enabled and not (dist.is_available() and dist.is_initialized()):
            raise L.TTTSError("enable_flat_allreduce needs an initialised torch.distributed process group")
        self._flat_allreduce = (process_group if process_group is not None else True) if enabled else None

    def _after_backward(self, accumulated=False):
        pg = getattr(self, "_flat_allreduce", None)
        if pg is None:
            return
        if accumulated:
            # .grad already held (reduced) gradients that this backward added to: reducing the buffer again would average them twice
            raise L.TTTSError("enable_flat_allreduce: gradients were accumulated across backward calls; zero_grad(set_to_none=True) between "
                              "steps, or use ttts_b200.gpt.train.FusedStep(accumulate=n), which reduces once per optimizer step")
        import torch.distributed as dist
        group = None if pg is True else pg
        world = dist.get_world_size(group)
        if world > 1:
            g = self._engine().grads
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
            g.mul_(1.0 / world)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        # older HF versions stored causal-mask buffers in the checkpoint; they carry no information
        sd = {k: v for k, v in state_dict.items() if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias"))}
        out = super().load_state_dict(sd, strict=strict)
        if self._eng is not None:
            self._eng.invalidate_shadow()       # the bf16 shadow the kernels read is re-cast from the loaded fp32 master at the next call
        return out

    # ------------------------------------------------------------------ reference helpers kept for API parity
    def build_aligned_inputs_and_targets(self, input, start_token, stop_token):
        inp = torch.nn.functional.pad(input, (1, 0), value=start_token)
        tar = torch.nn.functional.pad(input, (0, 1), value=stop_token)
        return inp, tar

    def set_mel_padding(self, mel_input_tokens, wav_lengths):
        """ttts/gpt/model.py:402-414 (vectorised; in place on the caller's tensor like the reference)."""
        mel_lengths = torch.div(wav_lengths, self.mel_length_compression, rounding_mode="trunc")
        pos = torch.arange(mel_input_tokens.shape[-1], device=mel_input_tokens.device)[None, :]
        mel_input_tokens.masked_fill_(pos >= (mel_lengths[:, None] + 1), self.stop_mel_token)
        return mel_input_tokens

    # ------------------------------------------------------------------ forward
    def forward(self, text_inputs, text_lengths, mel_codes, wav_lengths, types=None, text_first=True, raw_mels=None,
                return_attentions=False, return_latent=False, clip_inputs=True):
        if not text_first or raw_mels is not None or return_attentions:
            raise NotImplementedError("only the text_first / mel-code training path of the reference is implemented")
        L.require_cuda(text_inputs, text_lengths, mel_codes, wav_lengths)
        if types is not None:
            text_inputs = text_inputs * (1 + types).unsqueeze(-1)
        TL, CL = text_inputs.shape[1], mel_codes.shape[1]
        if clip_inputs:
            # same host sync as the reference (ttts/gpt/model.py:477-480)
            TL = min(TL, int(text_lengths.max()))
            CL = min(CL, int(wav_lengths.max()) // self.mel_length_compression)
        if TL + 2 > self.max_text_tokens + 2 or CL + 2 > self.max_mel_tokens + 2:
            raise IndexError("sequence longer than the position tables (max_text_tokens=%d, max_mel_tokens=%d)"
                             % (self.max_text_tokens, self.max_mel_tokens))
        text_inputs = text_inputs if (text_inputs.dtype == torch.int64 and text_inputs.stride(-1) == 1) else text_inputs.long().contiguous()
        if mel_codes.dtype != torch.int64 or mel_codes.stride(-1) != 1:
            raise TypeError("mel_codes must be a contiguous int64 tensor (it is padded in place, like the reference)")
        wav_lengths = wav_lengths.long().contiguous()
        eng = self._engine()
        if return_latent:
            with torch.no_grad():
                eng.refresh_shadow(force=not eng.trust_version)
                eng.forward(text_inputs, mel_codes, wav_lengths, TL, CL, save=False, want_latent=True)
                B = text_inputs.shape[0]
                T = TL + CL + 4
                lat = eng.ws_view(E.WS_LATENT, B, TL, CL, False, torch.float32, (B, T, self.model_dim))
                return lat[:, -(CL + 2):][:, :-2].clone()
        drop_p = self.dropout_p if self.training else 0.0
        self._seed_counter += 1
        seed = (self._base_seed * 6364136223846793005 + self._seed_counter * 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        params = [self._params_by_name[n] for n in self._param_names]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        loss_text, loss_mel, mel_logits = _GPTStepFn.apply(self, text_inputs, mel_codes, wav_lengths, TL, CL, drop_p, seed, need_grad, *params)
        return loss_text, loss_mel, mel_logits.permute(0, 2, 1)

    def post_init_gpt2_config(self, use_deepspeed=False, kv_cache=False, half=False, kv_positions=None):
        """ttts/gpt/model.py:357-394 builds a HF GPT2InferenceModel around the trained modules; here generation runs on the same engine
        as training, so there is nothing to build.  `kv_cache=False` (what ttts/api_zh.py:51 uses): every step re-runs the training forward
        over the sequence so far.  `kv_cache=True`: one-token decode steps against cached keys / values (`ttts_gpt_decode_step`).  The
        reference's cached branch indexes the position table one row later than its uncached branch does for the same token
        (model.py:144-147 vs :134-142; pinned by tests/golden/gpt_kvstep.npz), so the two settings generate different codes THERE;
        `kv_positions="reference"` (default with kv_cache=True) keeps that rule, `kv_positions="uncached"` runs the cached kernels with the
        uncached branch's positions, i.e. the kv_cache=False results at a fraction of the cost."""
        if use_deepspeed or half:
            raise NotImplementedError("deepspeed / fp16 inference wrappers are not part of this path (the engine computes in bf16 already)")
        if kv_positions not in (None, "reference", "uncached"):
            raise ValueError("kv_positions must be 'reference' or 'uncached'")
        self._kv_cache = bool(kv_cache)
        self._kv_pos_shift = 1 if (kv_cache and kv_positions in (None, "reference")) else 0
        self.eval()

    @torch.no_grad()
    def inference_speech(self, text_inputs, mel_codes, input_tokens=None, num_return_sequences=1, max_generate_length=None,
                         typical_sampling=False, typical_mass=.9, generator=None, **hf_generate_kwargs):
        """Autoregressive code generation, ttts/gpt/model.py:533-562 with kv_cache=False (the configuration ttts/api_zh.py:51 uses).

        Prompt = [start, text..., stop] followed by [start_mel, conditioning codes...]; every step runs the SAME forward as training
        (`ttts_gpt_forward`, no activations saved) over the sequence so far and reads the mel logits at the last token -- exactly what the
        reference's uncached `GPT2InferenceModel.forward` recomputes per step (model.py:132-142, 159-171).  Token selection follows HF
        `generate` (ttts_b200/gpt/sampling.py).  Returns the generated ids [B * num_return_sequences, n] after the prompt, padded with
        stop_mel_token once a sequence has emitted it (pad_token_id = eos_token_id = stop_mel_token, model.py:555-556)."""
        from . import sampling as S
        kw = dict(hf_generate_kwargs)
        do_sample = bool(kw.pop("do_sample", False))
        top_p, top_k = kw.pop("top_p", None), kw.pop("top_k", None)
        temperature = kw.pop("temperature", 1.0)
        repetition_penalty = kw.pop("repetition_penalty", 1.0)
        kw.pop("length_penalty", None)                               # beam search only
        if kw.pop("num_beams", 1) != 1:
            raise NotImplementedError("beam search is not used by the reference's call (ttts/api_zh.py:78-86) and is not implemented")
        if not do_sample and num_return_sequences != 1:              # HF GenerationConfig.validate
            raise ValueError("Greedy methods (do_sample != True) without beam search do not support `num_return_sequences` different than 1 "
                             "(got %d)." % num_return_sequences)
        if kw:
            raise TypeError("unsupported generate() arguments: %s" % sorted(kw))
        L.require_cuda(text_inputs, mel_codes)
        dev = text_inputs.device
        text = text_inputs.long()
        cond = mel_codes.long()
        if input_tokens is not None:
            # model.py:546-550 repeats the prompt itself AND passes num_return_sequences on to generate(), which expands once more
            assert num_return_sequences % input_tokens.shape[0] == 0, "The number of return sequences must be divisible by the number of input sequences"
            text = text.repeat(num_return_sequences, 1)
            cond = cond.repeat(num_return_sequences, 1)
            cond = torch.cat([cond, input_tokens.long().repeat(num_return_sequences // input_tokens.shape[0], 1)], dim=1)
        if num_return_sequences > 1:
            text = text.repeat_interleave(num_return_sequences, 0)   # HF expands each prompt num_return_sequences times, neighbours together
            cond = cond.repeat_interleave(num_return_sequences, 0)
        B, TL, m = text.shape[0], text.shape[1], cond.shape[1]
        trunc_index = (TL + 2) + (mel_codes.shape[1] + 1)            # fake text ids + [start_mel, conditioning codes]
        max_length = trunc_index + (self.max_mel_tokens - 1 if max_generate_length is None else max_generate_length)
        n_max = max_length - (TL + 2) - 1                            # codes after start_mel the sequence may hold
        n_max = min(n_max, self.max_mel_tokens + 1)                  # the position table ends there (the reference would raise an index error)
        if TL + 2 > self.max_text_tokens + 2 or m > self.max_mel_tokens:
            raise IndexError("prompt longer than the position tables (max_text_tokens=%d, max_mel_tokens=%d)" % (self.max_text_tokens, self.max_mel_tokens))
        codes = torch.full((B, max(n_max, m) + 1), self.stop_mel_token, dtype=torch.int64, device=dev)
        codes[:, :m] = cond
        text = text.contiguous()
        eng = self._engine()
        eng.refresh_shadow(force=not eng.trust_version)
        ld = L.lib().ttts_gpt_logits_ld(self.number_mel_codes)

        def step_logits(n):
            wav = torch.full((B,), (n + 1) * self.mel_length_compression, dtype=torch.int64, device=dev)      # nothing is stop-padded
            eng.forward(text, codes, wav, TL, n, save=False)
            logits = eng.ws_view(E.WS_MEL_LOGITS, B, TL, n, False, torch.bfloat16, (B, n + 2, ld))
            return logits[:, n, :self.number_mel_codes].float()      # position of the last real token (mel_in[n])

        if getattr(self, "_kv_cache", False):
            # cached decoding: the prompt [start, text, stop | start_mel, conditioning codes] goes through the training forward once (its
            # per-layer c_attn outputs fill the cache), every later code is ONE decode step.  Cache slot of code j = TL + 3 + j.
            use_graph = os.environ.get("TTTS_DECODE_GRAPH", "0") == "1"
            eng.decode_setup(B, TL + 3 + max(n_max, m) + 1)
            shift = self._kv_pos_shift

            def step_logits(n):                                      # noqa: F811  (replaces the uncached closure above)
                if n == m:
                    wav = torch.full((B,), (m + 1) * self.mel_length_compression, dtype=torch.int64, device=dev)
                    io = eng.forward(text, codes, wav, TL, m, save=True)
                    eng.kv_prefill(io, TL + 3 + m)                   # text slots + start_mel + m codes
                    logits = eng.ws_view(E.WS_MEL_LOGITS, B, TL, m, True, torch.bfloat16, (B, m + 2, ld))
                    return logits[:, m, :self.number_mel_codes].float()
                return eng.decode_step(codes, TL + 2, shift, graph=use_graph).clone()      # feeds codes[:, n - 1]

        n = S.generate_codes(step_logits, codes, m, n_max, TL + 2, self.start_mel_token, self.stop_mel_token, do_sample, temperature, top_k, top_p,
                             repetition_penalty, typical_sampling, typical_mass, generator)
        return codes[:, mel_codes.shape[1]:n].clone()                # gen[:, trunc_index:] (input_tokens, if any, included -- as in the reference)
