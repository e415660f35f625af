"""Host side of the UnifiedVoice GPT engine: flat parameter / gradient buffers, workspace, and the ctypes calls into
`ttts_gpt_forward` / `ttts_gpt_backward` / `ttts_adamw_step` (include/ttts_b200.h).

Reference being replaced: the body of `Trainer.train`'s loop, ttts/gpt/train.py:99-121, and `UnifiedVoice.forward`,
ttts/gpt/model.py:453-510.  PyTorch is used for device memory, streams and torch.distributed only.
"""
import ctypes

import torch

from .. import _lib as L


class GptConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "layers", "model_dim", "heads", "max_text_tokens", "max_mel_tokens", "n_text_vocab", "n_mel_vocab",
        "start_text_token", "stop_text_token", "start_mel_token", "stop_mel_token", "mel_length_compression")]


class GptIO(ctypes.Structure):
    _fields_ = [
        ("cfg", GptConfig),
        ("B", ctypes.c_int32), ("TL", ctypes.c_int32), ("CL", ctypes.c_int32),
        ("text", ctypes.c_void_p), ("ld_text", ctypes.c_int32),
        ("codes", ctypes.c_void_p), ("ld_codes", ctypes.c_int32),
        ("wav_lengths", ctypes.c_void_p),
        ("params", ctypes.c_void_p), ("params16", ctypes.c_void_p), ("grads", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
        ("losses", ctypes.c_void_p),
        ("save_acts", ctypes.c_int32), ("want_latent", ctypes.c_int32),
        ("drop_p", ctypes.c_float), ("seed", ctypes.c_uint64),
        ("gscale_text", ctypes.c_void_p), ("gscale_mel", ctypes.c_void_p),
        ("weight_text", ctypes.c_float), ("weight_mel", ctypes.c_float),
    ]


class GptDecode(ctypes.Structure):
    """ttts_gpt_decode (include/ttts_b200.h): one KV-cache decode step"""
    _fields_ = [
        ("cfg", GptConfig),
        ("B", ctypes.c_int32), ("T_max", ctypes.c_int32), ("text_positions", ctypes.c_int32), ("pos_shift", ctypes.c_int32),
        ("codes", ctypes.c_void_p), ("ld_codes", ctypes.c_int32),
        ("slot", ctypes.c_void_p),
        ("params", ctypes.c_void_p), ("params16", ctypes.c_void_p),
        ("kv", ctypes.c_void_p), ("kv_bytes", ctypes.c_int64),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
        ("logits", ctypes.c_void_p),
    ]


# tensor ids (ttts_gpt_tensor)
(P_TEXT_EMB, P_MEL_EMB, P_TEXT_POS, P_MEL_POS, P_LN1_W, P_LN1_B, P_ATTN_W, P_ATTN_B, P_PROJ_W, P_PROJ_B, P_LN2_W, P_LN2_B,
 P_FC_W, P_FC_B, P_PR_W, P_PR_B, P_LNF_W, P_LNF_B, P_FN_W, P_FN_B, P_TEXT_HEAD_W, P_TEXT_HEAD_B, P_MEL_HEAD_W, P_MEL_HEAD_B) = range(24)
WS_MEL_LOGITS, WS_TEXT_LOGITS, WS_LATENT, WS_RESID, WS_TOKENS = range(5)


def _setup_prototypes(lib):
    if getattr(lib, "_gpt_protos", False):
        return
    lib.ttts_gpt_param_offset.restype = ctypes.c_int64
    lib.ttts_gpt_param_offset.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int32, ctypes.c_int32]
    lib.ttts_gpt_param_numel.restype = ctypes.c_int64
    lib.ttts_gpt_param_numel.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int32]
    lib.ttts_gpt_param_count.restype = ctypes.c_int64
    lib.ttts_gpt_param_count.argtypes = [ctypes.POINTER(GptConfig)]
    lib.ttts_gpt_stage_range.restype = ctypes.c_int32
    lib.ttts_gpt_stage_range.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    lib.ttts_gpt_workspace_bytes.restype = ctypes.c_int64
    lib.ttts_gpt_workspace_bytes.argtypes = [ctypes.POINTER(GptConfig)] + [ctypes.c_int32] * 4
    lib.ttts_gpt_workspace_offset.restype = ctypes.c_int64
    lib.ttts_gpt_workspace_offset.argtypes = [ctypes.POINTER(GptConfig)] + [ctypes.c_int32] * 6
    lib.ttts_gpt_logits_ld.restype = ctypes.c_int32
    lib.ttts_gpt_logits_ld.argtypes = [ctypes.c_int32]
    lib.ttts_gpt_forward.argtypes = [ctypes.POINTER(GptIO), ctypes.c_void_p]
    lib.ttts_gpt_backward.argtypes = [ctypes.POINTER(GptIO), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.ttts_gpt_kv_bytes.restype = ctypes.c_int64
    lib.ttts_gpt_kv_bytes.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int32, ctypes.c_int32]
    lib.ttts_gpt_decode_workspace_bytes.restype = ctypes.c_int64
    lib.ttts_gpt_decode_workspace_bytes.argtypes = [ctypes.POINTER(GptConfig), ctypes.c_int32]
    lib.ttts_gpt_kv_prefill.argtypes = [ctypes.POINTER(GptIO), ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.ttts_gpt_decode_step.argtypes = [ctypes.POINTER(GptDecode), ctypes.c_void_p]
    lib.ttts_cast_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    lib.ttts_gpt_dropout_mask.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p]
    lib.ttts_grad_norm.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ttts_adamw_step.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int64, ctypes.c_void_p] + [ctypes.c_float] * 7 + [ctypes.c_int32, ctypes.c_void_p]
    lib._gpt_protos = True


# reference state_dict name -> (tensor id, shape fn)   
def tensor_table(cfg):
    d, L = cfg.model_dim, cfg.layers
    t = [("text_embedding.weight", P_TEXT_EMB, 0, (cfg.n_text_vocab, d)), ("mel_embedding.weight", P_MEL_EMB, 0, (cfg.n_mel_vocab, d))]
    for i in range(L):
        p = "gpt.h.%d." % i
        t += [(p + "ln_1.weight", P_LN1_W, i, (d,)), (p + "ln_1.bias", P_LN1_B, i, (d,)),
              (p + "attn.c_attn.weight", P_ATTN_W, i, (d, 3 * d)), (p + "attn.c_attn.bias", P_ATTN_B, i, (3 * d,)),
              (p + "attn.c_proj.weight", P_PROJ_W, i, (d, d)), (p + "attn.c_proj.bias", P_PROJ_B, i, (d,)),
              (p + "ln_2.weight", P_LN2_W, i, (d,)), (p + "ln_2.bias", P_LN2_B, i, (d,)),
              (p + "mlp.c_fc.weight", P_FC_W, i, (d, 4 * d)), (p + "mlp.c_fc.bias", P_FC_B, i, (4 * d,)),
              (p + "mlp.c_proj.weight", P_PR_W, i, (4 * d, d)), (p + "mlp.c_proj.bias", P_PR_B, i, (d,))]
    t += [("gpt.ln_f.weight", P_LNF_W, 0, (d,)), ("gpt.ln_f.bias", P_LNF_B, 0, (d,)),
          ("mel_pos_embedding.emb.weight", P_MEL_POS, 0, (cfg.max_mel_tokens + 2, d)),
          ("text_pos_embedding.emb.weight", P_TEXT_POS, 0, (cfg.max_text_tokens + 2, d)),
          ("final_norm.weight", P_FN_W, 0, (d,)), ("final_norm.bias", P_FN_B, 0, (d,)),
          ("text_head.weight", P_TEXT_HEAD_W, 0, (cfg.n_text_vocab, d)), ("text_head.bias", P_TEXT_HEAD_B, 0, (cfg.n_text_vocab,)),
          ("mel_head.weight", P_MEL_HEAD_W, 0, (cfg.n_mel_vocab, d)), ("mel_head.bias", P_MEL_HEAD_B, 0, (cfg.n_mel_vocab,))]
    return t


class Layout:
    """Offsets of every reference tensor inside the flat fp32 / bf16 / grad buffers."""

    def __init__(self, cfg):
        lib = L.lib()
        _setup_prototypes(lib)
        self.cfg = cfg
        self.total = lib.ttts_gpt_param_count(ctypes.byref(cfg))
        if self.total <= 0:
            raise L.TTTSError("bad GPT config: " + lib.ttts_last_error().decode())
        self.entries = []
        for name, tid, layer, shape in tensor_table(cfg):
            off = lib.ttts_gpt_param_offset(ctypes.byref(cfg), tid, layer)
            n = 1
            for s in shape:
                n *= s
            assert off >= 0 and n == lib.ttts_gpt_param_numel(ctypes.byref(cfg), tid), name
            self.entries.append((name, off, n, shape))

    def stage_range(self, stage):
        b, e = ctypes.c_int64(), ctypes.c_int64()
        L.check(L.lib().ttts_gpt_stage_range(ctypes.byref(self.cfg), stage, ctypes.byref(b), ctypes.byref(e)), "ttts_gpt_stage_range")
        return b.value, e.value

    def views(self, flat):
        return {name: flat[off:off + n].view(shape) for name, off, n, shape in self.entries}


class Engine:
    """Owns workspace + bf16 shadow + losses for one flat parameter buffer on one CUDA device."""

    def __init__(self, cfg, flat_params):
        lib = L.lib()
        _setup_prototypes(lib)
        L.require_cuda(flat_params)
        if not lib.ttts_device_ok():
            raise L.TTTSError("ttts_b200 needs an sm_100 (B200) device; no fallback path exists")
        self.cfg = cfg
        self.layout = Layout(cfg)
        self.flat = flat_params
        assert flat_params.dtype == torch.float32 and flat_params.numel() == self.layout.total
        dev = flat_params.device
        self.device = dev
        self.flat16 = torch.empty(self.layout.total, dtype=torch.bfloat16, device=dev)
        self.grads = torch.zeros(self.layout.total, dtype=torch.float32, device=dev)
        self.losses = torch.zeros(2, dtype=torch.float32, device=dev)
        self.ws = None
        self._shadow_version = -1
        # version key of the fp32 master: the flat buffer's counter PLUS every parameter's own counter -- `p.data = view` (UnifiedVoice._apply)
        # gives each Parameter its own version counter, so load_state_dict / in-place edits through a parameter do not move flat._version.
        # The owner (UnifiedVoice) installs `version_fn`; stand-alone engines see only the flat buffer.
        self.version_fn = None
        self.trust_version = False   # True: only re-cast the bf16 shadow when the version key moved (fused trainer)
        self._last = None      # (B, TL, CL, save, drop_p, seed, text, codes, wav) of the last forward with save_acts
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(2048, dtype=torch.float32, device=dev)
        self.exp_avg = None
        self.exp_avg_sq = None

    # ---- bf16 shadow ----
    def _version(self):
        return self.flat._version if self.version_fn is None else self.flat._version + self.version_fn()

    def invalidate_shadow(self):
        """the fp32 master changed behind the engine's back (checkpoint load, external optimizer): re-cast at the next use"""
        self._shadow_version = -1

    def refresh_shadow(self, force=False):
        v = self._version()
        if force or v != self._shadow_version:
            L.check(L.lib().ttts_cast_bf16(self.flat.data_ptr(), self.flat16.data_ptr(), self.layout.total, L.stream_ptr().value), "ttts_cast_bf16")
            self._shadow_version = v

    # ---- workspace ----
    def _workspace(self, B, TL, CL, save):
        need = L.lib().ttts_gpt_workspace_bytes(ctypes.byref(self.cfg), B, TL, CL, int(save))
        if need <= 0:
            raise L.TTTSError("workspace query failed: " + L.lib().ttts_last_error().decode())
        if self.ws is None or self.ws.numel() < need:
            self.ws = None
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return need

    def ws_view(self, item, B, TL, CL, save, dtype, shape, layer=0):
        off = L.lib().ttts_gpt_workspace_offset(ctypes.byref(self.cfg), B, TL, CL, int(save), item, layer)
        assert off >= 0
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        return self.ws[off:off + nbytes].view(dtype).view(shape)

    def _io(self, B, TL, CL, text, codes, wav_lengths, save, want_latent, drop_p, seed):
        io = GptIO()
        io.cfg = self.cfg
        io.B, io.TL, io.CL = B, TL, CL
        io.text, io.ld_text = text.data_ptr(), text.stride(0)
        io.codes, io.ld_codes = codes.data_ptr(), codes.stride(0)
        io.wav_lengths = wav_lengths.data_ptr()
        io.params, io.params16, io.grads = self.flat.data_ptr(), self.flat16.data_ptr(), self.grads.data_ptr()
        io.workspace, io.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        io.losses = self.losses.data_ptr()
        io.save_acts, io.want_latent = int(save), int(want_latent)
        io.drop_p, io.seed = float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF
        io.weight_text = io.weight_mel = 1.0
        return io

    def forward(self, text, codes, wav_lengths, TL, CL, save=True, want_latent=False, drop_p=0.0, seed=0):
        """text [B,>=TL] int64, codes [B,>=CL] int64 (mutated in place like the reference), wav_lengths [B] int64."""
        L.require_cuda(text, codes, wav_lengths)
        assert text.dtype == torch.int64 and codes.dtype == torch.int64 and wav_lengths.dtype == torch.int64
        assert text.stride(1) == 1 and codes.stride(1) == 1 and wav_lengths.is_contiguous()
        B = text.shape[0]
        self.refresh_shadow()
        self._workspace(B, TL, CL, save)
        io = self._io(B, TL, CL, text, codes, wav_lengths, save, want_latent, drop_p, seed)
        L.check(L.lib().ttts_gpt_forward(ctypes.byref(io), L.stream_ptr().value), "ttts_gpt_forward")
        if save:
            self._last = (B, TL, CL, drop_p, seed, text, codes, wav_lengths)
        return io

    def backward(self, gscale_text=None, gscale_mel=None, weight_text=1.0, weight_mel=1.0, stage_begin=0, stage_end=None):
        assert self._last is not None, "backward() without a saved forward"
        B, TL, CL, drop_p, seed, text, codes, wav = self._last
        io = self._io(B, TL, CL, text, codes, wav, True, False, drop_p, seed)
        io.gscale_text = gscale_text.data_ptr() if gscale_text is not None else None
        io.gscale_mel = gscale_mel.data_ptr() if gscale_mel is not None else None
        io.weight_text, io.weight_mel = float(weight_text), float(weight_mel)
        if stage_end is None:
            stage_end = self.cfg.layers + 2
        L.check(L.lib().ttts_gpt_backward(ctypes.byref(io), stage_begin, stage_end, L.stream_ptr().value), "ttts_gpt_backward")

    # ---- KV-cache decode (inference_speech with kv_cache=True) ----
    def decode_setup(self, B, T_max):
        """Allocate (or reuse) the cache [layers, 2, B, heads, T_max, 64] bf16, the decode workspace, the device slot counter and the logits."""
        lib = L.lib()
        st = getattr(self, "_dec", None)
        if st is not None and st["B"] == B and st["T_max"] >= T_max:
            return st
        T_max = (T_max + 63) // 64 * 64
        kv_bytes = lib.ttts_gpt_kv_bytes(ctypes.byref(self.cfg), B, T_max)
        ws_bytes = lib.ttts_gpt_decode_workspace_bytes(ctypes.byref(self.cfg), B)
        if kv_bytes <= 0 or ws_bytes <= 0:
            raise L.TTTSError("decode size query failed: " + lib.ttts_last_error().decode())
        self._dec = None
        st = dict(B=B, T_max=T_max,
                  kv=torch.empty(kv_bytes, dtype=torch.uint8, device=self.device),
                  ws=torch.empty(ws_bytes, dtype=torch.uint8, device=self.device),
                  slot=torch.zeros(1, dtype=torch.int32, device=self.device),
                  logits=torch.zeros(B, self.cfg.n_mel_vocab, dtype=torch.float32, device=self.device),
                  graph=None, graph_key=None)
        self._dec = st
        return st

    def kv_prefill(self, io, n_pos):
        """Copy K / V of positions [0, n_pos) of every layer out of the workspace of the forward(save=True) that produced `io`."""
        st = self._dec
        L.check(L.lib().ttts_gpt_kv_prefill(ctypes.byref(io), st["kv"].data_ptr(), st["kv"].numel(), st["T_max"], int(n_pos), L.stream_ptr().value),
                "ttts_gpt_kv_prefill")
        st["slot"].fill_(int(n_pos))

    def decode_step(self, codes, text_positions, pos_shift=0, graph=False):
        """Feed codes[:, slot - text_positions - 1] (slot lives on the device and is advanced by the step); returns the fp32 logits buffer [B, V]."""
        st = self._dec
        assert codes.dtype == torch.int64 and codes.stride(1) == 1 and codes.shape[0] == st["B"]
        a = GptDecode()
        a.cfg = self.cfg
        a.B, a.T_max, a.text_positions, a.pos_shift = st["B"], st["T_max"], int(text_positions), int(pos_shift)
        a.codes, a.ld_codes = codes.data_ptr(), codes.stride(0)
        a.slot = st["slot"].data_ptr()
        a.params, a.params16 = self.flat.data_ptr(), self.flat16.data_ptr()
        a.kv, a.kv_bytes = st["kv"].data_ptr(), st["kv"].numel()
        a.workspace, a.workspace_bytes = st["ws"].data_ptr(), st["ws"].numel()
        a.logits = st["logits"].data_ptr()
        if not graph:
            L.check(L.lib().ttts_gpt_decode_step(ctypes.byref(a), L.stream_ptr().value), "ttts_gpt_decode_step")
            return st["logits"]
        # every kernel of the step reads the slot from device memory, so ONE captured step replays for every position
        key = (codes.data_ptr(), codes.stride(0), int(text_positions), int(pos_shift))
        if st["graph"] is None or st["graph_key"] != key:
            g = torch.cuda.CUDAGraph()
            slot0 = st["slot"].clone()
            with torch.cuda.graph(g):
                L.check(L.lib().ttts_gpt_decode_step(ctypes.byref(a), L.stream_ptr().value), "ttts_gpt_decode_step (capture)")
            st["slot"].copy_(slot0)                  # capture does not execute; keep the counter where it was
            st["graph"], st["graph_key"] = g, key
        st["graph"].replay()
        return st["logits"]

    def dropout_masks(self, drop_p=None, seed=None, B=None, T=None):
        """The keep masks the last forward(save=True) drew (or those of an explicit (drop_p, seed, B, T)), from ttts_gpt_dropout_mask:
        {"embd", "attn_p<l>", "attn_o<l>", "mlp_o<l>"} as uint8 tensors, plus "scale" = 32768 / (32768 - round(p * 32768))."""
        if drop_p is None:
            B, TL, CL, drop_p, seed = self._last[:5]
            T = TL + CL + 4
        lib = L.lib()
        d, H, nl = self.cfg.model_dim, self.cfg.heads, self.cfg.layers
        out = {"scale": 1.0 / (1.0 - int(drop_p * 32768.0 + 0.5) / 32768.0)}

        def one(site, layer, rows, cols, shape):
            m = torch.empty(shape, dtype=torch.uint8, device=self.device)
            L.check(lib.ttts_gpt_dropout_mask(m.data_ptr(), site, layer, rows, cols, float(drop_p), int(seed) & 0xFFFFFFFFFFFFFFFF, L.stream_ptr().value),
                    "ttts_gpt_dropout_mask")
            return m
        out["embd"] = one(0, 0, B * T, d, (B, T, d))
        for l in range(nl):
            out["attn_p%d" % l] = one(1, l, B * H, T, (B, H, T, T))
            out["attn_o%d" % l] = one(2, l, B * T, d, (B, T, d))
            out["mlp_o%d" % l] = one(3, l, B * T, d, (B, T, d))
        return out

    # ---- step tail ----
    def grad_norm(self):
        L.check(L.lib().ttts_grad_norm(self.grads.data_ptr(), self.layout.total, self._scratch.data_ptr(), self.norm.data_ptr(),
                                       L.stream_ptr().value), "ttts_grad_norm")
        return self.norm

    def adamw(self, lr, step, betas=(0.9, 0.96), eps=1e-8, weight_decay=0.01, max_norm=1.0, grad_scale=1.0, use_norm=True):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
        norm_ptr = self.norm.data_ptr() if use_norm else None
        L.check(L.lib().ttts_adamw_step(self.flat.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        self.flat16.data_ptr(), self.layout.total, norm_ptr, float(max_norm), float(grad_scale), float(lr),
                                        float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), L.stream_ptr().value),
                "ttts_adamw_step")
        self._shadow_version = self._version()   # kernel wrote the bf16 shadow itself
