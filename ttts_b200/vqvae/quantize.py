"""`ResidualVectorQuantizer` / `ResidualVectorQuantization` / `VectorQuantization` / `EuclideanCodebook` -- drop-ins for
ttts/vqvae/quantize.py:28-118 and ttts/vqvae/core_vq.py:96-382 whose hot arithmetic (L2-nearest-neighbour lookup, row gather,
straight-through, commitment loss, EMA codebook update) runs in hand-written sm_100a CUDA (csrc/vq.cu).

Same constructor signatures, same `forward / encode / decode` contracts, same buffer names
(`vq.layers.{i}._codebook.{inited,cluster_size,embed,embed_avg}`) so checkpoints interchange.

The rare, RNG-consuming maintenance steps of the reference -- k-means initialisation on the first training batch
(core_vq.py:60-93,141-150) and dead-code expiry (core_vq.py:152-168) -- are host-side torch logic on the device, exactly
as in the reference; they are not part of the per-step hot path.  No CPU fallback.
"""
import ctypes
import typing as tp

import torch
from torch import nn

from .. import _lib as L


def _protos(lib):
    if getattr(lib, "_vq_protos", False):
        return
    lib.ttts_vq_workspace_floats.restype = ctypes.c_int64
    lib.ttts_vq_workspace_floats.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.ttts_vq_forward.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                    ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.ttts_vq_ema_update.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int32, ctypes.c_int32, ctypes.c_float, ctypes.c_float, ctypes.c_void_p,
                                                             ctypes.c_void_p]
    lib.ttts_vq_backward.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib._vq_protos = True


def vq_lookup(x, embed, layout_bdn, want_quantized=True, straight_through=False, want_commit=False, hist=None, embed_sum=None):
    """Raw kernel call.  x: [B,D,N] (layout_bdn) or [N,D]; returns (codes int64 [N_total], quantized|None, commit|None)."""
    lib = L.lib(); _protos(lib)
    L.require_cuda(x, embed)
    assert x.dtype == torch.float32 and embed.dtype == torch.float32 and x.is_contiguous() and embed.is_contiguous()
    if layout_bdn:
        B, D, Nn = x.shape
        N = B * Nn
    else:
        B, D = x.shape
        Nn, N = 1, B
    K = embed.shape[0]
    assert embed.shape[1] == D
    codes = torch.empty(N, dtype=torch.int64, device=x.device)
    q = torch.empty_like(x) if want_quantized else None
    commit = torch.zeros((), dtype=torch.float32, device=x.device) if want_commit else None
    ws = torch.empty(lib.ttts_vq_workspace_floats(N, K), dtype=torch.float32, device=x.device)
    L.check(lib.ttts_vq_forward(x.data_ptr(), B, D, Nn, int(layout_bdn), embed.data_ptr(), K, codes.data_ptr(),
                                q.data_ptr() if q is not None else None, int(straight_through),
                                commit.data_ptr() if commit is not None else None,
                                hist.data_ptr() if hist is not None else None,
                                embed_sum.data_ptr() if embed_sum is not None else None, ws.data_ptr(), L.stream_ptr().value),
            "ttts_vq_forward")
    return codes, q, commit


class _VQFn(torch.autograd.Function):
    """quantize = x + (q - x).detach() ; commit = mse(q.detach(), x)   (core_vq.py:311-318) with the fused backward."""

    @staticmethod
    def forward(ctx, x, codebook, training):
        cb = codebook
        hist = embed_sum = None
        if training:
            hist = torch.zeros(cb.codebook_size, dtype=torch.float32, device=x.device)
            embed_sum = torch.zeros_like(cb.embed)
        codes, q, commit = vq_lookup(x, cb.embed, True, True, straight_through=training, want_commit=training, hist=hist, embed_sum=embed_sum)
        ctx.save_for_backward(x, codes, cb.embed.clone() if training else cb.embed)
        ctx.training = training
        ctx.mark_non_differentiable(codes)
        if training:
            # reference order (core_vq.py:209-228): lookup -> expire_codes_ (RNG-consuming; its row replacement is then
            # overwritten by the EMA renormalisation below, exactly as in the reference) -> EMA update
            B, D, Nn = x.shape
            cb.expire_codes_(x.detach().permute(0, 2, 1).reshape(B * Nn, D))
            cb._ema_update(hist, embed_sum)
        else:
            commit = torch.zeros((), dtype=torch.float32, device=x.device)
        return q, codes, commit

    @staticmethod
    def backward(ctx, dq, _dcodes, dcommit):
        x, codes, embed = ctx.saved_tensors
        if not ctx.training:
            return None, None, None      # eval output is a buffer gather: no gradient path (core_vq.py:310)
        lib = L.lib(); _protos(lib)
        B, D, Nn = x.shape
        dx = torch.empty_like(x)
        dq_c = dq.contiguous().float() if dq is not None else None
        dc = dcommit.contiguous().float() if dcommit is not None else None
        L.check(lib.ttts_vq_backward(x.data_ptr(), B, D, Nn, 1, embed.data_ptr(), codes.data_ptr(),
                                     dq_c.data_ptr() if dq_c is not None else None, dc.data_ptr() if dc is not None else None,
                                     dx.data_ptr(), L.stream_ptr().value), "ttts_vq_backward")
        return dx, None, None


def _sample_vectors(samples, num):
    n = samples.shape[0]
    if n >= num:
        idx = torch.randperm(n, device=samples.device)[:num]
    else:
        idx = torch.randint(0, n, (num,), device=samples.device)
    return samples[idx]


class EuclideanCodebook(nn.Module):
    """ttts/vqvae/core_vq.py:96-230 (buffers only; no gradient flows into the codebook)."""

    def __init__(self, dim, codebook_size, kmeans_init=False, kmeans_iters=10, decay=0.99, epsilon=1e-5, threshold_ema_dead_code=2):
        super().__init__()
        self.decay = decay
        if kmeans_init:
            embed = torch.zeros(codebook_size, dim)
        else:
            embed = torch.empty(codebook_size, dim)
            nn.init.kaiming_uniform_(embed)
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.epsilon = epsilon
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())
        self._scratch = None
        self._inited_host = False        # host copy of `inited` once it has been seen set: no device read per training forward

    def _load_from_state_dict(self, *args, **kwargs):
        self._inited_host = False        # a checkpoint may carry inited = 0 or 1: look again
        return super()._load_from_state_dict(*args, **kwargs)

    # ---- rare host-side maintenance (same algorithm / RNG consumption pattern as the reference) ----
    @torch.no_grad()
    def init_embed_(self, data):
        if self._inited_host:
            return
        if bool(self.inited.item()):
            self._inited_host = True
            return
        samples = data[:500, :]
        means = _sample_vectors(samples, self.codebook_size)
        for _ in range(self.kmeans_iters):
            codes, _, _ = vq_lookup(samples.contiguous(), means.contiguous(), False, want_quantized=False)
            bins = torch.bincount(codes, minlength=self.codebook_size)
            zero = bins == 0
            new_means = torch.zeros_like(means)
            new_means.index_add_(0, codes, samples)
            new_means = new_means / bins.clamp(min=1)[:, None]
            means = torch.where(zero[:, None], means, new_means)
        self.embed.data.copy_(means)
        self.embed_avg.data.copy_(means)
        self.cluster_size.data.copy_(bins.float())
        self.inited.data.fill_(1.0)
        self._inited_host = True

    @torch.no_grad()
    def expire_codes_(self, batch_samples):
        """core_vq.py:153-161.  The reference returns early (and draws no random numbers) when no code is dead, which costs a device -> host
        read per training forward; on the GPU the replacement is applied unconditionally instead -- `where(expired, sample, embed)` with an
        all-false mask is the identity -- so the training step has no host synchronisation and can be captured in a CUDA graph.  The values
        are the reference's whenever nothing expired; when something did, the replacement rows are random batch vectors in both."""
        if self.threshold_ema_dead_code == 0:
            return
        expired = self.cluster_size < self.threshold_ema_dead_code
        if not batch_samples.is_cuda and not bool(torch.any(expired)):
            return
        flat = batch_samples.reshape(-1, batch_samples.shape[-1])
        self.embed.data.copy_(torch.where(expired[:, None], _sample_vectors(flat, self.codebook_size), self.embed))

    # ---- data-parallel training ----
    # The reference keeps replicas consistent only through DDP's broadcast_buffers (rank 0's codebook overwrites the others before every
    # forward; the Encodec broadcast_tensors calls are commented out, core_vq.py:150,168), i.e. only rank 0's batch ever trains the codebook.
    # `sync="allreduce"` (default under an initialised process group) instead sums the per-rank code histogram and embedding sums
    # (1024 + 1024*192 floats = 0.79 MB) before the EMA update, so every rank applies the identical update computed from the global
    # batch (Tortoise's dvae.py:116-118 does the same).  `sync="rank0"` reproduces the reference: every rank takes rank 0's statistics.
    # `sync=None` leaves replicas to DDP.
    sync = "allreduce"

    @staticmethod
    def sync_stats(hist, embed_sum, mode, group=None):
        """In-place synchronisation of the EMA statistics across ranks (host logic; tested on gloo in tests/test_ddp_gloo_cpu.py)."""
        import torch.distributed as dist
        if mode is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return hist, embed_sum
        if mode == "allreduce":
            flat = torch.cat([hist.reshape(-1), embed_sum.reshape(-1)])            # one collective, not two
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            hist.copy_(flat[:hist.numel()].view_as(hist))
            embed_sum.copy_(flat[hist.numel():].view_as(embed_sum))
        elif mode == "rank0":
            flat = torch.cat([hist.reshape(-1), embed_sum.reshape(-1)])
            dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            hist.copy_(flat[:hist.numel()].view_as(hist))
            embed_sum.copy_(flat[hist.numel():].view_as(embed_sum))
        else:
            raise ValueError("EuclideanCodebook.sync must be 'allreduce', 'rank0' or None, got %r" % (mode,))
        return hist, embed_sum

    @torch.no_grad()
    def _ema_update(self, hist, embed_sum):
        self.sync_stats(hist, embed_sum, self.sync)
        lib = L.lib(); _protos(lib)
        if self._scratch is None or self._scratch.device != self.embed.device:
            self._scratch = torch.zeros(4, dtype=torch.float32, device=self.embed.device)
        L.check(lib.ttts_vq_ema_update(self.embed.data_ptr(), self.embed_avg.data_ptr(), self.cluster_size.data_ptr(), hist.data_ptr(),
                                       embed_sum.data_ptr(), self.codebook_size, self.embed.shape[1], float(self.decay), float(self.epsilon),
                                       self._scratch.data_ptr(), L.stream_ptr().value), "ttts_vq_ema_update")

    # ---- API parity helpers ----
    def quantize(self, x):
        codes, _, _ = vq_lookup(x.contiguous().float(), self.embed, False, want_quantized=False)
        return codes

    def dequantize(self, embed_ind):
        return torch.nn.functional.embedding(embed_ind, self.embed)

    def encode(self, x):
        shape = x.shape
        return self.quantize(x.reshape(-1, shape[-1])).view(*shape[:-1])

    def decode(self, embed_ind):
        return self.dequantize(embed_ind)


class VectorQuantization(nn.Module):
    """ttts/vqvae/core_vq.py:233-322 (codebook_dim == dim only: project_in/out are Identity in the reference's use)."""

    def __init__(self, dim, codebook_size, codebook_dim=None, decay=0.99, epsilon=1e-5, kmeans_init=True, kmeans_iters=50,
                 threshold_ema_dead_code=2, commitment_weight=1.0):
        super().__init__()
        if codebook_dim is not None and codebook_dim != dim:
            raise NotImplementedError("codebook_dim != dim (projection) is not used by the reference's quantizer")
        self.project_in = nn.Identity()
        self.project_out = nn.Identity()
        self.epsilon = epsilon
        self.commitment_weight = commitment_weight
        self._codebook = EuclideanCodebook(dim=dim, codebook_size=codebook_size, kmeans_init=kmeans_init, kmeans_iters=kmeans_iters,
                                           decay=decay, epsilon=epsilon, threshold_ema_dead_code=threshold_ema_dead_code)
        self.codebook_size = codebook_size

    @property
    def codebook(self):
        return self._codebook.embed

    def encode(self, x):                       # x [B, D, N] -> [B, N]
        L.require_cuda(x)
        B, D, N = x.shape
        codes, _, _ = vq_lookup(x.contiguous().float(), self._codebook.embed, True, want_quantized=False)
        return codes.view(B, N)

    def decode(self, embed_ind):               # [B, N] -> [B, D, N]
        return self._codebook.decode(embed_ind).permute(0, 2, 1)

    def forward(self, x):                      # x [B, D, N]
        L.require_cuda(x)
        cb = self._codebook
        x = x.contiguous().float()
        B, D, N = x.shape
        if self.training:
            with torch.no_grad():
                flat = x.detach().permute(0, 2, 1).reshape(B * N, D)
                cb.init_embed_(flat)                   # first training batch only (core_vq.py:209)
        quantize, codes, commit = _VQFn.apply(x, cb, self.training)
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        if self.training and self.commitment_weight > 0:
            loss = loss + commit * self.commitment_weight
        return quantize, codes.view(B, N), loss


class ResidualVectorQuantization(nn.Module):
    """ttts/vqvae/core_vq.py:325-382."""

    def __init__(self, *, num_quantizers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantization(**kwargs) for _ in range(num_quantizers)])

    def forward(self, x, n_q: tp.Optional[int] = None, layers: tp.Optional[list] = None):
        quantized_out = 0.0
        residual = x
        all_losses, all_indices, out_quantized = [], [], []
        n_q = n_q or len(self.layers)
        for i, layer in enumerate(self.layers[:n_q]):
            quantized, indices, loss = layer(residual)
            residual = residual - quantized
            quantized_out = quantized_out + quantized
            all_indices.append(indices)
            all_losses.append(loss)
            if layers and i in layers:
                out_quantized.append(quantized)
        out_losses, out_indices = map(torch.stack, (all_losses, all_indices))
        return quantized_out, out_indices, out_losses, out_quantized

    def encode(self, x, n_q: tp.Optional[int] = None, st: tp.Optional[int] = None):
        residual = x
        all_indices = []
        n_q = n_q or len(self.layers)
        st = st or 0
        for layer in self.layers[st:n_q]:
            indices = layer.encode(residual)
            all_indices.append(indices)
            if len(self.layers[st:n_q]) > 1:
                residual = residual - layer.decode(indices)
        return torch.stack(all_indices)

    def decode(self, q_indices, st: int = 0):
        quantized_out = torch.tensor(0.0, device=q_indices.device)
        for i, indices in enumerate(q_indices):
            quantized_out = quantized_out + self.layers[st + i].decode(indices)
        return quantized_out


class ResidualVectorQuantizer(nn.Module):
    """ttts/vqvae/quantize.py:28-118; constructed as (dimension=192, n_q=1, bins=1024) at ttts/vqvae/vq2.py:835."""

    def __init__(self, dimension=256, n_q=8, bins=1024, decay=0.99, kmeans_init=True, kmeans_iters=50, threshold_ema_dead_code=2):
        super().__init__()
        self.n_q = n_q
        self.dimension = dimension
        self.bins = bins
        self.decay = decay
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.vq = ResidualVectorQuantization(dim=dimension, codebook_size=bins, num_quantizers=n_q, decay=decay, kmeans_init=kmeans_init,
                                             kmeans_iters=kmeans_iters, threshold_ema_dead_code=threshold_ema_dead_code)

    def forward(self, x, n_q: tp.Optional[int] = None, layers: tp.Optional[list] = None):
        n_q = n_q if n_q else self.n_q
        if layers and max(layers) >= n_q:
            raise ValueError(f"Last layer index in layers: A {max(layers)}. Number of quantizers in RVQ: B {self.n_q}. A must less than B.")
        quantized, codes, commit_loss, quantized_list = self.vq(x, n_q=n_q, layers=layers)
        return quantized, codes, torch.mean(commit_loss), quantized_list

    def encode(self, x, n_q: tp.Optional[int] = None, st: tp.Optional[int] = None):
        n_q = n_q if n_q else self.n_q
        st = st or 0
        return self.vq.encode(x, n_q=n_q, st=st)

    def decode(self, codes, st: int = 0):
        return self.vq.decode(codes, st=st)
