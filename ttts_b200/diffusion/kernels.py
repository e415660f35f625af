"""The product backend of the diffusion training graph: `CudaKernels` of the VQ-VAE tape (convolutions, element-wise ops, the latents' cross
attention, masked mean) plus the ops this model adds, each ONE call into libttts_b200.so (csrc/diffusion_kernels.cu) on the current stream.
Device tensors only; no CPU fallback (off-GPU every method raises through `require_cuda`).

The wide convolutions of AA_diffusion (512 / 1024 / 1536 channels: GEMM-shaped, 56 % of the exact-fp32 step in the r2o launch list) run on the
tcgen05 GEMM with split-bf16 operands -- the route `CudaKernels.conv_fwd / conv_bwd` take for every wide layer of the training tapes
(ttts_b200/vqvae/train_encoder.py: `_gemm_conv_*`); `TTTS_DIFF_TC=0` or `TTTS_TRAIN_GEMM=0` keeps every convolution on the exact-fp32
CUDA-core kernels."""
import ctypes
import os

import torch

from ..vqvae.train_encoder import CudaKernels


class DiffusionKernelsMixin:
    @staticmethod
    def _diff_protos(lib):
        if not getattr(lib, "_diff_protos", False):
            vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
            lib.ttts_groupnorm.argtypes = [vp] * 7 + [i32] * 5 + [vp]
            lib.ttts_groupnorm_bwd.argtypes = [vp] * 13 + [i32] * 5 + [vp]
            lib.ttts_silu.argtypes = [vp, vp, vp, i64, i32, vp]
            lib.ttts_attn_bias.argtypes = [vp] * 5 + [i32] * 4 + [vp]
            lib.ttts_attn_bias_bwd_scratch_floats.argtypes = [i32, i32, i32]
            lib.ttts_attn_bias_bwd_scratch_floats.restype = i64
            lib.ttts_attn_bias_bwd.argtypes = [vp] * 9 + [i32] * 4 + [vp]
            lib.ttts_diff_q_sample.argtypes = [vp, vp, vp, vp, i32, i64, vp]
            lib.ttts_diff_loss.argtypes = [vp] * 9 + [i32] * 3 + [vp]
            lib.ttts_diff_loss_bwd.argtypes = [vp] * 8 + [i32] * 3 + [vp]
            lib._diff_protos = True

    def _reqi(self, t):
        self._device_check([t])
        assert t.is_contiguous() and t.dtype == torch.int32

    def gn_fwd(self, x, gamma, beta, groups, scale, shift, silu):
        self._diff_protos(self.lib)
        self._req(x, gamma, beta, scale, shift)
        B, C, T = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(B, groups, 2, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_groupnorm(self._p(x), self._p(gamma), self._p(beta), self._p(scale), self._p(shift), self._p(y), self._p(stats),
                                          B, C, T, groups, int(bool(silu)), self._st()), "ttts_groupnorm")
        return y, stats

    def gn_bwd(self, dy, x, stats, gamma, beta, groups, scale, shift, silu):
        dy = dy.contiguous()
        self._req(dy, x, stats, gamma, beta, scale, shift)
        B, C, T = x.shape
        dx, dg, db = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(beta)
        dsc = torch.empty_like(scale) if scale is not None else None
        dsh = torch.empty_like(shift) if shift is not None else None
        scratch = torch.empty(B * C * 2, dtype=torch.float32, device=x.device)
        self._chk(self.lib.ttts_groupnorm_bwd(self._p(dy), self._p(x), self._p(stats), self._p(gamma), self._p(beta), self._p(scale), self._p(shift),
                                              self._p(dx), self._p(dg), self._p(db), self._p(dsc), self._p(dsh), self._p(scratch), B, C, T, groups,
                                              int(bool(silu)), self._st()), "ttts_groupnorm_bwd")
        return dx, dg, db, dsc, dsh

    def silu_fwd(self, x):
        self._diff_protos(self.lib)
        self._req(x)
        o = torch.empty_like(x)
        self._chk(self.lib.ttts_silu(self._p(x), None, self._p(o), x.numel(), 0, self._st()), "ttts_silu")
        return o

    def silu_bwd(self, dy, x):
        dy = dy.contiguous()
        self._req(dy, x)
        o = torch.empty_like(x)
        self._chk(self.lib.ttts_silu(self._p(x), self._p(dy), self._p(o), x.numel(), 1, self._st()), "ttts_silu")
        return o

    def attn_bias_fwd(self, qkv, table, heads, diag):
        self._diff_protos(self.lib)
        self._req(qkv, table); self._reqi(diag)
        B, W, T = qkv.shape
        C = W // 3
        assert diag.numel() == 2 * T - 1 and tuple(table.shape) == (32, heads)
        out = torch.empty(B, C, T, dtype=torch.float32, device=qkv.device)
        lse = torch.empty(B, heads, T, dtype=torch.float32, device=qkv.device)
        self._chk(self.lib.ttts_attn_bias(self._p(qkv), self._p(table), self._p(diag), self._p(out), self._p(lse), B, C, T, heads, self._st()), "ttts_attn_bias")
        return out, lse

    def attn_bias_bwd(self, do, qkv, out, lse, table, heads, diag):
        do = do.contiguous()
        self._req(do, qkv, out, lse, table); self._reqi(diag)
        B, W, T = qkv.shape
        C = W // 3
        dqkv, dtab = torch.empty_like(qkv), torch.empty_like(table)
        scratch = torch.empty(int(self.lib.ttts_attn_bias_bwd_scratch_floats(B, T, heads)), dtype=torch.float32, device=qkv.device)
        self._chk(self.lib.ttts_attn_bias_bwd(self._p(do), self._p(qkv), self._p(out), self._p(lse), self._p(table), self._p(diag), self._p(dqkv),
                                              self._p(dtab), self._p(scratch), B, C, T, heads, self._st()), "ttts_attn_bias_bwd")
        return dqkv, dtab

    def q_sample(self, x_start, noise, coef):
        self._diff_protos(self.lib)
        self._req(x_start, noise, coef)
        xt = torch.empty_like(x_start)
        B = x_start.shape[0]
        self._chk(self.lib.ttts_diff_q_sample(self._p(x_start), self._p(noise), self._p(coef), self._p(xt), B, x_start.numel() // B, self._st()),
                  "ttts_diff_q_sample")
        return xt

    def diff_loss_fwd(self, out, x_start, x_t, noise, coef, t_is0):
        self._req(out, x_start, x_t, noise, coef); self._reqi(t_is0)
        B, Cn, T = x_start.shape
        assert tuple(out.shape) == (B, 2 * Cn, T)
        terms = torch.empty(2, B, dtype=torch.float32, device=out.device)
        loss = torch.empty(1, dtype=torch.float32, device=out.device)
        scratch = torch.empty(B * 32 * 2, dtype=torch.float32, device=out.device)
        self._chk(self.lib.ttts_diff_loss(self._p(out), self._p(x_start), self._p(x_t), self._p(noise), self._p(coef), self._p(t_is0), self._p(terms),
                                          self._p(loss), self._p(scratch), B, Cn, T, self._st()), "ttts_diff_loss")
        return loss, (terms[0], terms[1])

    def diff_loss_bwd(self, dL, out, x_start, x_t, noise, coef, t_is0):
        dL = dL.contiguous()
        self._req(dL, out, x_start, x_t, noise, coef); self._reqi(t_is0)
        B, Cn, T = x_start.shape
        d = torch.empty_like(out)
        self._chk(self.lib.ttts_diff_loss_bwd(self._p(dL), self._p(out), self._p(x_start), self._p(x_t), self._p(noise), self._p(coef), self._p(t_is0),
                                              self._p(d), B, Cn, T, self._st()), "ttts_diff_loss_bwd")
        return d


class DiffusionCudaKernels(DiffusionKernelsMixin, CudaKernels):
    pass
