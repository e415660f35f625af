"""The product backend of the diffusion training graph: `CudaKernels` of the VQ-VAE tape (convolutions, element-wise ops, the latents' cross
attention, masked mean) plus the ops this model adds, each ONE call into libttts_b200.so (csrc/diffusion_kernels.cu) on the current stream.
Device tensors only; no CPU fallback (off-GPU every method raises through `require_cuda`).

The wide stride-1 convolutions of AA_diffusion (512 / 1024 / 1536 channels, K = 1 or 3: GEMM-shaped, 56 % of the exact-fp32 step in the r2o
launch list) run on the tcgen05 GEMM (`ttts_gemm_bf16`) with SPLIT-bf16 operands: x = hi + lo, w = hi + lo, x w ~ hi hi + lo hi + hi lo, fp32
accumulation in TMEM -- fp32-grade results (~1e-5 relative, the recipe of conv1d_tcs) at tensor-core speed.  [B, C, T] activations are
converted to position-major [hi | lo] rows with zero rows between the clips (`ttts_cl_split`), so a tap of a K = 3 convolution is the same
buffer read one row earlier / later; forward, input gradient and weight gradient are then plain GEMMs:
    forward : D[m, co]   = sum_k [hi | lo](m + k - pad) . [wh_k | wh_k]^T  +  hi(m + k - pad) . wl_k^T            (2 launches per tap)
    dgrad   : dX[m, ci]  = sum_k [dyh | dyl](m - k + pad) . [wh_k ; wh_k]  +  dyh(m - k + pad) . wl_k
    wgrad   : dW_k       = dyh^T . [xh | xl](. + k - pad)  (two column blocks, summed)  +  dyl^T . xh(. + k - pad)      (split-K, fp32 red.add)
`TTTS_DIFF_TC=0` keeps every convolution on the exact-fp32 CUDA-core kernels."""
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
            lib.ttts_cl_split.argtypes = [vp, vp, i32, i32, i32, vp]
            lib.ttts_cl_unpack.argtypes = [vp, vp, i32, i32, i32, i32, vp]
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

    # ---------------------------------------------------------------- tensor-core convolutions ----------------------------------------------------------------
    TC_MIN_POSITIONS = 2048

    def _tc_ok(self, x, w, stride, dil, pad, pre_lrelu, groups):
        ct = self.__dict__.get("conv_tc")
        if ct is None:
            ct = self.conv_tc = os.environ.get("TTTS_DIFF_TC", "1") != "0"
        if not ct:
            return False
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        return (x.is_cuda and groups == 1 and stride == 1 and dil == 1 and K in (1, 3) and 2 * pad == K - 1 and not pre_lrelu
                and Cin % 64 == 0 and Cout % 64 == 0 and B * T >= self.TC_MIN_POSITIONS)

    def _buf(self, tag, shape, dtype, dev, zero=False):
        """transient buffers, one per (tag, shape): every use is ordered on the current stream"""
        pool = self.__dict__.setdefault("_tc_pool", {})
        key = (tag, tuple(shape), dtype, dev)
        if key not in pool:
            pool[key] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=dev)
        return pool[key]

    def _cl_split(self, tag, x):
        """x [B,C,T] -> [2 + B (T + 1), 2C] bf16 rows [hi | lo]; the rows the kernel never writes are the zero padding"""
        self._diff_protos(self.lib)
        B, C, T = x.shape
        buf = self._buf(tag, (2 + B * (T + 1), 2 * C), torch.bfloat16, x.device, zero=True)
        self._chk(self.lib.ttts_cl_split(self._p(x), self._p(buf), B, C, T, self._st()), "ttts_cl_split")
        return buf

    def _cl_unpack(self, D, B, C, T):
        y = torch.empty(B, C, T, dtype=torch.float32, device=D.device)
        self._chk(self.lib.ttts_cl_unpack(self._p(D), self._p(y), B, C, T, D.stride(0), self._st()), "ttts_cl_unpack")
        return y

    @staticmethod
    def _split_weights(w):
        wk = w.permute(2, 0, 1).contiguous()                          # [K, Cout, Cin]
        wh = wk.bfloat16()
        wl = (wk - wh.float()).bfloat16()
        return wh, wl

    def conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        if not self._tc_ok(x, w, stride, dil, pad, pre_lrelu, groups):
            return super().conv_fwd(x, w, b, stride, dil, pad, pre_lrelu, groups)
        L = self.L
        self._req(x, w, b)
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        M = B * (T + 1)
        X = self._cl_split("x", x)
        wh, wl = self._split_weights(w)
        D = self._buf("D", (M, Cout), torch.float32, x.device)
        bias = b.clone() if (b is not None and b.data_ptr() % 16) else b
        first = True
        for k in range(K):
            A = X[1 + k - pad:1 + k - pad + M]
            L.gemm(A, torch.cat([wh[k], wh[k]], dim=1), D, epi=L.EPI_F32 if first else L.EPI_F32_ADD, bias=bias if first else None)
            L.gemm(A[:, :Cin], wl[k], D, epi=L.EPI_F32_ADD)
            first = False
        return self._cl_unpack(D, B, Cout, T)

    def conv_bwd(self, dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups=1):
        if not self._tc_ok(x, w, stride, dil, pad, pre_lrelu, groups):
            return super().conv_bwd(dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups)
        L = self.L
        dy = dy.contiguous()
        self._req(dy, x, w)
        B, Cin, T = x.shape
        Cout, _, K = w.shape
        M = B * (T + 1)
        DY = self._cl_split("dy", dy)
        wh, wl = self._split_weights(w)
        dx = None
        if need_dx:
            D = self._buf("D", (M, Cin), torch.float32, x.device)
            first = True
            for k in range(K):
                A = DY[1 + pad - k:1 + pad - k + M]
                L.gemm(A, torch.cat([wh[k], wh[k]], dim=0), D, b_mn=True, epi=L.EPI_F32 if first else L.EPI_F32_ADD)
                L.gemm(A[:, :Cout], wl[k], D, b_mn=True, epi=L.EPI_F32_ADD)
                first = False
            dx = self._cl_unpack(D, B, Cin, T)
        X = self._cl_split("x", x)
        acc = torch.zeros(K, Cout, 2 * Cin, dtype=torch.float32, device=x.device)
        for k in range(K):
            Xk = X[1 + k - pad:1 + k - pad + M]
            L.gemm(DY[1:1 + M, :Cout], Xk, acc[k], a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
            L.gemm(DY[1:1 + M, Cout:], Xk[:, :Cin], acc[k][:, :Cin], a_mn=True, b_mn=True, epi=L.EPI_F32_ADD, split_k=16)
        dw = (acc[:, :, :Cin] + acc[:, :, Cin:]).permute(1, 2, 0).contiguous()
        db = None
        if need_db:
            db = torch.zeros(Cout, dtype=torch.float32, device=x.device)
            self._chk(self.lib.ttts_bias_grad(self._p(dy), self._p(db), B, Cout, T, self._st()), "ttts_bias_grad")
        return dx, dw, db


class DiffusionCudaKernels(DiffusionKernelsMixin, CudaKernels):
    pass
