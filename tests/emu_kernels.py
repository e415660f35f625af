"""TEST INFRASTRUCTURE: `EmuKernels` = the PRODUCT backend class `CudaKernels` (ttts_b200/vqvae/train_encoder.py) with its library handle
replaced by the host builds of the same CUDA sources (tests/emu) and host tensors allowed -- so that the argument marshalling the GPU will see
(pointer order, shapes, flags) runs end to end on the CPU emulation.  The few forward ops whose kernels live in GPU-only translation units
(ttts_conv1d_f32, ttts_weight_norm, ttts_snake_aa, ttts_mha_small, ttts_masked_mean, ttts_posterior_sample, the VQ lookup, the log-mel
forward -- all validated on a B200 already) are served by the op contract tests/ref_kernels.py."""
import ctypes
import os
import subprocess

import torch

from emu_build import compile_emu

from ref_kernels import TorchRefKernels
from ttts_b200.vqvae.train_encoder import CudaKernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ("conv_bwd_emu.cpp", "encoder_bwd_emu.cpp", "gan_losses_emu.cpp", "stft_bwd_emu.cpp", "text_encoder_emu.cpp", "gconv_emu.cpp", "diffusion_emu.cpp")


def build_all(outdir):
    libs = []
    for src in SOURCES:
        so = os.path.join(outdir, "lib" + src[:-4] + ".so")
        compile_emu(src, so)
        libs.append(ctypes.CDLL(so))
    return libs


class EmuLib:
    """one namespace over the emulation libraries (each holds the extern "C" entry points of one .cu file)"""

    def __init__(self, libs):
        self._libs, self._cache = libs, {}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._cache:
            for l in self._libs:
                try:
                    self._cache[name] = getattr(l, name)
                    break
                except AttributeError:
                    continue
            else:
                raise AttributeError("no emulation library exports " + name)
        return self._cache[name]


class EmuKernels(CudaKernels):
    def __init__(self, libs):
        self.ref = TorchRefKernels()
        super().__init__(lib=EmuLib(libs))

    def _st(self):
        return None

    def _device_check(self, ts):
        assert all(not t.is_cuda for t in ts)

    def _chk(self, rc, what):
        assert rc == 0, what

    # ---- forward ops of GPU-only translation units: the op contract ----
    def conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        if groups > 1:
            return super().conv_fwd(x, w, b, stride, dil, pad, pre_lrelu, groups)
        return self.ref.conv_fwd(x, w, b, stride, dil, pad, pre_lrelu).contiguous()

    def wn_fwd(self, v, g):
        return self.ref.wn_fwd(v, g.reshape(-1)).contiguous()

    def snake_fwd(self, x, la, lb, filt):
        return self.ref.snake_fwd(x, la, lb, filt).contiguous()

    def mha_fwd(self, q, k, v, lens, heads, temperature):
        return self.ref.mha_fwd(q, k, v, lens, heads, temperature).contiguous()

    def masked_mean_fwd(self, x, lens):
        return self.ref.masked_mean_fwd(x, lens).contiguous()

    def posterior_fwd(self, stats, eps, mask):
        return self.ref.posterior_fwd(stats, eps, mask).contiguous()

    def vq_fwd(self, x, embed):
        q, c, codes = self.ref.vq_fwd(x, embed)
        return q.contiguous(), c, codes

    def vq_bwd(self, dq, dcommit, x, embed, codes):
        return self.ref.vq_bwd(dq, dcommit, x, embed, codes)

    def logmel_fwd(self, wav):
        return self.ref.logmel_fwd(wav).contiguous()


def emu_diffusion_kernels(libs):
    """the diffusion backend (ttts_b200/diffusion/kernels.py) over the same host builds"""
    from ttts_b200.diffusion.kernels import DiffusionKernelsMixin

    class EmuDiffusionKernels(DiffusionKernelsMixin, EmuKernels):
        pass
    return EmuDiffusionKernels(libs)
