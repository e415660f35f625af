"""CPU: the host side of the split-bf16 GEMM route of the wide training convolutions (ttts_b200/vqvae/train_encoder.py: `_gemm_conv_fwd`,
`_gemm_conv_bwd`) -- row geometry of the position-major buffers, the OVERLAPPED tap-concatenated operand view, flipped / transposed weights
of the input gradient, strided scatter of the per-tap form, the weight-gradient slices -- with the layout kernels (ttts_cl_split / ttts_cl_unpack,
ttts_lrelu, ttts_bias_grad) running from their CUDA source on the CPU emulation (tests/emu) and the tcgen05 GEMM replaced by a torch matmul on
the SAME bf16 operand views and epilogue flags.  What it cannot show is the TMA side of the overlapped view (row pitch < row length in the
tensor map); that is tests/test_gpu_diffusion.py::test_tensor_core_convolution_vs_torch on a B200."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu_kernels as EK  # noqa: E402
from ttts_b200 import _lib as RL  # noqa: E402


class TorchGemm:
    """stands in for ttts_b200._lib on the GEMM calls of the route: same operand conventions (A [M,K] or [K,M] with a_mn; B [N,K] or [K,N] with
    b_mn; bf16 operands, fp32 accumulation), same epilogue meaning (EPI_F32 = overwrite (+ bias), EPI_F32_ADD = accumulate)"""
    EPI_F32, EPI_F32_ADD = RL.EPI_F32, RL.EPI_F32_ADD

    def __init__(self):
        self.calls = 0

    def gemm(self, A, B, out, *, a_mn=False, b_mn=False, epi=None, bias=None, split_k=1, **kw):
        assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16 and out.dtype == torch.float32 and A.stride(-1) == 1 and B.stride(-1) == 1
        self.calls += 1
        Af = A.float().t() if a_mn else A.float()
        Bf = B.float() if b_mn else B.float().t()
        R = Af @ Bf
        if bias is not None:
            R = R + bias
        if epi == self.EPI_F32:
            out.copy_(R)
        else:
            assert epi == self.EPI_F32_ADD
            out.add_(R)
        return out


@pytest.fixture(scope="module")
def K(tmp_path_factory):
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    k = EK.EmuKernels(EK.build_all(str(tmp_path_factory.mktemp("emu"))))
    return k


@pytest.mark.parametrize("tapcat", [True, False])
@pytest.mark.parametrize("B,Cin,T,Cout,Kw,stride,dil,pad,lrelu", [
    (2, 64, 50, 128, 3, 1, 1, 1, False), (2, 128, 40, 64, 5, 3, 1, 2, False), (1, 64, 45, 64, 7, 1, 1, 3, True), (2, 64, 33, 64, 1, 1, 1, 0, False),
    (1, 64, 60, 64, 7, 2, 3, 9, True), (2, 64, 37, 128, 5, 1, 1, 0, False), (1, 64, 41, 64, 4, 1, 1, 3, False),
    (2, 32, 50, 32, 7, 1, 1, 3, True), (1, 32, 70, 128, 5, 3, 1, 2, False), (2, 64, 31, 32, 3, 1, 1, 1, False),
    # dilated stride-1 layers whose padding is a multiple of the dilation: de-interleaved sub-clips, tap-concatenated like dilation 1
    (1, 64, 60, 64, 7, 1, 3, 9, True), (2, 64, 50, 128, 3, 1, 5, 5, False), (1, 32, 48, 32, 5, 1, 3, 6, True), (1, 64, 45, 64, 3, 1, 3, 0, False),
    (2, 64, 40, 64, 3, 1, 5, 10, True), (2, 64, 41, 64, 7, 1, 3, 9, True), (1, 32, 53, 64, 3, 1, 5, 5, True), (1, 64, 44, 64, 3, 1, 3, 0, False)])
def test_gemm_route_geometry(K, monkeypatch, tapcat, B, Cin, T, Cout, Kw, stride, dil, pad, lrelu):
    F = torch.nn.functional
    g = torch.Generator().manual_seed(Cin + T + Kw)
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, Kw, generator=g) / (Cin * Kw) ** 0.5
    b = torch.randn(Cout, generator=g)
    if not tapcat and (Cin % 64 or Cout % 64):
        pytest.skip("narrow layers take the route in the tap-concatenated form only")
    fake = TorchGemm()
    monkeypatch.setattr(K, "L", fake)
    monkeypatch.setattr(K, "tap_concat", tapcat)
    monkeypatch.setattr(K, "wgrad_concat", tapcat)
    xr, wr, br = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.conv1d(F.leaky_relu(xr, 0.1) if lrelu else xr, wr, br, stride=stride, dilation=dil, padding=pad)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy.double())
    y = K._gemm_conv_fwd(x, w, b, stride, dil, pad, lrelu)[0]
    n_fwd = fake.calls
    dx, dw, db = K._gemm_conv_bwd(dy, x, w, stride, dil, pad, lrelu, True, True)
    rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm())
    assert y.shape == yr.shape and dx.shape == x.shape and dw.shape == w.shape
    # three bf16 products per fp32 product: ~1e-5; the tolerance leaves room for the dropped lo . lo term
    assert rel(y, yr) <= 3e-5, rel(y, yr)
    assert rel(dx, xr.grad) <= 3e-5, rel(dx, xr.grad)
    assert rel(dw, wr.grad) <= 3e-5, rel(dw, wr.grad)
    assert rel(db, br.grad) <= 1e-5
    if tapcat and Kw > 1 and (dil == 1 or (stride == 1 and pad % dil == 0)):
        assert n_fwd == 2                                            # one GEMM pair for all taps
    else:
        assert n_fwd == 2 * Kw
