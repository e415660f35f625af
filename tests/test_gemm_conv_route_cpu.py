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
    libs = EK.build_all(str(tmp_path_factory.mktemp("emu")))
    k = EK.EmuKernels(libs)
    k._libs_for_route = libs
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


class RouteKernels(EK.EmuKernels):
    """the emulation backend with EVERY eligible convolution of a training graph sent down the product's GEMM route (host code of
    ttts_b200/vqvae/train_encoder.py, layout / weight / bias-gradient kernels from their CUDA source on the emulation, the tcgen05 GEMM as a
    torch matmul on the same bf16 operands): the route's shape rule applies from one output position, the device check is the only thing skipped"""

    def __init__(self, libs):
        super().__init__(libs)
        self.L = TorchGemm()
        self.GEMM_MIN_POSITIONS = 1
        self.routed = {"fwd": 0, "bwd": 0, "dilated": 0, "narrow": 0}

    def conv_fwd(self, x, w, b, stride, dil, pad, pre_lrelu, groups=1):
        if groups == 1 and self._gemm_shape_ok(x, w, stride, dil, pad):
            self.routed["fwd"] += 1
            self.routed["dilated"] += dil > 1
            self.routed["narrow"] += min(w.shape[0], w.shape[1]) < 128
            return self._gemm_conv_fwd(x.contiguous(), w.contiguous(), b, stride, dil, pad, pre_lrelu)[0]
        return super().conv_fwd(x, w, b, stride, dil, pad, pre_lrelu, groups)

    def conv_bwd(self, dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups=1):
        if groups == 1 and self._gemm_shape_ok(x, w, stride, dil, pad):
            self.routed["bwd"] += 1
            return self._gemm_conv_bwd(dy, x.contiguous(), w.contiguous(), stride, dil, pad, pre_lrelu, need_dx, need_db)
        return super().conv_bwd(dy, x, w, stride, dil, pad, pre_lrelu, need_dx, need_db, groups)


@pytest.mark.skipif(os.environ.get("TTTS_SLOW_TESTS") != "1", reason="minutes on the emulation (one OS thread per CUDA thread): TTTS_SLOW_TESTS=1")
def test_decoder_training_graph_through_the_route(K, golden_dir):
    """The Generator's training graph (ResBlock1 stacks with dilations 1 / 3 / 5 at 256 ... 32 channels, conv_pre / conv_post; vq2.py Generator,
    modules.py:224-318) with every eligible convolution on the route -- dilated layers de-interleaved, narrow layers, the leaky-ReLU derivative
    folded into the conversion, one split of dy for both gradients -- against the REAL reference's waveform and parameter gradients
    (tests/golden/decoder.npz).  Tolerances: 2e-4 on the waveform (2e-5 in the fp32 graph test, tests/test_train_decoder_cpu.py: three bf16
    products per fp32 product through ~40 layers in sequence), gradients as in the GPU tests.  ~12 minutes on the emulation."""
    import numpy as np
    from oracle import decoder_oracle as DO
    from ttts_b200.vqvae.train_decoder import DecoderGraph
    dec = np.load(os.path.join(golden_dir, "decoder.npz"))
    P = DO.init_params(seed=9)
    RK = RouteKernels(K._libs_for_route)
    graph = DecoderGraph(RK, P)
    y = graph.forward(torch.tensor(dec["z"]), torch.tensor(dec["g"]))
    assert np.linalg.norm(y.v.numpy() - dec["y"]) <= 2e-4 * np.linalg.norm(dec["y"])
    R = torch.randn(y.v.shape, generator=torch.Generator().manual_seed(32))
    grads = graph.backward(R)
    assert RK.routed["fwd"] >= 30 and RK.routed["bwd"] >= 30 and RK.routed["dilated"] >= 8 and RK.routed["narrow"] >= 8, RK.routed
    _check_routed_grads(grads, [str(n) for n in dec["names"]], dec["norm"], dec["proj"])


def _check_routed_grads(grads, names, norms, projs):
    """The kink-aware comparison of the GPU tests (tests/test_gpu_encoder.py::check_param_grads) with a wider kink allowance: the route's
    arithmetic is fp32-GRADE (1e-5 relative per layer), not fp32 (1e-7 between summation orders), so at the golden shapes (a few hundred positions
    per level) a leaky-ReLU input lands within rounding of zero ~100x more often; one such flip moves the gradient rows of one channel through
    the convolutions of its ResBlock.  Every tensor must be within 6 %, the tensors of at most 8 convolutions may be looser than 2e-3 / 1e-2,
    and the root-mean-square norm error over ALL tensors must stay below 1e-3.  (First run: decoder 4 convolutions loose, worst norm error 7e-4.)"""
    import numpy as np
    from test_gpu_encoder import check_param_grads
    check_param_grads(grads, names, norms, projs, kink_layers=8)
    num = sum((float(grads[k].norm()) - float(n)) ** 2 for k, n in zip(names, norms))
    den = sum(float(n) ** 2 for n in norms)
    assert (num / den) ** 0.5 < 1e-3, (num / den) ** 0.5


@pytest.mark.skipif(os.environ.get("TTTS_SLOW_TESTS") != "1", reason="minutes on the emulation (one OS thread per CUDA thread): TTTS_SLOW_TESTS=1")
def test_encoder_training_graph_through_the_route(K, golden_dir):
    """The encode half's training graph (MelStyleEncoder, PosteriorAudioEncoder with its dilated 96 / 128 / 192-channel ResBlock1 stacks and the
    WN layers, proj; vq2.py:667-745, modules.py:136-318) with every eligible convolution on the route, against the REAL reference's outputs and
    its 414 parameter gradients (tests/golden/encoder.npz, encoder_grads.npz) -- the path the quantizer's codes and kl_ssl hang on."""
    import numpy as np
    from oracle import encoder_oracle as EO
    from oracle import vq_mel_oracle as V
    from ttts_b200.vqvae.train_encoder import EncoderGraph
    enc = np.load(os.path.join(golden_dir, "encoder.npz"))
    g = np.load(os.path.join(golden_dir, "encoder_grads.npz"))
    P = EO.init_params(seed=5)
    RK = RouteKernels(K._libs_for_route)
    graph = EncoderGraph(RK, P)
    z, x = graph.forward(torch.tensor(V.spectrogram(enc["wav"])), torch.tensor(enc["wav"]), lengths=torch.tensor(enc["lengths"]), eps=torch.tensor(enc["eps"]))
    assert np.abs(z.v.numpy() - enc["z"]).max() <= 5e-4 * np.abs(enc["z"]).max()
    assert np.abs(x.v.numpy() - enc["x"]).max() <= 5e-4 * np.abs(enc["x"]).max()
    R = torch.randn(3, 192, 36, generator=torch.Generator().manual_seed(123))
    grads = graph.backward(dz=R, dx=x.v / x.v.numel())
    assert RK.routed["fwd"] >= 30 and RK.routed["bwd"] >= 30 and RK.routed["dilated"] >= 8, RK.routed
    _check_routed_grads(grads, [str(n) for n in g["names"]], g["norm"], g["proj"])
