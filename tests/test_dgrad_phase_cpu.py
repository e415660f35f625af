"""The phase decomposition of a strided convolution's input gradient (ttts_b200/vqvae/train_encoder.py::dgrad_by_phase) with torch's conv1d in
place of the CUDA forward kernel: index arithmetic against torch.autograd for the stride / kernel / padding combinations of the
down-sampling stack, the period discriminators and `proj`, and against conv_transpose1d for the Generator's up-sampling layers."""
import pytest
import torch
import torch.nn.functional as F

from ttts_b200.vqvae.train_encoder import dgrad_by_phase

FN = lambda a, wt, pd: F.conv1d(a, wt, None, padding=pd)


@pytest.mark.parametrize("B,Cin,T,Cout,K,s,pad", [(2, 5, 160, 7, 16, 8, 7), (2, 4, 61, 6, 7, 2, 3), (1, 3, 100, 4, 16, 10, 3), (2, 6, 40, 3, 2, 2, 0),
                                                 (2, 3, 33, 4, 5, 3, 2), (1, 2, 17, 3, 3, 5, 1), (2, 4, 50, 5, 8, 2, 3), (1, 2, 9, 2, 4, 4, 0),
                                                 (1, 3, 31, 2, 16, 10, 7), (1, 2, 5, 2, 4, 2, 5)])
def test_strided_input_gradient(B, Cin, T, Cout, K, s, pad):
    g = torch.Generator().manual_seed(K * 31 + s)
    x = torch.randn(B, Cin, T, generator=g, requires_grad=True)
    w = torch.randn(Cout, Cin, K, generator=g)
    y = F.conv1d(x, w, None, stride=s, padding=pad)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    got = dgrad_by_phase(FN, dy, w, T, s, pad)
    assert float((got - x.grad).abs().max()) <= 1e-5 * max(1.0, float(x.grad.abs().max()))


@pytest.mark.parametrize("Cin,Cout,K,s,pad,T", [(12, 6, 16, 10, 3, 7), (8, 4, 16, 8, 4, 9), (6, 3, 8, 2, 3, 20), (5, 40, 2, 2, 0, 33)])
def test_transposed_convolution_forward(Cin, Cout, K, s, pad, T):
    g = torch.Generator().manual_seed(Cin + K)
    x, w = torch.randn(2, Cin, T, generator=g), torch.randn(Cin, Cout, K, generator=g)
    y = F.conv_transpose1d(x, w, None, stride=s, padding=pad)
    got = dgrad_by_phase(FN, x, w, (T - 1) * s - 2 * pad + K, s, pad)
    assert float((got - y).abs().max()) <= 1e-5 * max(1.0, float(y.abs().max()))
