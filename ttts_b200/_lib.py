"""ctypes binding of libttts_b200.so (the C ABI declared in include/ttts_b200.h).

The product path has no CPU fallback: if the shared library is missing or the device is not
sm_100, every op raises.  (The CPU oracle lives under oracle/ and is test-only.)
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libttts_b200.so")

_lib = None


class TTTSError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int32), ("a_mn", ctypes.c_int32),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int32), ("b_mn", ctypes.c_int32),
        ("epi", ctypes.c_int32),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("aux", ctypes.c_void_p), ("ldaux", ctypes.c_int32),
        ("aux_out", ctypes.c_void_p), ("ldaux_out", ctypes.c_int32),
        ("split_k", ctypes.c_int32),
        ("drop_thresh16", ctypes.c_uint32), ("drop_scale", ctypes.c_float), ("drop_seed", ctypes.c_uint64),
    ]


EPI_BF16, EPI_GELU, EPI_RESID, EPI_DGELU, EPI_F32_ADD, EPI_F32 = range(6)


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TTTSError(
                "libttts_b200.so not found at %s -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.ttts_last_error.restype = ctypes.c_char_p
        L.ttts_version.restype = ctypes.c_int
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().ttts_last_error().decode("utf-8", "replace")
        raise TTTSError("%s failed (%d): %s" % (what or "ttts call", rc, msg))


# torch.cuda.current_stream() builds a Stream object and resolves the device index through several Python layers (~10 us); the training tapes
# ask for the stream once per kernel call, ~12 000 times per VQ-VAE-GAN step, and that step is bound by the host (r2al: enqueue time = step
# time, current_stream() 20 % of the profile).  The raw-stream accessor torch's own launchers use returns the same pointer in < 1 us.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise TTTSError("ttts_b200 ops run on sm_100a only; got a %s tensor (no CPU fallback)" % t.device)


def gemm(A, B, out, *, a_mn=False, b_mn=False, epi=EPI_BF16, bias=None, aux=None, aux_out=None, split_k=1,
         M=None, N=None, K=None, drop_p=0.0, drop_seed=0):
    """D[M,N] = A*B with fused epilogue.  A: [M,K] (a_mn=False) or [K,M]; B: [N,K] (b_mn=False) or [K,N]."""
    require_cuda(A, B, out)
    if M is None:
        M = A.shape[1] if a_mn else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn else B.shape[0]
    g = GemmArgs()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn = A.data_ptr(), A.stride(0), int(a_mn)
    g.B, g.ldb, g.b_mn = B.data_ptr(), B.stride(0), int(b_mn)
    g.epi = epi
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    g.bias = bias.data_ptr() if bias is not None else None
    g.aux, g.ldaux = (aux.data_ptr(), aux.stride(0)) if aux is not None else (None, 0)
    g.aux_out, g.ldaux_out = (aux_out.data_ptr(), aux_out.stride(0)) if aux_out is not None else (None, 0)
    g.split_k = split_k
    if drop_p > 0:
        g.drop_thresh16 = 2 * int(drop_p * 32768.0 + 0.5)
        g.drop_scale = 1.0 / (1.0 - g.drop_thresh16 / 65536.0)
        g.drop_seed = drop_seed
    check(lib().ttts_gemm_bf16(ctypes.byref(g), stream_ptr()), "ttts_gemm_bf16")
    return out
