"""Unit tests of individual CUDA kernels against plain PyTorch fp32 references (floating-point kernels
keep a torch reference; the path-level oracle is in test_gpu_parity.py).  pytest -m gpu."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(prec, A, B, bias, relu):
    from graph_pde_b200 import _lib
    L = _lib.lib()
    _lib.check(L.nnconv_init())
    M, K = A.shape
    N = B.shape[0]
    C = torch.empty(M, N, dtype=A.dtype, device=A.device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.nnconv_gemm_16b(_lib.PREC[prec], ctypes.c_void_p(A.data_ptr()), M, K, ctypes.c_void_p(B.data_ptr()),
                                 N, ctypes.c_void_p(bias.data_ptr() if bias is not None else 0), int(relu),
                                 ctypes.c_void_p(C.data_ptr()), st))
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize('prec,dtype', [('f16', torch.float16), ('bf16', torch.bfloat16)])
@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (1000, 256, 128), (77, 128, 64), (4096, 1024, 1024),
                                   (300, 2048, 64), (129, 192, 320), (40000, 64, 256)])
def test_tcgen05_gemm_vs_torch(prec, dtype, M, N, K):
    torch.manual_seed(M + N + K)
    dev = torch.device('cuda:0')
    A = (torch.randn(M, K, device=dev) * 0.5).to(dtype)
    B = (torch.randn(N, K, device=dev) * 0.5).to(dtype)
    bias = torch.randn(N, device=dev)
    ref = A.float() @ B.float().t()
    out = _gemm(prec, A, B, None, False).float()
    tol = 2e-3 if prec == 'f16' else 1.6e-2          # output rounding to the 16-bit type dominates
    scale = ref.abs().max()
    assert float((out - ref).abs().max() / scale) < tol
    out2 = _gemm(prec, A, B, bias, True).float()
    ref2 = torch.relu(ref + bias)
    assert float((out2 - ref2).abs().max() / ref2.abs().max()) < tol
