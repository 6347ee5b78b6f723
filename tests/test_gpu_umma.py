"""tcgen05 building block (-m gpu): the K-major no-swizzle operand layout, descriptors, TMEM accumulate and
the 3-pass split-TF32 scheme reproduce an fp32 GEMM."""
import pytest
import torch

from tests import helpers as H  # noqa: F401
from dynrf_b200 import lib as L, ops

gpu = pytest.mark.gpu


@gpu
@pytest.mark.parametrize("K,N", [(8, 64), (32, 64), (96, 64), (152, 64), (108, 128), (72, 32), (131, 128)])
def test_umma_gemm(K, N):
    lib = H.selftest_lib()
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g).cuda()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    ref = A.double() @ B.double().T
    scale = float(ref.abs().max())
    for passes, tol in ((3, 3e-6), (1, 3e-3)):
        C = torch.full((128, N), float("nan"), device="cuda")
        L.check(lib.dynrf_selftest_umma_gemm(ops._p(A), ops._p(B), K, N, passes, ops._p(C), None), "selftest")
        torch.cuda.synchronize()
        err = float((C.double() - ref).abs().max()) / scale
        print(K, N, passes, err)
        assert err <= tol, (K, N, passes, err)
    # fp32 reference error for context
    print("fp32 torch err", float(((A @ B.T).double() - ref).abs().max()) / scale)


@gpu
@pytest.mark.parametrize("K,N,P", [(152, 64, 5000), (128, 128, 4096), (108, 128, 777), (216, 27, 3001), (64, 3, 1000), (96, 64, 31)])
def test_wgrad_gemm(K, N, P):
    """dW = XT @ DYT^T over samples on tensor cores (split TF32) == fp64 reference to fp32-level accuracy."""
    lib = L.load()
    g = torch.Generator().manual_seed(K + N + P)
    ld = (P + 3) // 4 * 4 + 8
    XT = torch.randn(K, ld, generator=g).cuda()
    DYT = torch.randn(N, ld, generator=g).cuda()
    dw = torch.zeros(N, K, device="cuda")
    cnt = torch.tensor([P], dtype=torch.int32, device="cuda")
    for _ in range(2):  # accumulates
        L.check(lib.dynrf_wgrad_gemm(ops._p(XT), ld, K, ops._p(DYT), ld, N, ld, ops._p(cnt), ops._p(dw), K, None), "wgrad_gemm")
    torch.cuda.synchronize()
    ref = 2 * (DYT[:, :P].double() @ XT[:, :P].double().T)
    err = float((dw.double() - ref).abs().max() / ref.abs().max())
    print("wgrad_gemm", K, N, P, err)
    assert err <= 5e-6, err
