import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from dynrf_b200 import lib as L, ops
lib = L.load()
for K, N, P in [(224, 128, 1048576), (128, 128, 1048576)]:
  for pad in (0, 288, 1056, 8224):
    ld = P + pad
    XT = torch.randn(K, ld, device="cuda"); DYT = torch.randn(N, ld, device="cuda"); dw = torch.zeros(N, K, device="cuda")
    f = lambda: lib.dynrf_wgrad_gemm(ops._p(XT), ld, K, ops._p(DYT), ld, N, P, None, ops._p(dw), K, None)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = (K + N) * P * 4 / 1e9
    print(f"wgrad K={K} N={N} P={P} pad={pad}: {ms:.3f} ms, {gb/ms*1e3:.0f} GB/s read, {2*K*N*P/ms/1e9:.1f} TFLOP/s")
