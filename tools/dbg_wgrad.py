import os, sys, ctypes as C, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from dynrf_b200 import lib as L, ops
lib = L.load()
raw = lib._cdll
f = raw.dynrf_selftest_umma_wgrad_dbg
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]; f.restype = C.c_int
K = 128
for (lbo, sbo) in [(128, 2048), (2048, 128)]:
    A = torch.ones(128, K).cuda(); B = torch.ones(128, 64).cuda()
    Cc = torch.full((128, 64), float("nan"), device="cuda")
    rc = f(A.data_ptr(), B.data_ptr(), K, 1 | (lbo_flags := 0), Cc.data_ptr(), lbo, sbo, None); torch.cuda.synchronize()
    print("all-ones lbo", lbo, "sbo", sbo, "C min/max", float(Cc.min()), float(Cc.max()))
    for flags, name in ((0x100, "A only MN"), (0x200, "B only MN"), (0x300, "neither (K-major both)")):
        Cc = torch.full((128, 64), float("nan"), device="cuda")
        rc = f(A.data_ptr(), B.data_ptr(), K, 1 | flags, Cc.data_ptr(), lbo, sbo, None); torch.cuda.synchronize()
        print("   ", name, "C min/max", float(Cc.min()), float(Cc.max()))
