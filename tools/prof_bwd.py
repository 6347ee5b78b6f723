import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from dynrf_b200 import train_step as TS
case = H.build_pair([331, 368, 220], 4096, 256, ray_type="ndc", seed=1)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
target = torch.rand(4096, 3, device="cuda")
step = TS.Step(static, dyn, 256, "ndc", "train", 1)
for _ in range(2):
    step(rays, ts, target)
torch.cuda.synchronize()
