import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests import helpers as H
from dynrf_b200 import ops
case = H.build_pair([331, 368, 220], 4096, 256, ray_type="ndc", seed=1)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
cfg = dyn.native_cfg("ndc")
zvec = ops.make_zvec(cfg, 256, "ndc", False, rays.device)
xyz, z, valid = ops.sample_points(cfg, rays, zvec)
for _ in range(3):
    ops.dynamic_forward_raw(dyn, rays, ts, xyz, z, valid, "ndc")
torch.cuda.synchronize()
