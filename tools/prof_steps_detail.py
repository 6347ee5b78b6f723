"""Per-step GPU time of the train step + per-entry-point breakdown, to see how the workload drifts as Adam moves the fields."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robust-dynrf_b200"))
from tests import helpers as H  # noqa: E402
from dynrf_b200 import lib as L  # noqa: E402
from dynrf_b200 import train_step as TS  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
case = H.build_pair([331, 368, 220], 4096, 256, ray_type="ndc", seed=1)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
target = torch.rand(4096, 3, device="cuda")
step = TS.Step(static, dyn, 256, "ndc", mode, 1)
names = [n for n in L.PROTOTYPES if n not in L._NO_LAUNCH]
for i in range(nsteps):
    L.profile_events = {n: [] for n in names}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step(rays, ts, target, fused=False)
    e1.record()
    torch.cuda.synchronize()
    ev = L.profile_events
    L.profile_events = None
    per = {n: sum(a.elapsed_time(b) for a, b in v) for n, v in ev.items() if v}
    tot = e0.elapsed_time(e1)
    top = sorted(per.items(), key=lambda kv: -kv[1])[:9]
    print(f"step {i}: {tot:.2f} ms, kernels {sum(per.values()):.2f} ms | " + ", ".join(f"{k.replace('dynrf_', '')} {v:.2f}" for k, v in top))
