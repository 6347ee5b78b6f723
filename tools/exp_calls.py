"""Per-CALL GPU time (CUDA events) of every C-ABI entry point in one eager train step at cfg 2 (static and dynamic launches of
the same entry point listed separately, in call order).  Used with DYNRF_VARIANT=<experiment build> for timing experiments."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robust-dynrf_b200"))
from tests import helpers as H  # noqa: E402
from dynrf_b200 import lib as L  # noqa: E402
from dynrf_b200 import train_step as TS  # noqa: E402

case = H.build_pair([331, 368, 220], 4096, 256, ray_type="ndc", seed=1)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
target = torch.rand(4096, 3, device="cuda")
step = TS.Step(static, dyn, 256, "ndc", "train", 1, graph=False)
names = [n for n in L.PROTOTYPES if n not in L._NO_LAUNCH]
for i in range(3):
    L.profile_events = {n: [] for n in names}
    step(rays, ts, target, fused=False)
    torch.cuda.synchronize()
    ev = L.profile_events
    L.profile_events = None
per = {n: [round(a.elapsed_time(b), 4) for a, b in v] for n, v in ev.items() if v}
tot = sum(sum(v) for v in per.values())
print(os.environ.get("DYNRF_VARIANT", "product"), f"kernels {tot:.3f} ms |", {k.replace("dynrf_", ""): v for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])) if sum(v) > 0.05})
