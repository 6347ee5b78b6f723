"""CPU-side vs GPU-side time of the train step: enqueue time (no sync), GPU event time, cProfile of the host code."""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "robust-dynrf_b200"))
from tests import helpers as H  # noqa: E402
from dynrf_b200 import train_step as TS  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
case = H.build_pair([331, 368, 220], 4096, 256, ray_type="ndc", seed=1)
static, dyn = H.native_models(case, "cuda:0")
rays, ts = case["rays"].cuda(), case["ts"].cuda()
target = torch.rand(4096, 3, device="cuda")
step = TS.Step(static, dyn, 256, "ndc", mode, 1, graph=(len(sys.argv) > 2 and sys.argv[2] == "graph"))
for _ in range(4):
    step(rays, ts, target, fused=False)
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        step(rays, ts, target, fused=False)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rep {rep}: enqueue {(t1 - t0) / 5 * 1e3:.2f} ms/step, gpu {e0.elapsed_time(e1) / 5:.2f} ms/step, wall {(t2 - t0) / 5 * 1e3:.2f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step(rays, ts, target, fused=False)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
print(torch.cuda.memory_summary(abbreviated=True)[:1500])
