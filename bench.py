#!/usr/bin/env python
"""bench.py -- rays/s of the RoDynRF ray-march hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, rays sharded (weak scaling)

Workload (BASELINE.json configs[1], SURVEY.md 8d cfg 2): static TensorVMSplit + dynamic
TensorVMSplit_TimeEmbedding at grid [331,368,220] ("300^3"), 4096 NDC rays x 256 samples per GPU,
synthetic seeded rays and reference-init random parameters.

One "step" = one pass of the hot path over one ray batch:
  --mode fwd   : sampleXYZ(is_train) -> static field -> dynamic field -> raw2outputs   (evaluation pass)
  --mode train : the PASS-1 structure of the reference train loop (train.py:1092-1162): the pass above with
                 autograd, an MSE loss on the three composited colours, backward to every dynamic + static
                 parameter, one Adam update (and, for N>1, one NCCL all-reduce of the flat gradient buffer).
The step is replayed as a CUDA graph after three eager warm-up steps (--no-graph launches it eagerly): at ~16 ms of GPU
work per step the ~100 kernel launches + ~300 small tensor ops of the eager step cost as much host time as the GPU needs.
`--impl reference` runs the SAME step (train: forward + autograd backward + Adam) of the CPU oracle port on all host threads.
`value` times the step with inputs resident in HBM; `e2e` times the same step through the public Python
API starting from PINNED HOST ray buffers (H2D inside the timed region) and ending with the composited
per-ray maps back on the host (D2H inside).  Between timed iterations a 512 MiB buffer is rewritten to
flush the 126 MB L2.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = [331, 368, 220]
N_RAYS = 4096
N_SAMPLES = 256
WORKLOAD = "cfg2: Nvidia.txt Balloon1 shape, static+dynamic VM fields grid [331,368,220], 4096 NDC rays x 256 samples per GPU"
METRIC = "rays/sec (train step)"
BYTES_PER_SAMPLE_K1 = 588  # 576 B gathered + 12 B written, samples generated in-kernel
BYTES_PER_SAMPLE_SBB = 2 * 3456 + 512 + 512 + 57  # k_dyn_mlp_bwd_tc<SAVED>, see DESIGN.md 6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


# ------------------------------------------------------------------------------------------- clocks


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- reference arm (CPU)


class CpuReference:
    """The reference's CPU path (oracle port, torch fp32 on all host threads) on a bounded sample of the workload.

    mode "fwd": sampleXYZ -> both fields -> raw2outputs under no_grad.  mode "train": the same pass with autograd, the
    3-term MSE loss of the native arm, loss.backward() and an Adam update of every parameter (train.py:1092-1162)."""

    def __init__(self, n_rays, mode, seed=1):
        from tests import helpers as H
        from oracle import rodynrf_oracle as O

        self.H, self.O, self.n, self.mode = H, O, n_rays, mode
        self.case = H.build_pair(GRID, n_rays, N_SAMPLES, ray_type="ndc", seed=seed)
        if mode == "train":
            c = self.case
            self.ss = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["sd_static"].items()}
            self.sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["sd_dynamic"].items()}
            params = [v for v in list(self.ss.values()) + list(self.sd.values()) if v.requires_grad]
            self.opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.99))
            self.target = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(5))

    def step(self):
        if self.mode != "train":
            self.H.oracle_pass(self.case)
            return
        O, c = self.O, self.case
        xyz, z, valid = O.sampleXYZ(c["cd"], c["rays"], c["n_samples"], c["ray_type"], True)
        fs = O.field_forward(c["cs"], self.ss, c["rays"], c["ts"], xyz, z, valid, c["ray_type"])
        fd = O.field_forward(c["cd"], self.sd, c["rays"], c["ts"], xyz, z, valid, c["ray_type"])
        o = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], c["rays"], False, c["ray_type"])
        t = self.target
        loss = ((o[0] - t) ** 2).mean() + ((o[4] - t) ** 2).mean() + ((o[8] - t) ** 2).mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()

    def time(self, reps, warm):
        for _ in range(warm):
            self.step()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            self.step()
            ts.append(time.perf_counter() - t0)
        return ts

    def describe(self, ts):
        what = "forward + backward + Adam" if self.mode == "train" else "forward pass"
        return (f"{self.n} rays x {N_SAMPLES} samples per step at the full grid {GRID}, {what}, torch fp32 CPU oracle port "
                f"({sum(ts):.1f} s of CPU work)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mode = "train" if args.mode in ("auto", "train") else "fwd"
    n = 1024
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)
    torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)))
    cores = torch.get_num_threads()
    ref = CpuReference(n, mode)
    ts = ref.time(args.steps, args.warmup)
    dt = sum(ts)
    v = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mode": mode, "note": "reference CPU implementation of the path (oracle port of the reference's PyTorch code)"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": ref.describe(ts)},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------- native arm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fwd", "train"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying CUDA graphs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py (native arm) needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from tests import helpers as H
    from dynrf_b200 import lib as L
    import renderer as R
    from dynrf_b200 import train_step as TS

    mode = args.mode
    if mode == "auto":
        mode = "train" if TS.backward_available() else "fwd"

    # identical parameters on every rank (replicated), a different ray shard per rank
    case = H.build_pair(GRID, N_RAYS, N_SAMPLES, ray_type="ndc", seed=1)
    static, dyn = H.native_models(case, dev)
    g = torch.Generator().manual_seed(1000 + rank)
    rays_h, ts_h = H.synth_rays(N_RAYS, "ndc", g)
    target_h = torch.rand(N_RAYS, 3, generator=g)
    rays_pin, ts_pin, target_pin = rays_h.pin_memory(), ts_h.pin_memory(), target_h.pin_memory()
    rays_d, ts_d, target_d = rays_h.to(dev), ts_h.to(dev), target_h.to(dev)
    out_pin = torch.empty(N_RAYS, 16, pin_memory=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    step = TS.Step(static, dyn, N_SAMPLES, "ndc", mode, world, graph=not args.no_graph)

    def one(rays, ts, target, fused=True):
        return step(rays, ts, target, fused=fused)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None  # samples through warm-up + timed region (under load)
    for _ in range(args.warmup):
        one(rays_d, ts_d, target_d)
    barrier()

    # ---- device-resident timing
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        one(rays_d, ts_d, target_d)
        ev[i][1].record()
    barrier()
    launches = step.launches_per_step * args.steps  # counted per eager step; a graph replay launches the same kernels
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end: pinned host -> device -> step -> host
    for _ in range(2):
        one(rays_pin.to(dev, non_blocking=True), ts_pin.to(dev, non_blocking=True), target_pin.to(dev, non_blocking=True), fused=False)
    barrier()
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev2[i][0].record()
        r = rays_pin.to(dev, non_blocking=True)
        t = ts_pin.to(dev, non_blocking=True)
        tg = target_pin.to(dev, non_blocking=True)
        o = one(r, t, tg, fused=False)
        out_pin.copy_(o[:N_RAYS], non_blocking=True)
        ev2[i][1].record()
    barrier()
    t2_ms = sum(a.elapsed_time(b) for a, b in ev2)

    # ---- per-kernel times for the roofline lines: eager forward passes (the same forward kernels the train step runs),
    # CUDA events around the two entry points on the launching stream, L2 flushed between passes
    DOM = "dynrf_dyn_sigma_blend_fwd_tc2"
    prof = TS.Step(static, dyn, N_SAMPLES, "ndc", "fwd", 1, graph=False)
    prof.forward_fused(rays_d, ts_d)
    L.profile_events = {"dynrf_static_sigma_weight_fwd": [], DOM: [], "dynrf_dyn_sigma_blend_fwd": []}
    for i in range(max(5, min(args.steps, 20))):
        flush.fill_(i & 0xFF)
        prof.forward_fused(rays_d, ts_d)
    torch.cuda.synchronize()
    k1 = [a.elapsed_time(b) for a, b in L.profile_events["dynrf_static_sigma_weight_fwd"]]
    ksb = [a.elapsed_time(b) for a, b in (L.profile_events[DOM] or L.profile_events["dynrf_dyn_sigma_blend_fwd"])]
    L.profile_events = None

    # dominant kernel of the train step (largest share in profiles/r01_launch_shares_train.json): eager train steps on a
    # throw-away Step (it re-points p.grad, so it runs after every timed loop)
    kdom = []
    if mode == "train":
        DOMB = "dynrf_dyn_mlp_bwd_tc"
        prof2 = TS.Step(static, dyn, N_SAMPLES, "ndc", "train", 1, graph=False)
        prof2(rays_d, ts_d, target_d)
        L.profile_events = {DOMB: []}
        for i in range(5):
            flush.fill_(i & 0xFF)
            prof2(rays_d, ts_d, target_d)
        torch.cuda.synchronize()
        kdom = [a.elapsed_time(b) for a, b in L.profile_events[DOMB]]
        L.profile_events = None

    tt = torch.tensor([t_ms, t2_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t2_ms = float(tt[0]), float(tt[1])

    if rank == 0:
        pk = peaks()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_k1_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")
        total_rays = N_RAYS * world * args.steps
        value = total_rays / (t_ms * 1e-3)
        e2e = total_rays / (t2_ms * 1e-3)
        k1_ms = sum(k1) / max(len(k1), 1)
        k1_gbs = N_RAYS * N_SAMPLES * BYTES_PER_SAMPLE_K1 / (k1_ms * 1e-3) / 1e9 if k1 else None
        ksb_ms = sum(ksb) / max(len(ksb), 1)
        sb_flops = N_RAYS * N_SAMPLES * 2 * 2 * (152 * 64 + 64)  # density + blending MLPs, MAC x 2
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": mode, "rays_per_gpu": N_RAYS, "samples_per_ray": N_SAMPLES,
                       "grid": GRID, "parallelism": f"rays sharded x{world}, parameters replicated",
                       "cuda_graph": not args.no_graph,
                       "l2": "512 MiB buffer rewritten between timed iterations (L2 flush); parameters 93 MB + per-sample tensors"},
            "roofline": {"kernel": "k_static_sigma_weight (static density + sigma->alpha + transmittance scan, a6+a8+a9)",
                         "bound": "hbm", "achieved": k1_gbs, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": (k1_gbs / pk["hbm"]) if k1_gbs else None, "traffic": traffic,
                         "ms_per_launch": k1_ms, "bytes_per_sample": BYTES_PER_SAMPLE_K1, "peak_source": pk["src"],
                         "note": "algorithmic bytes (every tap counted); the 10 MB density tables are L2-resident so DRAM traffic is far lower -- see profiles/"},
            "roofline_tensor": {"kernel": "k_dyn_sigma_blend_tc2 (a16+a17: 3-scale lookups + two 152->64->1 MLPs, tcgen05 3-pass split-TF32)",
                                  "bound": "tensor", "achieved": sb_flops / (ksb_ms * 1e-3) / 1e12 if ksb else None,
                                  "peak": pk["tf_sust"], "unit": "TFLOP/s",
                                  "frac": (sb_flops / (ksb_ms * 1e-3) / 1e12 / pk["tf_sust"]) if ksb else None,
                                  "ms_per_launch": ksb_ms, "peak_source": pk["src"],
                                  "note": "useful fp32-equivalent flops (the tensor cores execute 3x this as TF32 passes); bf16 dense peak quoted as the denominator the task names"},
            "roofline_dominant": None if not kdom else {
                "kernel": "k_dyn_mlp_bwd_tc<SAVED> (backward of a16+a17: dH from the saved activations, dgrad on tcgen05, warp-aggregated scatter from TMEM)",
                "bound": "hbm", "achieved": N_RAYS * N_SAMPLES * BYTES_PER_SAMPLE_SBB / (sum(kdom) / len(kdom) * 1e-3) / 1e9,
                "peak": pk["hbm"], "unit": "GB/s",
                "frac": N_RAYS * N_SAMPLES * BYTES_PER_SAMPLE_SBB / (sum(kdom) / len(kdom) * 1e-3) / 1e9 / pk["hbm"],
                "traffic": None, "ms_per_launch": sum(kdom) / len(kdom), "bytes_per_sample": BYTES_PER_SAMPLE_SBB,
                "peak_source": pk["src"],
                "note": "algorithmic bytes per sample: 3456 re-read by the scatter + 3456 reduced + 512 read (H^T) + 512 written (dH^T) + 57 of per-sample I/O; "
                        "the 21 MB of tables are L2-resident, so DRAM sees ~1 GB/launch of activation streams; the kernel is bound by the L2 reduction rate "
                        "and instruction issue, not by HBM"},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": int(rays_pin.numel() * 4 + ts_pin.numel() * 4 + target_pin.numel() * 4),
                    "d2h_bytes_per_step": int(out_pin.numel() * 4), "ms_per_step": t2_ms / args.steps},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            n_cpu = 1024
            ref = CpuReference(n_cpu, mode)
            tsamp = ref.time(reps=3, warm=1)
            line["cpu_baseline"] = {"value": n_cpu / min(tsamp), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": ref.describe(tsamp) + ", best of 3 after 1 warm-up"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
