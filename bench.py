#!/usr/bin/env python
"""bench.py -- rays/s of the RoDynRF ray-march hot path on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU, rays sharded (weak scaling)
    python bench.py --config 3|4|5 ...                       # the other BASELINE.json configs (extra evidence lines)

Default workload (BASELINE.json configs[1], SURVEY.md 8d cfg 2): static TensorVMSplit + dynamic
TensorVMSplit_TimeEmbedding at grid [331,368,220] ("300^3"), 4096 NDC rays x 256 samples per GPU,
synthetic seeded rays in train order (spatially random origins, train.py:88-93) and reference-init random parameters.

One "step" = one pass of the hot path over one ray batch:
  --mode fwd   : sampleXYZ(is_train) -> static field -> dynamic field -> raw2outputs   (evaluation pass)
  --mode train : the PASS-1 structure of the reference train loop (train.py:1092-1162): the pass above with
                 autograd, an MSE loss on the three composited colours, backward to every dynamic + static
                 parameter, one Adam update (and, for N>1, one NCCL all-reduce of the flat gradient buffer).
  --mode train9: (--config 4) the 9-pass structure of the reference's full train step, see dynrf_b200/train9.py.
  --mode render: (--config 5) the no-grad evaluation_path chunk loop, frames sharded over ranks, frames/s.
The step is replayed as a CUDA graph after three eager warm-up steps (--no-graph launches it eagerly).
`--impl reference` runs the SAME step (train: forward + autograd backward + Adam) of the CPU oracle port on the host cores,
on a bounded ray sample sized by a calibration step so that the whole arm ends within its wall-clock budget; a watchdog
prints the line with the steps completed so far if the budget is exceeded (the arm can never end without a line).
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

METRIC = "rays/sec (train step)"

# BASELINE.json configs -> synthetic restatements
WORKLOADS = {
    2: dict(name="cfg2: Nvidia.txt Balloon1 shape, static+dynamic VM fields grid [331,368,220], 4096 NDC rays x 256 samples per GPU",
            grid=[331, 368, 220], n_rays=4096, n_samples=256, ray_type="ndc", static_mode="MLP_Fea", rays_grad=False),
    3: dict(name="cfg3: DAVIS.txt no-poses shape, grid [256,256,256], 4096 contracted rays x 192 samples per GPU, rays carry "
                 "requires_grad (pose / focal gradient through the static field)",
            grid=[256, 256, 256], n_rays=4096, n_samples=192, ray_type="contract", static_mode="MLP_Fea_TimeEmbedding",
            rays_grad=True),
    4: dict(name="cfg4: Nvidia_no_poses.txt shape, grid [706,786,471] ('640^3'), 4096 NDC rays x 256 samples",
            grid=[706, 786, 471], n_rays=4096, n_samples=256, ray_type="ndc", static_mode="MLP_Fea", rays_grad=True),
    5: dict(name="cfg5: render_path, 1920x1080 spiral (change_view_time), 270 frames, 8192-ray chunks, cfg-2 models, no_grad",
            grid=[331, 368, 220], n_rays=8192, n_samples=256, ray_type="ndc", static_mode="MLP_Fea", rays_grad=False),
}
BYTES_PER_SAMPLE_K1 = 588  # 576 B gathered + 12 B written, samples generated in-kernel
BYTES_PER_SAMPLE_K1_MAT = 605  # same with materialised xyz / z / valid inputs (the variant the train step runs)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


def config_dict(w, mode, world, graph):
    """The `config` object of the JSON line -- identical for the native and the reference arm (the driver compares them)."""
    return {"workload": w["name"], "mode": mode, "rays_per_gpu": w["n_rays"], "samples_per_ray": w["n_samples"], "grid": w["grid"],
            "ray_type": w["ray_type"], "ray_order": "train (spatially random origins, train.py:88-93)",
            "parallelism": f"rays sharded x{world}, parameters replicated", "cuda_graph": graph,
            "l2": "512 MiB buffer rewritten between timed iterations (L2 flush)"}


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


# ------------------------------------------------------------------------------------------- reference arm (oracle port)


class OracleStep:
    """The reference's own PyTorch path (oracle port, torch fp32) on a ray sample of the workload, on `device`
    ("cpu": the reference arm / cpu_baseline; "cuda": the eager-CUDA baseline, BASELINE.md "the real bar").

    mode "fwd": sampleXYZ -> both fields -> raw2outputs under no_grad.  mode "train": the same pass with autograd, the
    3-term MSE loss of the native arm, loss.backward() and an Adam update of every parameter (train.py:1092-1162)."""

    def __init__(self, w, n_rays, mode, seed=1, device="cpu"):
        from tests import helpers as H
        from oracle import rodynrf_oracle as O

        self.H, self.O, self.n, self.mode, self.w = H, O, n_rays, mode, w
        dev = torch.device(device)
        c = H.build_pair(w["grid"], n_rays, w["n_samples"], ray_type=w["ray_type"], static_mode=w["static_mode"], seed=seed)
        if dev.type != "cpu":
            for k in ("sd_static", "sd_dynamic"):
                c[k] = {n: v.to(dev) for n, v in c[k].items()}
            c["rays"], c["ts"] = c["rays"].to(dev), c["ts"].to(dev)
            for k in ("cs", "cd"):
                c[k].aabb = c[k].aabb.to(dev)
        self.case = c
        if mode == "train":
            self.ss = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["sd_static"].items()}
            self.sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["sd_dynamic"].items()}
            params = [v for v in list(self.ss.values()) + list(self.sd.values()) if v.requires_grad]
            self.opt = torch.optim.Adam(params, lr=1e-4, betas=(0.9, 0.99))
            self.target = torch.rand(n_rays, 3, generator=torch.Generator().manual_seed(5)).to(dev)

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

    def describe(self, secs):
        what = "forward + backward + Adam" if self.mode == "train" else "forward pass"
        return (f"{self.n} rays x {self.w['n_samples']} samples per step at the full grid {self.w['grid']}, {what}, torch fp32 "
                f"oracle port of the reference's PyTorch path ({secs:.1f} s of CPU work timed)")


def _ref_threads():
    """Host threads for the CPU arm: torch's own default (= physical cores) unless the launcher forced one thread
    (torchrun exports OMP_NUM_THREADS=1 to every rank); never more than 32 -- the port barely threads
    (1 thread 154 rays/s, 64 threads 192 rays/s measured in round 1) and oversubscribed boxes stalled the round-1 arm."""
    n = torch.get_num_threads()
    if n <= 1 or os.environ.get("OMP_NUM_THREADS") == "1":
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        try:
            import psutil

            phys = psutil.cpu_count(logical=False) or avail
        except Exception:
            phys = max(1, avail // 2)
        n = max(1, min(avail, phys))
    n = max(1, min(n, 32))
    torch.set_num_threads(n)
    return n


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    mode = "train" if args.mode in ("auto", "train", "train9") else "fwd"
    budget = float(os.environ.get("DYNRF_REF_BUDGET_S", "240"))
    t_start = time.perf_counter()
    cores = _ref_threads()
    state = {"ts": [], "n": None, "done": False, "ref": None}
    lock = threading.Lock()

    def emit(partial):
        with lock:
            if state["done"]:
                return
            state["done"] = True
            ts, n = list(state["ts"]), state["n"] or 0
            dt = sum(ts)
            v = (n * len(ts) / dt) if dt > 0 else 0.0
            sample = state["ref"].describe(dt) if state["ref"] is not None else "no step completed"
            if partial:
                sample += f"; wall-clock budget of {budget:.0f} s reached after {len(ts)} of {args.steps} timed steps"
            line = {
                "impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": (dt / len(ts) * 1e3) if ts else None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(w, mode, max(args.gpus, 1), not args.no_graph),
                "sample": {"rays_per_step": n, "steps_timed": len(ts), "partial": bool(partial)},
                "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            }
            print(json.dumps(line), flush=True)

    def watchdog():
        emit(partial=True)
        os._exit(0)

    timer = threading.Timer(budget, watchdog)
    timer.daemon = True
    timer.start()

    # calibration: one step of 128 rays (fixed costs -- Adam over 23 M parameters, gradient zeroing -- plus 128 rays of work)
    cal = OracleStep(w, 128, mode)
    t0 = time.perf_counter()
    cal.step()
    t_small = time.perf_counter() - t0
    t0 = time.perf_counter()
    cal.step()
    t_small = min(t_small, time.perf_counter() - t0)
    del cal
    left = budget * 0.8 - (time.perf_counter() - t_start)
    per_step = max(left, 1.0) / max(args.steps + args.warmup, 1)
    n = 1024
    while n > 128 and t_small * (n / 128.0) > per_step:   # upper bound: cost taken as linear in rays
        n //= 2
    sys.stderr.write(f"[reference arm] {cores} threads, calibration {t_small:.2f} s per 128-ray step -> {n} rays per step\n")
    ref = OracleStep(w, n, mode)
    state["n"], state["ref"] = n, ref
    for _ in range(args.warmup):
        ref.step()
    for i in range(args.steps):
        t0 = time.perf_counter()
        ref.step()
        dt = time.perf_counter() - t0
        with lock:
            state["ts"].append(dt)
        sys.stderr.write(f"[reference arm] step {i + 1}/{args.steps}: {dt:.2f} s\n")
    timer.cancel()
    emit(partial=False)


# ------------------------------------------------------------------------------------------- native arm


def _events(n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def _kernel_table():
    """profiles/r02_kernel_table.json: ncu-measured DRAM / L2 / L1 bytes per launch of the step's kernels on the final build
    (written by tools/kernel_table.py from an `ncu --set full` capture); None when absent."""
    p = os.path.join(ROOT, "profiles", "r02_kernel_table.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fwd", "train", "train9", "render"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-eager", action="store_true", help="skip the eager-CUDA run of the reference port")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--frames", type=int, default=0, help="--config 5: frames per rank to render (0 = 270 / world)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    w = WORKLOADS[args.config]
    if args.impl == "reference":
        run_reference(args, w)
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

    sys.path.insert(0, os.path.join(ROOT, "robust-dynrf_b200"))
    if args.config == 5 or args.mode == "render":
        from dynrf_b200 import bench_extra as BX

        BX.run_render(args, WORKLOADS[5], rank, world, dev, dist, ClockSampler, config_dict)
        return
    if args.mode == "train9":
        from dynrf_b200 import bench_extra as BX

        BX.run_train9(args, w, rank, world, dev, dist, ClockSampler, config_dict, METRIC)
        return

    sys.path.insert(0, os.path.join(ROOT, "robust-dynrf_b200"))
    from dynrf_b200 import synth as SY      # product-only workload construction
    from dynrf_b200 import lib as L
    from dynrf_b200 import train_step as TS

    mode = args.mode
    if mode == "auto":
        mode = "train" if TS.backward_available() else "fwd"
    N_RAYS, N_SAMPLES, GRID, RT = w["n_rays"], w["n_samples"], w["grid"], w["ray_type"]
    strong = os.environ.get("DYNRF_STRONG", "0") == "1"     # strong scaling: the 4096-ray GLOBAL batch is split over the ranks
    if strong:
        N_RAYS = max(N_RAYS // world, 1)

    # identical parameters on every rank (replicated), a different ray shard per rank
    static, dyn = SY.make_models(GRID, RT, w["static_mode"], dev, seed=1)
    g = torch.Generator().manual_seed(1000 + rank)
    rays_h, ts_h = SY.synth_rays(N_RAYS, RT, g)
    target_h = torch.rand(N_RAYS, 3, generator=g)
    rays_pin, ts_pin, target_pin = rays_h.pin_memory(), ts_h.pin_memory(), target_h.pin_memory()
    rays_d, ts_d, target_d = rays_h.to(dev), ts_h.to(dev), target_h.to(dev)
    out_pin = torch.empty(N_RAYS, 16, pin_memory=True)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    step = TS.Step(static, dyn, N_SAMPLES, RT, mode, world, graph=not args.no_graph, rays_grad=w["rays_grad"] and mode == "train")

    def one(rays, ts, target, fused=True):
        return step(rays, ts, target, fused=fused)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None  # samples through warm-up + both timed regions (under load)
    for _ in range(args.warmup):
        one(rays_d, ts_d, target_d)
    barrier()

    # Every step is a real optimisation step, so the model -- and with it the number of samples that pass the appearance mask --
    # drifts while the bench runs (round 1: `e2e` looked faster than `value` only because it ran later).  Both timed loops therefore
    # start from the SAME snapshot of the parameters and of the optimiser state.
    snap = None
    if mode == "train":
        snap = ([p.detach().clone() for p in step.params], step.opt.m.clone(), step.opt.v.clone(), step.opt.state.clone())

    def restore():
        if snap is None:
            return
        from dynrf_b200 import ops as _ops

        with torch.no_grad():
            for p, q in zip(step.params, snap[0]):
                p.copy_(q)
            step.opt.m.copy_(snap[1])
            step.opt.v.copy_(snap[2])
            step.opt.state.copy_(snap[3])
        _ops.PARAM_EPOCH[0] += 1
        torch.cuda.synchronize()

    # ---- device-resident timing
    ev = _events(args.steps)
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev[i][0].record()
        one(rays_d, ts_d, target_d)
        ev[i][1].record()
    barrier()
    launches = step.launches_per_step * args.steps  # counted per eager step; a graph replay launches the same kernels
    it_ms = [a.elapsed_time(b) for a, b in ev]
    t_ms = sum(it_ms)

    # ---- end-to-end: pinned host -> device -> step -> host
    for _ in range(2):
        one(rays_pin.to(dev, non_blocking=True), ts_pin.to(dev, non_blocking=True), target_pin.to(dev, non_blocking=True), fused=False)
    restore()
    barrier()
    ev2 = _events(args.steps)
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
    it2_ms = [a.elapsed_time(b) for a, b in ev2]
    t2_ms = sum(it2_ms)
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel times for the roofline lines: EAGER steps of the SAME mode (the kernel variants the timed step runs),
    # CUDA events around each entry point on the launching stream, L2 flushed between steps
    names = ["dynrf_static_sigma_weight_fwd", "dynrf_dyn_sigma_blend_fwd_tc2", "dynrf_dyn_mlp_bwd_tc", "dynrf_basis_bwd",
             "dynrf_head_bwd_tc", "dynrf_app_fwd_tc", "dynrf_warp_fwd_tc", "dynrf_warp_fwd_tc2", "dynrf_warp_bwd", "dynrf_warp_bwd_tc", "dynrf_wgrad_gemm",
             "dynrf_static_density_bwd", "dynrf_composite_fwd", "dynrf_composite_bwd", "dynrf_alpha_weight_fwd",
             "dynrf_alpha_weight_bwd", "dynrf_sample_points", "dynrf_render_static_fwd", "dynrf_render_composite_fwd"]
    prof = TS.Step(static, dyn, N_SAMPLES, RT, mode, 1, graph=False, rays_grad=w["rays_grad"] and mode == "train")
    prof(rays_d, ts_d, target_d, fused=False)
    L.profile_events = {k: [] for k in names}
    nprof = 5
    for i in range(nprof):
        flush.fill_(i & 0xFF)
        prof(rays_d, ts_d, target_d, fused=False)
    torch.cuda.synchronize()
    ktimes = {k: [a.elapsed_time(b) for a, b in v] for k, v in L.profile_events.items() if v}
    L.profile_events = None
    kms = {k: sum(v) / nprof for k, v in ktimes.items()}          # ms per step spent in this entry point
    klaunch = {k: len(v) / nprof for k, v in ktimes.items()}

    tt = torch.tensor([t_ms, t2_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t2_ms = float(tt[0]), float(tt[1])

    if rank == 0:
        pk = peaks()
        ktab = _kernel_table() or {}
        total_rays = N_RAYS * world * args.steps
        value = total_rays / (t_ms * 1e-3)
        e2e = total_rays / (t2_ms * 1e-3)
        P = N_RAYS * N_SAMPLES

        def traffic_of(kernel):
            row = (ktab.get("kernels") or {}).get(kernel)
            if not row:
                return None, None
            return row.get("dram_bytes"), {"source": ktab.get("source"), "dram_bytes": row.get("dram_bytes"),
                                           "l2_bytes": row.get("lts_bytes"), "l1_bytes": row.get("l1tex_bytes"),
                                           "ms_under_ncu": row.get("ms")}

        # K1: the north-star kernel (static density + sigma->alpha + transmittance scan), the variant THIS step runs
        k1_ms = kms.get("dynrf_static_sigma_weight_fwd")
        k1_bytes = BYTES_PER_SAMPLE_K1_MAT   # materialised xyz / z / valid (public-API path)
        k1_gbs = (P * k1_bytes / (k1_ms * 1e-3) / 1e9) if k1_ms else None
        k1_traffic, k1_meas = traffic_of("k_static_sigma_weight<0>")
        k1_bind = None
        try:
            k1t = json.load(open(os.path.join(ROOT, "profiles", "r02_k1_table.json")))
            gname = {(331, 368, 220): "300^3", (706, 786, 471): "640^3"}.get(tuple(GRID))
            row = next(r for r in k1t["rows"] if r["grid"] == gname and r["order"] == "train" and r["gen"] == 0)
            pct = {k: v for k, v in row["ncu_pct_of_peak"].items() if v is not None and k in ("l2", "l1tex", "sm")}
            k1_bind = {"unit": "L1TEX", "pct_of_peak_ncu": round(pct["l1tex"], 1), "l2_pct": round(pct["l2"], 1), "dram_bytes": row["dram_bytes"],
                       "l2_bytes": row["lts_bytes"], "l1_bytes": row["l1tex_bytes"], "source": "profiles/r02_k1_table.json"}
            if k1_traffic is None:
                k1_traffic = row["dram_bytes"]
        except Exception:
            pass
        roof_k1 = {
            "binding_unit": k1_bind,
            "kernel": "k_static_sigma_weight<GEN=0> (a6+a8+a9: static density lookup + sigma->alpha + transmittance scan), as launched by the timed step",
            "bound": "hbm", "achieved": k1_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": (k1_gbs / pk["hbm"]) if k1_gbs else None,
            "traffic": k1_traffic, "ms_per_launch": k1_ms, "bytes_per_sample": k1_bytes, "peak_source": pk["src"],
            "dram_frac": (k1_traffic / (k1_ms * 1e-3) / 1e9 / pk["hbm"]) if (k1_traffic and k1_ms) else None, "measured": k1_meas,
            "note": "achieved = ALGORITHMIC bytes (every tap counted as moved, SURVEY 8d) / launch time: an effective-bandwidth figure. The "
                    "density tables (10 MB at 300^3, 47 MB at 640^3) are L2-resident, so this is NOT a DRAM fraction (it can exceed 1); "
                    "dram_frac uses the ncu-measured DRAM bytes. profiles/r02_k1_table.json holds the 300^3 / 640^3 x train-order / "
                    "render-order table with DRAM, L2 and L1 bytes per launch and the fraction of each resource's peak."}
        # dominant kernel of the step by measured time
        shares = sorted(kms.items(), key=lambda kv: -kv[1])
        dom_name, dom_ms = shares[0] if shares else (None, None)
        ksum = sum(kms.values())
        dom_kernel = {"dynrf_dyn_mlp_bwd_tc": "k_dyn_mlp_bwd_tc<1>", "dynrf_basis_bwd": "k_basis_bwd<1>", "dynrf_head_bwd_tc": "k_head_bwd_tc<0>",
                      "dynrf_dyn_sigma_blend_fwd_tc2": "k_dyn_sigma_blend_tc2", "dynrf_app_fwd_tc": "k_app_fwd_tc<1>",
                      "dynrf_warp_bwd": "k_warp_bwd", "dynrf_warp_bwd_tc": "k_warp_bwd_tc", "dynrf_warp_fwd_tc2": "k_warp_fwd_tc2"}.get(dom_name, dom_name)
        dom_traffic, dom_meas = traffic_of(dom_kernel) if dom_name else (None, None)
        dom_ms_launch = (dom_ms / klaunch[dom_name]) if dom_name else None
        dom_row = (ktab.get("kernels") or {}).get(dom_kernel) or {}
        dom_pct = {k: v for k, v in (dom_row.get("pct_of_peak") or {}).items() if v is not None and k in ("dram", "lts", "l1tex", "sm")}
        dom_bind = max(dom_pct.items(), key=lambda kv: kv[1]) if dom_pct else None
        roofline = {
            "binding_unit": None if dom_bind is None else {"unit": {"lts": "L2", "l1tex": "L1TEX", "sm": "SM pipes", "dram": "HBM"}[dom_bind[0]],
                                                           "pct_of_peak_ncu": round(dom_bind[1], 1)},
            "kernel": f"{dom_kernel} (largest share of the timed step: {dom_ms:.3f} of {ksum:.3f} ms of this library's kernels per step)" if dom_name else None,
            "bound": "hbm", "unit": "GB/s", "peak": pk["hbm"], "peak_source": pk["src"],
            "achieved": (dom_traffic / (dom_ms_launch * 1e-3) / 1e9) if (dom_traffic and dom_ms_launch) else None,
            "frac": (dom_traffic / (dom_ms_launch * 1e-3) / 1e9 / pk["hbm"]) if (dom_traffic and dom_ms_launch) else None,
            "traffic": dom_traffic, "ms_per_launch": dom_ms_launch, "measured": dom_meas,
            "note": "achieved = ncu-measured DRAM bytes per launch (profiles/r02_kernel_table.json, final build) / live launch time. The kernel is "
                    "not DRAM-bound: the tables are L2-resident and the binding resources are L2 reductions / instruction issue -- see profiles/."}
        sb_ms = kms.get("dynrf_dyn_sigma_blend_fwd_tc2")
        sb_flops = P * 2 * 2 * (152 * 64 + 64)  # density + blending MLPs, MAC x 2
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(w, mode, world, not args.no_graph), rays_per_gpu=N_RAYS),
            "roofline": roofline,
            "roofline_k1": roof_k1,
            "roofline_tensor": {"kernel": "k_dyn_sigma_blend_tc2 (a16+a17: 3-scale lookups + two 152->64->1 MLPs, tcgen05 3-pass split-TF32)",
                                  "bound": "tensor", "achieved": sb_flops / (sb_ms * 1e-3) / 1e12 if sb_ms else None,
                                  "peak": pk["tf_sust"], "unit": "TFLOP/s",
                                  "frac": (sb_flops / (sb_ms * 1e-3) / 1e12 / pk["tf_sust"]) if sb_ms else None,
                                  "ms_per_launch": sb_ms, "peak_source": pk["src"],
                                  "note": "useful fp32-equivalent flops (the tensor cores execute 3x this as TF32 passes); bf16 dense peak quoted as the denominator the task names"},
            "kernel_ms_per_step": {k: round(v, 4) for k, v in shares},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": int(rays_pin.numel() * 4 + ts_pin.numel() * 4 + target_pin.numel() * 4),
                    "d2h_bytes_per_step": int(out_pin.numel() * 4), "ms_per_step": t2_ms / args.steps},
            "ms_per_iter": {"value": [round(x, 3) for x in it_ms], "e2e": [round(x, 3) for x in it2_ms],
                            "note": "every step is a real Adam update, so the model (and its active-sample count) drifts along each list; "
                                    "both lists start from the same parameter / optimiser snapshot"},
            "gpu_launches": launches, "clocks": clocks,
        }
        del prof
        if not args.no_cuda_eager and world == 1:
            # "the real bar" (BASELINE.md 3 / SURVEY 8d): the reference's eager PyTorch chain on this same B200 (oracle port, device=cuda)
            try:
                torch.cuda.empty_cache()
                n_eager = N_RAYS
                ref = OracleStep(w, n_eager, mode, device=dev)
                ref.step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    ref.step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                line["ref_cuda_eager"] = {"value": n_eager / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms,
                                          "sample": f"{n_eager} rays x {N_SAMPLES} samples, same step ({mode}), oracle port of the reference's "
                                                    "eager PyTorch chain on this GPU (library kernels: ATen grid_sampler / cuBLAS sgemm / cumprod), "
                                                    f"mean of {reps} after 1 warm-up",
                                          "speedup_value": value / world / (n_eager / (ms * 1e-3))}
                del ref
                torch.cuda.empty_cache()
            except Exception as e:  # a baseline must never take the line down
                line["ref_cuda_eager"] = {"error": repr(e)[:300]}
        if not args.no_cpu_baseline and world == 1:
            n_cpu = 1024 if args.config != 4 else 256
            _ref_threads()
            ref = OracleStep(w, n_cpu, mode)
            ref.step()
            tsamp = []
            for _ in range(3):
                t0 = time.perf_counter()
                ref.step()
                tsamp.append(time.perf_counter() - t0)
            line["cpu_baseline"] = {"value": n_cpu / min(tsamp), "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": ref.describe(sum(tsamp)) + ", best of 3 after 1 warm-up"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
