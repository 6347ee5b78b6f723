#!/usr/bin/env python
"""Per-stage CUDA-event timings of the forward path at a BASELINE config shape (developer tool)."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import helpers as H  # noqa: E402
from dynrf_b200 import lib as L, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[331, 368, 220])
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--ray-type", default="ndc")
    a = ap.parse_args()
    dev = "cuda:0"
    case = H.build_pair(a.grid, a.rays, a.samples, ray_type=a.ray_type, seed=1)
    static, dyn = H.native_models(case, dev)
    rays, ts = case["rays"].to(dev), case["ts"].to(dev)
    rt = a.ray_type
    cfg = static.native_cfg(rt)
    zvec = ops.make_zvec(cfg, a.samples, rt, False, rays.device)
    xyz, z, valid = ops.sample_points(cfg, rays, zvec)
    res = {}
    res["sample_points"] = timeit(lambda: ops.sample_points(cfg, rays, zvec))
    res["static_sigma_weight(materialised)"] = timeit(lambda: ops.static_forward_raw(static, rays, xyz, z, valid, rt))
    res["static_sigma_weight(generated)"] = timeit(lambda: ops.static_forward_raw(static, rays, None, None, None, rt, zvec=zvec))
    st = ops.static_forward_raw(static, rays, xyz, z, valid, rt)
    res["static_app"] = timeit(lambda: ops.static_app_raw(static, st, rt))
    res["static_active_frac"] = float(st["cnt"]) / (a.rays * a.samples)
    res["dynamic_forward(all)"] = timeit(lambda: ops.dynamic_forward_raw(dyn, rays, ts, xyz, z, valid, rt))
    d = ops.dynamic_forward_raw(dyn, rays, ts, xyz, z, valid, rt)
    res["dynamic_active_frac"] = float(d["cnt"]) / (a.rays * a.samples)
    lib = L.load()
    ly = dyn.packed_layers()
    cfgd = dyn.native_cfg(rt)
    n, s = a.rays, a.samples
    p, S_ = ops._p, ops._stream
    res["dyn.time_embed"] = timeit(lambda: lib.dynrf_time_embed_fwd(p(ts), n, C.byref(ly["layer1"]), C.byref(ly["layer2"]), p(d["tfeat"]), S_()))
    res["dyn.warp"] = timeit(lambda: lib.dynrf_warp_fwd(C.byref(cfgd), p(xyz), n, s, p(d["tfeat"]), C.byref(ly["layer3"]), C.byref(ly["layer4"]), C.byref(ly["layer5"]), p(d["xyz_prime"]), S_()))
    vden, vbl, vapp = ops.vm_struct(dyn.density_plane, dyn.density_line), ops.vm_struct(dyn.blending_plane, dyn.blending_line), ops.vm_struct(dyn.app_plane, dyn.app_line)
    res["dyn.sigma_blend"] = timeit(lambda: lib.dynrf_dyn_sigma_blend_fwd(C.byref(cfgd), C.byref(vden), C.byref(vbl), C.byref(ly["d1"]), C.byref(ly["d2"]), C.byref(ly["b1"]), C.byref(ly["b2"]), p(xyz), p(d["xyz_prime"]), p(d["valid"]), p(d["tfeat"]), n, s, p(d["sigma"]), p(d["blend"]), S_()))
    res["dyn.app"] = timeit(lambda: lib.dynrf_dyn_app_fwd(C.byref(cfgd), C.byref(vapp), C.byref(ly["basis"]), C.byref(ly["h1"]), C.byref(ly["h2"]), C.byref(ly["h3"]), p(rays), p(xyz), p(d["xyz_prime"]), p(d["tfeat"]), n, s, p(d["idx"]), p(d["cnt"]), p(d["rgb"]), S_()))
    def dyn_only(tc, ver=2):
        ops.USE_TC, ops.TC_SB_VERSION = tc, ver
        return timeit(lambda: ops.dynamic_forward_raw(dyn, rays, ts, xyz, z, valid, rt))
    res["dynamic_forward(simt)"] = dyn_only(False)
    res["dynamic_forward(tc v1)"] = dyn_only(True, 1)
    res["dynamic_forward(tc v2)"] = dyn_only(True, 2)
    res["composite"] = timeit(lambda: ops.composite_raw(st["rgb"], st["sigma"], d["rgb"], d["sigma"], d["dists"], d["blend"], z, rays, False, rt))
    ns = a.rays * a.samples
    t_gen = res["static_sigma_weight(generated)"][1]
    res["K1_generated_GBs_algorithmic(588B)"] = ns * 588 / (t_gen * 1e-3) / 1e9
    t_mat = res["static_sigma_weight(materialised)"][1]
    res["K1_materialised_GBs_algorithmic(605B)"] = ns * 605 / (t_mat * 1e-3) / 1e9
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
