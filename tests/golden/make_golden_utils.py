#!/usr/bin/env python
"""Generate tests/golden/utils_small.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden_utils.py [/root/reference]

Vectors for the boundary utilities and the "next" rows of SURVEY.md 8(f) / a4:
  * regularisers: utils.TVLoss through TV_loss_density/app/blending, density_L1, blending_L1, vector_comp_diffs
    (models/tensoRF.py:63-116, 378-444; utils.py:157-181) and their autograd gradients;
  * get_optparam_groups order, upsample_volume_grid (models/tensoRF.py:49-61, 198-232, 813-850);
  * TensorBase.sample_ray (models/tensorBase.py:501-522), eval and seeded train mode;
  * scene flow + induce_flow chain (models/tensoRF.py:446-519, renderer.py:1266-1392), with gradients;
  * ray generation: camera.pose_to_mtx, get_ray_directions_lean, get_rays_lean, ndc_rays_blender2
    (camera.py:8-15, dataLoader/ray_utils.py:53-140), with gradients to the 9-D poses and the focal length.
Same import harness as make_golden.py; nothing from the reference is copied, only inputs and outputs are stored.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = MG.SEED


def np_(x):
    return x.detach().cpu().numpy()


def main():
    T, B, R = MG.import_reference()
    import utils as U
    import camera as CAM
    from dataLoader import ray_utils as RU

    out = {}
    torch.manual_seed(SEED + 100)
    aabb = torch.tensor([[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]])
    grid = [13, 15, 11]
    static, dyn = MG.build_models(T, aabb, grid, [0.0, 1.0], "MLP_Fea")
    MG.put_sd(out, "sd_static", static)
    MG.put_sd(out, "sd_dynamic", dyn)
    out["aabb"], out["grid"] = np_(aabb), np.array(grid)

    # ---- regularisers (values + gradients)
    reg = U.TVLoss()
    for tag, m in (("s", static), ("d", dyn)):
        names = ["TV_loss_density", "TV_loss_app", "density_L1"]
        # vector_comp_diffs exists on the static class only (models/tensoRF.py:63-78)
        names += ["TV_loss_blending", "blending_L1"] if tag == "d" else ["vector_comp_diffs"]
        for nme in names:
            m.zero_grad(set_to_none=True)
            fn = getattr(m, nme)
            v = fn(reg) if nme.startswith("TV") else fn()
            v.backward()
            out[f"reg_{tag}/{nme}"] = np_(v)
            for pn, p in m.named_parameters():
                if p.grad is not None:
                    out[f"reg_{tag}/{nme}/grad/{pn}"] = np_(p.grad)
    # density_L1 with the default softplus activation as well (DAVIS uses relu; keep both activations pinned)
    static.fea2denseAct, keep = "softplus", static.fea2denseAct
    static.zero_grad(set_to_none=True)
    v = static.density_L1()
    v.backward()
    out["reg_s/density_L1_softplus"] = np_(v)
    out["reg_s/density_L1_softplus/grad/density_plane.0"] = np_(static.density_plane[0].grad)
    out["reg_s/density_L1_softplus/grad/density_line.2"] = np_(static.density_line[2].grad)
    static.fea2denseAct = keep
    static.zero_grad(set_to_none=True)
    # the TV_loss_* sums above are NaN in the reference: a line is (1,C,L,1), so count_w == 0 and utils.py:172-174 divides 0 by 0
    # (the reference's own TODO).  Pin the finite part too: TVLoss of every plane, with its gradient.
    for tag, m in (("s", static), ("d", dyn)):
        for fam in ("density", "app") + (("blending",) if tag == "d" else ()):
            for i, p in enumerate(getattr(m, fam + "_plane")):
                m.zero_grad(set_to_none=True)
                v = reg(p)
                v.backward()
                out[f"tvp_{tag}/{fam}_plane.{i}"] = np_(v)
                out[f"tvp_{tag}/{fam}_plane.{i}/grad"] = np_(p.grad)
            ln = getattr(m, fam + "_line")[0]
            m.zero_grad(set_to_none=True)
            v = reg(ln, "w")          # the well-defined half of a line's TV (h direction only)
            v.backward()
            out[f"tvp_{tag}/{fam}_line.0_h"] = np_(v)
            out[f"tvp_{tag}/{fam}_line.0_h/grad"] = np_(ln.grad)
    static.zero_grad(set_to_none=True)
    dyn.zero_grad(set_to_none=True)
    # TVLoss on a raw tensor incl. the ignore_axis variants
    x = torch.randn(1, 5, 9, 7)
    out["tv_x"] = np_(x)
    out["tv_none"], out["tv_h"], out["tv_w"] = np_(reg(x)), np_(reg(x, "h")), np_(reg(x, "w"))

    # ---- optimizer groups: parameter shapes per group, in order
    for tag, m in (("s", static), ("d", dyn)):
        groups = m.get_optparam_groups(0.02, 0.001)
        desc = []
        for g in groups:
            desc.append(([tuple(p.shape) for p in g["params"]], g["lr"]))
        out[f"optgroups_{tag}"] = np.array(repr(desc))

    # ---- sample_ray (a4)
    g = torch.Generator().manual_seed(SEED + 101)
    n = 37
    ro = torch.cat((torch.rand(n, 2, generator=g) * 3 - 1.5, -1.3 * torch.ones(n, 1)), -1)
    rd = torch.nn.functional.normalize(torch.cat((0.3 * torch.randn(n, 2, generator=g), torch.ones(n, 1)), -1), dim=-1)
    rd[3, 0] = 0.0   # exercises the rays_d == 0 guard
    out["sr_o"], out["sr_d"] = np_(ro), np_(rd)
    out["sr_stepsize"] = np_(static.stepSize)
    with torch.no_grad():
        p, z, v = static.sample_ray(ro, rd, is_train=False, N_samples=29)
        out["sr_eval_xyz"], out["sr_eval_z"], out["sr_eval_valid"] = np_(p), np_(z), np_(v)
        torch.manual_seed(SEED + 102)
        p, z, v = static.sample_ray(ro, rd, is_train=True, N_samples=29)
        out["sr_train_xyz"], out["sr_train_z"], out["sr_train_valid"] = np_(p), np_(z), np_(v)

    # ---- upsample_volume_grid
    res = [17, 19, 14]
    with MG.contextlib.redirect_stdout(MG.io.StringIO()):
        static.upsample_volume_grid(res)
        dyn.upsample_volume_grid(res)
    out["up_res"] = np.array(res)
    MG.put_sd(out, "up_static", static)
    MG.put_sd(out, "up_dynamic", dyn)
    out["up_static_nSamples"], out["up_static_stepSize"] = np.array(static.nSamples), np_(static.stepSize)

    # ---- scene flow + induced flow (f2)
    g = torch.Generator().manual_seed(SEED + 103)
    n, s = 23, 17
    H, W = 270, 480
    focal = torch.tensor(418.0)
    for rt in ("ndc", "contract"):
        if rt == "ndc":
            o = torch.cat((torch.rand(n, 2, generator=g) * 2 - 1, -torch.ones(n, 1)), -1)
            d = torch.cat((0.2 * torch.randn(n, 2, generator=g), 2 * torch.ones(n, 1)), -1)
            z = torch.linspace(0, 1, s)[None].repeat(n, 1)
            pts = o[:, None] + d[:, None] * z[..., None]
        else:
            o = 0.05 * torch.randn(n, 3, generator=g)
            d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
            z = torch.linspace(0.1, 30, s)[None].repeat(n, 1)
            pts = o[:, None] + d[:, None] * z[..., None]
            nrm = pts.abs().amax(-1, keepdim=True)
            pts = torch.where(nrm > 1, (2 - 1 / nrm) * pts / nrm, pts)
        rays = torch.cat((o, d), -1)
        w = torch.rand(n, s, generator=g)
        w = w / w.sum(-1, keepdim=True) * torch.rand(n, 1, generator=g)     # acc in (0,1)
        sf = 0.02 * torch.randn(n, s, 3, generator=g)
        pose9 = torch.cat((torch.eye(3)[:2].reshape(1, 6).repeat(n, 1) + 0.05 * torch.randn(n, 6, generator=g),
                           0.1 * torch.randn(n, 3, generator=g)), -1)
        c2w = CAM.pose_to_mtx(pose9)
        pts2d = torch.stack((torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H), -1)
        pre = f"if_{rt}/"
        out[pre + "rays"], out[pre + "w"], out[pre + "pts"], out[pre + "sf"] = np_(rays), np_(w), np_(pts), np_(sf)
        out[pre + "c2w"], out[pre + "pts2d"] = np_(c2w), np_(pts2d)
        wg, ptsg, sfg = w.clone().requires_grad_(True), pts.clone().requires_grad_(True), sf.clone().requires_grad_(True)
        c2wg, fg, raysg = c2w.clone().requires_grad_(True), focal.clone().requires_grad_(True), rays.clone().requires_grad_(True)
        if rt == "ndc":
            moved = ptsg + sfg
        else:
            moved = torch.clamp(ptsg + sfg, min=-2.0 + 1e-6, max=2.0 - 1e-6)   # train.py:1374-1378
        flow, disp = R.induce_flow(H, W, fg, c2wg, wg, moved, pts2d, raysg, ray_type=rt)
        out[pre + "flow"], out[pre + "disp"] = np_(flow), np_(disp)
        gw = torch.randn(flow.shape, generator=g)
        gd = torch.randn(disp.shape, generator=g)
        out[pre + "gflow"], out[pre + "gdisp"] = np_(gw), np_(gd)
        ((flow * gw).sum() + (disp * gd).sum()).backward()
        for nme, t in (("w", wg), ("pts", ptsg), ("sf", sfg), ("c2w", c2wg), ("focal", fg), ("rays", raysg)):
            out[pre + "grad_" + nme] = np_(t.grad)
    out["if_H"], out["if_W"], out["if_focal"] = np.array(H), np.array(W), np_(focal)
    # standalone pieces
    ptn = torch.rand(31, 3, generator=g) * 2 - 1
    out["n2w_in"] = np_(ptn)
    out["n2w_out"] = np_(R.NDC2world(ptn, H, W, focal))
    out["w2n_out"] = np_(R.world2NDC(R.NDC2world(ptn, H, W, focal), H, W, focal))
    # scene-flow MLP
    up_pts = torch.rand(9, 7, 3, generator=g) * 2 - 1
    tsamp = torch.rand(9, generator=g) * 2 - 1
    out["sf_pts"], out["sf_t"] = np_(up_pts), np_(tsamp)
    with torch.no_grad():
        f, b = dyn.get_forward_backward_scene_flow(up_pts, tsamp)
        out["sf_f"], out["sf_b"] = np_(f), np_(b)
        a1, a2, a3, a4 = dyn.get_forward_backward_scene_flow_point_single(up_pts[:, 0], tsamp)
        out["sfs_0"], out["sfs_1"], out["sfs_2"], out["sfs_3"] = np_(a1), np_(a2), np_(a3), np_(a4)

    # ---- ray generation (f1): ids -> pixel -> camera dirs -> world rays -> NDC, gradients to poses / focal
    g = torch.Generator().manual_seed(SEED + 104)
    F_, H, W = 5, 27, 48
    ids = torch.randint(0, F_ * H * W, (64,), generator=g)
    pose9 = torch.cat((torch.eye(3)[:2].reshape(1, 6).repeat(F_, 1) + 0.05 * torch.randn(F_, 6, generator=g),
                       0.1 * torch.randn(F_, 3, generator=g)), -1).requires_grad_(True)
    focal = torch.tensor(41.8, requires_grad=True)
    col, row, view = ids % W, (ids // W) % H, ids // (W * H)      # train.py:96-103 ids2pixel
    poses_mtx = CAM.pose_to_mtx(pose9)
    dirs = RU.get_ray_directions_lean(col, row, [focal, focal], [W / 2, H / 2])
    ro, rd = RU.get_rays_lean(dirs, poses_mtx[view])
    out["rg_ids"], out["rg_pose9"], out["rg_focal"] = np_(ids), np_(pose9), np_(focal)
    out["rg_F"], out["rg_H"], out["rg_W"] = np.array(F_), np.array(H), np.array(W)
    out["rg_mtx"], out["rg_dirs"], out["rg_world_o"], out["rg_world_d"] = np_(poses_mtx), np_(dirs), np_(ro), np_(rd)
    no, nd = RU.ndc_rays_blender2(H, W, [focal, focal], 1.0, ro, rd)
    rays = torch.cat((no, nd), -1)
    out["rg_ndc_rays"] = np_(rays)
    gr = torch.randn(rays.shape, generator=g)
    out["rg_grays"] = np_(gr)
    (rays * gr).sum().backward()
    out["rg_ndc_grad_pose9"], out["rg_ndc_grad_focal"] = np_(pose9.grad), np_(focal.grad)
    pose9.grad = None
    focal.grad = None
    poses_mtx = CAM.pose_to_mtx(pose9)
    dirs = RU.get_ray_directions_lean(col, row, [focal, focal], [W / 2, H / 2])
    ro, rd = RU.get_rays_lean(dirs, poses_mtx[view])
    (torch.cat((ro, rd), -1) * gr).sum().backward()
    out["rg_world_grad_pose9"], out["rg_world_grad_focal"] = np_(pose9.grad), np_(focal.grad)

    path = os.path.join(OUT, "utils_small.npz")
    np.savez_compressed(path, **out)
    print("utils_small ->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", len(out), "arrays")


if __name__ == "__main__":
    main()
