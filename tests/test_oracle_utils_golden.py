"""Pin oracle/rodynrf_extras.py (a4, f1, f2, f3 restatements) against vectors computed by the unmodified reference
(tests/golden/make_golden_utils.py), and check the drop-in modules' host-side utilities (optimizer-group order,
upsample_volume_grid, vector_comp_diffs, save -> load) against the same vectors.  CPU only."""
import ast
import math
import os
import tempfile

import numpy as np
import pytest
import torch

from oracle import rodynrf_extras as X
from oracle import rodynrf_oracle as O
from tests.conftest import Golden
from tests import helpers as _H  # noqa: F401  (puts the drop-in package on sys.path)

TOL = dict(rtol=1e-6, atol=1e-6)


@pytest.fixture(scope="module")
def gu():
    return Golden("utils_small")


def _cfgs(g, act="relu"):
    aabb = g.t("aabb")
    cs = O.FieldCfg(aabb=aabb, near_far=(0.0, 1.0), shadingMode="MLP_Fea", fea_pe=2, dynamic=False, fea2denseAct=act)
    cd = O.FieldCfg(aabb=aabb, near_far=(0.0, 1.0), shadingMode="MLP_Fea_late_view", fea_pe=0, dynamic=True, fea2denseAct=act)
    return cs, cd


def _grads_close(sd, g, prefix, **kw):
    n = 0
    for k in g.raw:
        if k.startswith(prefix + "/grad/"):
            name = k[len(prefix) + 6:]
            assert sd[name].grad is not None, name
            torch.testing.assert_close(sd[name].grad, g.t(k), **{**TOL, **kw})
            n += 1
    assert n > 0, prefix


def test_tv_loss_raw_tensor(gu):
    x = gu.t("tv_x")
    torch.testing.assert_close(X.tv_loss(x), gu.t("tv_none"), **TOL)
    torch.testing.assert_close(X.tv_loss(x, "h"), gu.t("tv_h"), **TOL)
    torch.testing.assert_close(X.tv_loss(x, "w"), gu.t("tv_w"), **TOL)


@pytest.mark.parametrize("tag", ["s", "d"])
def test_regularisers(gu, tag):
    cs, cd = _cfgs(gu)
    cfg = cs if tag == "s" else cd
    fams = ["density", "app"] + (["blending"] if tag == "d" else [])
    for fam in fams:
        sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static" if tag == "s" else "dynamic").items()}
        v = X.tv_loss_vm(sd, fam)
        want = gu.t(f"reg_{tag}/TV_loss_{fam}")
        # the reference's TV over a (1,C,L,1) line divides 0 by 0 (utils.py:172-174): NaN, reproduced
        assert math.isnan(float(want)) and math.isnan(float(v.detach()))
        for i in range(3):
            sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static" if tag == "s" else "dynamic").items()}
            p = sd[f"{fam}_plane.{i}"]
            val = X.tv_loss(p)
            val.backward()
            torch.testing.assert_close(val, gu.t(f"tvp_{tag}/{fam}_plane.{i}"), **TOL)
            torch.testing.assert_close(p.grad, gu.t(f"tvp_{tag}/{fam}_plane.{i}/grad"), **TOL)
        sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static" if tag == "s" else "dynamic").items()}
        ln = sd[f"{fam}_line.0"]
        val = X.tv_loss(ln, "w")
        val.backward()
        torch.testing.assert_close(val, gu.t(f"tvp_{tag}/{fam}_line.0_h"), **TOL)
        torch.testing.assert_close(ln.grad, gu.t(f"tvp_{tag}/{fam}_line.0_h/grad"), **TOL)
    for fam in ["density"] + (["blending"] if tag == "d" else []):
        sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static" if tag == "s" else "dynamic").items()}
        v = X.dense_l1(cfg, sd, fam)
        v.backward()
        torch.testing.assert_close(v, gu.t(f"reg_{tag}/{fam}_L1"), **TOL)
        _grads_close(sd, gu, f"reg_{tag}/{fam}_L1", rtol=1e-5, atol=1e-9)
    if tag == "s":
        sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static").items()}
        v = X.vector_comp_diffs(sd)
        v.backward()
        torch.testing.assert_close(v, gu.t("reg_s/vector_comp_diffs"), **TOL)
        _grads_close(sd, gu, "reg_s/vector_comp_diffs")
        cs2, _ = _cfgs(gu, "softplus")
        sd = {k: v.clone().requires_grad_(True) for k, v in gu.sd("static").items()}
        v = X.dense_l1(cs2, sd, "density")
        v.backward()
        torch.testing.assert_close(v, gu.t("reg_s/density_L1_softplus"), rtol=1e-5, atol=1e-10)
        torch.testing.assert_close(sd["density_plane.0"].grad, gu.t("reg_s/density_L1_softplus/grad/density_plane.0"), rtol=1e-4, atol=1e-12)
        torch.testing.assert_close(sd["density_line.2"].grad, gu.t("reg_s/density_L1_softplus/grad/density_line.2"), rtol=1e-4, atol=1e-12)


def test_sample_ray(gu):
    cs, _ = _cfgs(gu)
    ro, rd = gu.t("sr_o"), gu.t("sr_d")
    step = float(gu.raw["sr_stepsize"])
    p, z, v = X.sample_ray(cs, step, ro, rd, 29, False)
    torch.testing.assert_close(p, gu.t("sr_eval_xyz"), **TOL)
    torch.testing.assert_close(z, gu.t("sr_eval_z"), **TOL)
    assert torch.equal(v, gu.t("sr_eval_valid"))
    torch.manual_seed(20211202 + 102)
    p, z, v = X.sample_ray(cs, step, ro, rd, 29, True)
    torch.testing.assert_close(p, gu.t("sr_train_xyz"), **TOL)
    torch.testing.assert_close(z, gu.t("sr_train_z"), **TOL)
    assert torch.equal(v, gu.t("sr_train_valid"))


@pytest.mark.parametrize("rt", ["ndc", "contract"])
def test_induce_flow_chain(gu, rt):
    pre = f"if_{rt}/"
    H, W = int(gu.raw["if_H"]), int(gu.raw["if_W"])
    f = gu.t("if_focal").clone().requires_grad_(True)
    w, pts, sf = (gu.t(pre + k).clone().requires_grad_(True) for k in ("w", "pts", "sf"))
    c2w, rays = gu.t(pre + "c2w").clone().requires_grad_(True), gu.t(pre + "rays").clone().requires_grad_(True)
    moved = pts + sf if rt == "ndc" else torch.clamp(pts + sf, min=-2.0 + 1e-6, max=2.0 - 1e-6)
    flow, disp = X.induce_flow(H, W, f, c2w, w, moved, gu.t(pre + "pts2d"), rays, rt)
    torch.testing.assert_close(flow, gu.t(pre + "flow"), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(disp, gu.t(pre + "disp"), rtol=1e-5, atol=1e-5)
    ((flow * gu.t(pre + "gflow")).sum() + (disp * gu.t(pre + "gdisp")).sum()).backward()
    for nme, t in (("w", w), ("pts", pts), ("sf", sf), ("c2w", c2w), ("focal", f), ("rays", rays)):
        want = gu.t(pre + "grad_" + nme)
        torch.testing.assert_close(t.grad, want, rtol=1e-4, atol=1e-4 * float(want.abs().max()))


def test_ndc_world_and_scene_flow(gu):
    H, W, f = int(gu.raw["if_H"]), int(gu.raw["if_W"]), gu.t("if_focal")
    x = gu.t("n2w_in")
    torch.testing.assert_close(X.ndc2world(x, H, W, f), gu.t("n2w_out"), **TOL)
    torch.testing.assert_close(X.world2ndc(X.ndc2world(x, H, W, f), H, W, f), gu.t("w2n_out"), **TOL)
    _, cd = _cfgs(gu)
    sd = gu.sd("dynamic")
    fwd, bwd = X.scene_flow(cd, sd, gu.t("sf_pts"), gu.t("sf_t"))
    torch.testing.assert_close(fwd, gu.t("sf_f"), **TOL)
    torch.testing.assert_close(bwd, gu.t("sf_b"), **TOL)
    outs = X.scene_flow_point_single(cd, sd, gu.t("sf_pts")[:, 0], gu.t("sf_t"))
    for i, o in enumerate(outs):
        torch.testing.assert_close(o, gu.t(f"sfs_{i}"), **TOL)


def test_ray_generation(gu):
    H, W = int(gu.raw["rg_H"]), int(gu.raw["rg_W"])
    ids = gu.t("rg_ids")
    for rt, key in (("ndc", "rg_ndc"), ("world", "rg_world")):
        pose9 = gu.t("rg_pose9").clone().requires_grad_(True)
        focal = gu.t("rg_focal").clone().requires_grad_(True)
        rays = X.generate_rays(ids, pose9, focal, H, W, rt)
        if rt == "ndc":
            torch.testing.assert_close(rays, gu.t("rg_ndc_rays"), rtol=1e-5, atol=1e-6)
        else:
            torch.testing.assert_close(rays[:, :3], gu.t("rg_world_o"), **TOL)
            torch.testing.assert_close(rays[:, 3:], gu.t("rg_world_d"), **TOL)
        (rays * gu.t("rg_grays")).sum().backward()
        torch.testing.assert_close(pose9.grad, gu.t(key + "_grad_pose9"), rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(focal.grad, gu.t(key + "_grad_focal"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(X.pose_to_mtx(gu.t("rg_pose9")), gu.t("rg_mtx"), **TOL)


# ------------------------------------------------------------------------------------------- product host utilities (CPU tensors)


def _product_models(gu, device="cpu"):
    import models.tensoRF as T
    from tests.helpers import COMMON

    aabb = gu.t("aabb")
    grid = [int(v) for v in gu.raw["grid"]]
    kw = dict(COMMON, near_far=[0.0, 1.0])
    static = T.TensorVMSplit(aabb, grid, 12, torch.device(device), shadingMode="MLP_Fea", fea_pe=2, **kw)
    dyn = T.TensorVMSplit_TimeEmbedding(aabb, grid, 12, torch.device(device), shadingMode="MLP_Fea_late_view", fea_pe=0, **kw)
    static.load_state_dict(gu.sd("static"))
    dyn.load_state_dict(gu.sd("dynamic"))
    return static, dyn


def test_product_optimizer_group_order(gu):
    """train.py:2331,2335 read groups 0 and 4 by index: the order is part of the boundary."""
    static, dyn = _product_models(gu)
    for tag, m in (("s", static), ("d", dyn)):
        want = ast.literal_eval(str(gu.raw[f"optgroups_{tag}"]))
        got = [([tuple(p.shape) for p in g["params"]], g["lr"]) for g in m.get_optparam_groups(0.02, 0.001)]
        assert got == want


def test_product_vector_comp_diffs_and_upsample(gu):
    static, dyn = _product_models(gu)
    v = static.vector_comp_diffs()
    v.backward()
    torch.testing.assert_close(v.detach(), gu.t("reg_s/vector_comp_diffs"), **TOL)
    for k in gu.raw:
        if k.startswith("reg_s/vector_comp_diffs/grad/"):
            name = k.split("/grad/")[1]
            torch.testing.assert_close(dict(static.named_parameters())[name].grad, gu.t(k), **TOL)
    res = [int(v) for v in gu.raw["up_res"]]
    static.upsample_volume_grid(res)
    dyn.upsample_volume_grid(res)
    for tag, m in (("static", static), ("dynamic", dyn)):
        want = gu.sd("up_" + tag) if False else {k[len(f"up_{tag}/"):]: gu.t(k) for k in gu.raw if k.startswith(f"up_{tag}/")}
        got = m.state_dict()
        for k, v in want.items():
            assert got[k].shape == v.shape, k
            torch.testing.assert_close(got[k].contiguous(), v, **TOL)
    assert static.nSamples == int(gu.raw["up_static_nSamples"])
    torch.testing.assert_close(static.stepSize, gu.t("up_static_stepSize"), **TOL)
    # the oracle's restatement of the resize agrees too
    up = X.upsample_vm(gu.sd("static"), ("density", "app"), res)
    for k in ("density_plane.1", "app_line.2"):
        torch.testing.assert_close(up[k], gu.t("up_static/" + k), **TOL)


def test_product_save_load_roundtrip(gu):
    """models/tensorBase.py:460-485: checkpoint = {kwargs, state_dict}; planes land NCHW-contiguous on disk and
    reload into the live channel-last parameters without changing a value."""
    import models.tensoRF as T

    static, _ = _product_models(gu)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ck.th")
        static.save(torch.zeros(3, 9), torch.tensor(1.0), path)
        ck = torch.load(path, weights_only=False)
        assert set(ck) == {"kwargs", "state_dict"} and "se3_poses" in ck["kwargs"] and "focal_ratio_refine" in ck["kwargs"]
        for k, v in ck["state_dict"].items():
            assert v.is_contiguous(), k
        kw = dict(ck["kwargs"])
        kw.pop("se3_poses"), kw.pop("focal_ratio_refine")
        kw["device"] = torch.device("cpu")
        fresh = T.TensorVMSplit(**kw)
        fresh.load(ck)
    a, b = static.state_dict(), fresh.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert fresh.density_plane[0].stride(1) == 1      # still channel-last after load
