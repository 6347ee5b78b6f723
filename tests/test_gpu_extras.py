"""-m gpu: the kernels of the rows next to the ray-march path (a4, f1, f2, f3) against vectors produced by the UNMODIFIED
reference (tests/golden/utils_small.npz) and against the extras oracle at larger sizes."""
import math

import pytest
import torch

from oracle import rodynrf_extras as X
from oracle import rodynrf_oracle as O
from tests import helpers as H
from tests.conftest import Golden

gpu = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gu():
    return Golden("utils_small")


def _models(gu, act="relu", shift=-10.0):
    import models.tensoRF as T

    dev = torch.device(DEV)
    aabb = gu.t("aabb").to(dev)
    grid = [int(v) for v in gu.raw["grid"]]
    kw = dict(H.COMMON, near_far=[0.0, 1.0], fea2denseAct=act, density_shift=shift)
    static = T.TensorVMSplit(aabb, grid, 12, dev, shadingMode="MLP_Fea", fea_pe=2, **kw)
    dyn = T.TensorVMSplit_TimeEmbedding(aabb, grid, 12, dev, shadingMode="MLP_Fea_late_view", fea_pe=0, **kw)
    static.load_state_dict(gu.sd("static"))
    dyn.load_state_dict(gu.sd("dynamic"))
    return static, dyn


class _TVLoss(torch.nn.Module):
    """stand-in for the reference's utils.TVLoss instance train.py passes as `reg` (only the attribute is read)."""

    def __init__(self, w=1.0):
        super().__init__()
        self.TVLoss_weight = w


# ------------------------------------------------------------------------------------------- a4


@gpu
def test_sample_ray_matches_reference(gu):
    static, _ = _models(gu)
    ro, rd = gu.t("sr_o").to(DEV), gu.t("sr_d").to(DEV)
    assert abs(float(static.stepSize) - float(gu.raw["sr_stepsize"])) <= 1e-7
    with torch.no_grad():
        p, z, v = static.sample_ray(ro, rd, is_train=False, N_samples=29)
    torch.testing.assert_close(p.cpu(), gu.t("sr_eval_xyz"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(z.cpu(), gu.t("sr_eval_z"), rtol=1e-6, atol=1e-6)
    assert torch.equal(v.cpu(), gu.t("sr_eval_valid"))
    # train mode: same per-ray jitter fed to the oracle (the device RNG differs from the CPU generator)
    torch.manual_seed(5)
    with torch.no_grad():
        p, z, v = static.sample_ray(ro, rd, is_train=True, N_samples=29)
    torch.manual_seed(5)
    jit = torch.rand(ro.shape[0], device=DEV).cpu()[:, None]
    cs = O.FieldCfg(aabb=gu.t("aabb"), near_far=(0.0, 1.0))
    pr, zr, vr = X.sample_ray(cs, float(gu.raw["sr_stepsize"]), ro.cpu(), rd.cpu(), 29, True, jitter=jit)
    torch.testing.assert_close(p.cpu(), pr, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(z.cpu(), zr, rtol=1e-6, atol=1e-6)
    assert int((v.cpu() != vr).sum()) <= 1     # a point within 1 ulp of the box face may flip


# ------------------------------------------------------------------------------------------- f1


@gpu
@pytest.mark.parametrize("rt", ["ndc", "world"])
def test_ray_generation_matches_reference(gu, rt):
    import renderer as R

    Hh, Ww = int(gu.raw["rg_H"]), int(gu.raw["rg_W"])
    ids = gu.t("rg_ids").to(DEV)
    pose9 = gu.t("rg_pose9").to(DEV).requires_grad_(True)
    focal = gu.t("rg_focal").to(DEV).requires_grad_(True)
    rays = R.generate_rays(ids, pose9, focal, Hh, Ww, rt)
    if rt == "ndc":
        torch.testing.assert_close(rays.detach().cpu(), gu.t("rg_ndc_rays"), rtol=2e-5, atol=2e-6)
        key = "rg_ndc"
    else:
        torch.testing.assert_close(rays[:, :3].detach().cpu(), gu.t("rg_world_o"), rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(rays[:, 3:].detach().cpu(), gu.t("rg_world_d"), rtol=1e-5, atol=1e-6)
        key = "rg_world"
    (rays * gu.t("rg_grays").to(DEV)).sum().backward()
    torch.testing.assert_close(pose9.grad.cpu(), gu.t(key + "_grad_pose9"), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(focal.grad.cpu(), gu.t(key + "_grad_focal"), rtol=1e-3, atol=1e-4)


@gpu
def test_ray_generation_large_against_oracle():
    import renderer as R

    g = torch.Generator().manual_seed(3)
    F_, Hh, Ww = 24, 1080, 1920
    ids = torch.randint(0, F_ * Hh * Ww, (4096,), generator=g)
    pose9 = torch.cat((torch.eye(3)[:2].reshape(1, 6).repeat(F_, 1) + 0.05 * torch.randn(F_, 6, generator=g), 0.1 * torch.randn(F_, 3, generator=g)), -1)
    focal = torch.tensor(1000.0)
    gr = torch.randn(4096, 6, generator=g)
    p_ref, f_ref = pose9.clone().requires_grad_(True), focal.clone().requires_grad_(True)
    r_ref = X.generate_rays(ids, p_ref, f_ref, Hh, Ww, "ndc")
    (r_ref * gr).sum().backward()
    p, f = pose9.to(DEV).requires_grad_(True), focal.to(DEV).requires_grad_(True)
    r = R.generate_rays(ids.to(DEV), p, f, Hh, Ww, "ndc")
    (r * gr.to(DEV)).sum().backward()
    torch.testing.assert_close(r.detach().cpu(), r_ref.detach(), rtol=1e-4, atol=1e-5)
    assert float((p.grad.cpu() - p_ref.grad).norm() / p_ref.grad.norm()) <= 1e-3
    assert abs(float(f.grad) - float(f_ref.grad)) <= 1e-3 * abs(float(f_ref.grad)) + 1e-5


# ------------------------------------------------------------------------------------------- f2


@gpu
@pytest.mark.parametrize("rt", ["ndc", "contract"])
@pytest.mark.parametrize("fused_sf", [False, True])
def test_induce_flow_matches_reference(gu, rt, fused_sf):
    import renderer as R

    pre = f"if_{rt}/"
    Hh, Ww = int(gu.raw["if_H"]), int(gu.raw["if_W"])
    f = gu.t("if_focal").to(DEV).requires_grad_(True)
    w, pts, sf = (gu.t(pre + k).to(DEV).requires_grad_(True) for k in ("w", "pts", "sf"))
    c2w, rays = gu.t(pre + "c2w").to(DEV).requires_grad_(True), gu.t(pre + "rays").to(DEV).requires_grad_(True)
    clamp = None if rt == "ndc" else (-2.0 + 1e-6, 2.0 - 1e-6)
    if fused_sf:
        flow, disp = R.induce_flow(Hh, Ww, f, c2w, w, pts, gu.t(pre + "pts2d").to(DEV), rays, ray_type=rt, scene_flow=sf, clamp=clamp)
    else:
        moved = pts + sf if clamp is None else torch.clamp(pts + sf, min=clamp[0], max=clamp[1])
        flow, disp = R.induce_flow(Hh, Ww, f, c2w, w, moved, gu.t(pre + "pts2d").to(DEV), rays, ray_type=rt)
    torch.testing.assert_close(flow.detach().cpu(), gu.t(pre + "flow"), rtol=1e-4, atol=2e-3)
    torch.testing.assert_close(disp.detach().cpu(), gu.t(pre + "disp"), rtol=1e-4, atol=1e-5)
    ((flow * gu.t(pre + "gflow").to(DEV)).sum() + (disp * gu.t(pre + "gdisp").to(DEV)).sum()).backward()
    for nme, t in (("w", w), ("pts", pts), ("sf", sf), ("c2w", c2w), ("focal", f), ("rays", rays)):
        want = gu.t(pre + "grad_" + nme)
        torch.testing.assert_close(t.grad.cpu(), want, rtol=2e-3, atol=2e-3 * float(want.abs().max()))


@gpu
def test_weighted_point_fullsize():
    from dynrf_b200 import extras as E

    g = torch.Generator().manual_seed(1)
    n, s = 4096, 256
    w = torch.rand(n, s, generator=g) / s
    pts = torch.randn(n, s, 3, generator=g)
    sf = 0.01 * torch.randn(n, s, 3, generator=g)
    wd, pd, sd = (t.to(DEV).requires_grad_(True) for t in (w, pts, sf))
    pm, acc = E.weighted_point(wd, pd, sd, (-1.5, 1.5))
    mv = torch.clamp(pts + sf, -1.5, 1.5)
    torch.testing.assert_close(pm.detach().cpu(), (w[..., None] * mv).sum(1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(acc.detach().cpu(), w.sum(1), rtol=1e-5, atol=1e-6)
    gp, ga = torch.randn(n, 3, generator=g), torch.randn(n, generator=g)
    ((pm * gp.to(DEV)).sum() + (acc * ga.to(DEV)).sum()).backward()
    wr, pr = w.clone().requires_grad_(True), pts.clone().requires_grad_(True)
    mr = torch.clamp(pr + sf, -1.5, 1.5)
    (((wr[..., None] * mr).sum(1) * gp).sum() + (wr.sum(1) * ga).sum()).backward()
    torch.testing.assert_close(wd.grad.cpu(), wr.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(pd.grad.cpu(), pr.grad, rtol=1e-5, atol=1e-7)
    assert torch.equal(pd.grad, sd.grad)


# ------------------------------------------------------------------------------------------- f3


@gpu
@pytest.mark.parametrize("tag", ["s", "d"])
def test_tv_and_l1_match_reference(gu, tag):
    static, dyn = _models(gu)
    m = static if tag == "s" else dyn
    reg = _TVLoss()
    params = dict(m.named_parameters())
    fams = ["density", "app"] + (["blending"] if tag == "d" else [])
    for fam in fams:
        m.zero_grad(set_to_none=True)
        v = getattr(m, f"TV_loss_{fam}")(reg)
        # a (1,C,L,1) line makes utils.TVLoss divide 0 by 0: the reference's value is NaN, its gradient finite -- both reproduced
        assert math.isnan(float(v)) and math.isnan(float(gu.raw[f"reg_{tag}/TV_loss_{fam}"]))
        v.backward()
        n = 0
        for k in gu.raw:
            pre = f"reg_{tag}/TV_loss_{fam}/grad/"
            if k.startswith(pre):
                torch.testing.assert_close(params[k[len(pre):]].grad.cpu().contiguous(), gu.t(k), rtol=1e-5, atol=1e-9)
                n += 1
        assert n == 6
    # the finite part of the value: one plane at a time through the same kernel
    from dynrf_b200 import extras as E

    for fam in fams:
        for i, p in enumerate(getattr(m, fam + "_plane")):
            val = E._TvFn.apply(1.0, [1.0], p)
            torch.testing.assert_close(val.detach().cpu(), gu.t(f"tvp_{tag}/{fam}_plane.{i}"), rtol=1e-5, atol=1e-8)
    for fam in ["density"] + (["blending"] if tag == "d" else []):
        m.zero_grad(set_to_none=True)
        v = getattr(m, f"{fam}_L1")()
        v.backward()
        torch.testing.assert_close(v.detach().cpu(), gu.t(f"reg_{tag}/{fam}_L1"), rtol=1e-5, atol=1e-8)
        n = 0
        for k in gu.raw:
            pre = f"reg_{tag}/{fam}_L1/grad/"
            if k.startswith(pre):
                want = gu.t(k)
                torch.testing.assert_close(params[k[len(pre):]].grad.cpu().contiguous(), want, rtol=1e-4, atol=1e-5 * float(want.abs().max()))
                n += 1
        assert n == 6


@gpu
def test_density_l1_softplus_matches_reference(gu):
    static, _ = _models(gu, act="softplus")
    v = static.density_L1()
    v.backward()
    torch.testing.assert_close(v.detach().cpu(), gu.t("reg_s/density_L1_softplus"), rtol=1e-4, atol=1e-10)
    for nme in ("density_plane.0", "density_line.2"):
        want = gu.t(f"reg_s/density_L1_softplus/grad/{nme}")
        got = dict(static.named_parameters())[nme].grad.cpu().contiguous()
        torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-5 * float(want.abs().max()))


@gpu
def test_tv_and_l1_midsize_against_oracle():
    """[61,70,45] grid (ragged against every tile size of the kernels), softplus with a live shift, TV weight 0.5."""
    case = H.build_pair([61, 70, 45], 8, 8, seed=3, act="softplus", density_shift=-0.5)
    static, dyn = H.native_models(case, DEV)
    cd = case["cd"]
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd_dynamic"].items()}
    ref = X.dense_l1(cd, sd, "blending")
    ref.backward()
    got = dyn.blending_L1()
    got.backward()
    torch.testing.assert_close(got.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-9)
    for i in range(3):
        for fam in ("blending_plane", "blending_line"):
            want = sd[f"{fam}.{i}"].grad
            have = getattr(dyn, fam)[i].grad.cpu().contiguous()
            assert float((have - want).norm() / want.norm()) <= 1e-4, (fam, i)
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd_dynamic"].items()}
    for i in range(3):
        (X.tv_loss(sd[f"app_plane.{i}"]) * 0.5 * 1e-2 + X.tv_loss(sd[f"app_line.{i}"], "w") * 0.5 * 1e-3).backward()
    dyn.zero_grad(set_to_none=True)
    dyn.TV_loss_app(_TVLoss(0.5)).backward()
    for i in range(3):
        for fam in ("app_plane", "app_line"):
            want = sd[f"{fam}.{i}"].grad
            have = getattr(dyn, fam)[i].grad.cpu().contiguous()
            assert float((have - want).norm() / want.norm()) <= 1e-5, (fam, i)


@gpu
def test_tv_l1_fullsize_run_and_time():
    """cfg-2 grid: one launch each; prints the streaming rate (also covered by bench --config ... lines)."""
    case = H.build_pair([331, 368, 220], 8, 8, seed=2)
    static, dyn = H.native_models(case, DEV)
    reg = _TVLoss()
    for _ in range(2):
        dyn.zero_grad(set_to_none=True)
        (dyn.TV_loss_app(reg) + dyn.TV_loss_density(reg) + dyn.TV_loss_blending(reg)).backward()
        l1 = dyn.density_L1()
        l1.backward()
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    dyn.zero_grad(set_to_none=True)
    e0.record()
    tv = dyn.TV_loss_app(reg)
    e1.record()
    l1 = dyn.density_L1()
    e2.record()
    torch.cuda.synchronize()
    nbytes = sum(p.numel() * 4 for p in list(dyn.app_plane) + list(dyn.app_line))
    print(f"TV(app) fused fwd+grad: {e0.elapsed_time(e1):.3f} ms for {nbytes / 1e6:.1f} MB of tables (read once, gradient written once); "
          f"density_L1 fwd+grad over {331 * 368 * 220 / 1e6:.1f} M voxels: {e1.elapsed_time(e2):.3f} ms")
    assert math.isfinite(float(l1)) and float(l1) > 0
