"""Pin oracle/rodynrf_oracle.py against vectors computed by the unmodified reference
(tests/golden/make_golden.py).  CPU only.  Same torch, same ops -> expect bit equality
for most outputs; a 1e-6 tolerance absorbs op-ordering differences."""
import numpy as np
import torch

from oracle import rodynrf_oracle as O

TOL = dict(rtol=1e-6, atol=1e-6)


def close(a, b, **kw):
    kw = {**TOL, **kw}
    torch.testing.assert_close(a, b, **kw)


def test_pure_functions(golden_pure):
    g = golden_pure
    close(O.positional_encoding(g.t("pe_x"), 6), g.t("pe_6"))
    close(O.positional_encoding(g.t("pe_x"), 10), g.t("pe_10"))
    a, w, last = O.raw2alpha(g.t("r2a_sigma"), g.t("r2a_dist"))
    close(a, g.t("r2a_alpha"))
    close(w, g.t("r2a_weight"))
    close(last, g.t("r2a_last"))


def test_sampling(golden_case):
    g = golden_case
    cs, cd = g.cfgs()
    rt = g.meta["ray_type"]
    s = int(g.raw["n_samples"])
    xyz, z, valid = O.sampleXYZ(cd, g.t("rays"), s, rt, False)
    close(xyz, g.t("xyz"))
    close(z, g.t("z_vals"))
    assert torch.equal(valid, g.t("ray_valid"))
    torch.manual_seed(20211202 + 7)  # same seed as make_golden.py: RNG consumed identically
    xyz, z, valid = O.sampleXYZ(cd, g.t("rays"), s, rt, True)
    close(xyz, g.t("train_xyz"))
    close(z, g.t("train_z"))
    assert torch.equal(valid, g.t("train_valid"))


def test_unit_features(golden_case):
    g = golden_case
    cs, cd = g.cfgs()
    ss, sdd = g.sd("static"), g.sd("dynamic")
    xn, t = g.t("u_xn"), g.t("u_t")
    close(O.static_density_feature(ss, xn), g.t("u_s_density"))
    close(O.static_app_feature(ss, xn), g.t("u_s_app"))
    close(O.dynamic_density_feature(cd, sdd, xn, t), g.t("u_d_density"))
    close(O.dynamic_blending_feature(cd, sdd, xn, t), g.t("u_d_blending"))
    close(O.dynamic_app_feature(cd, sdd, xn, t), g.t("u_d_app"))
    close(O.warp_coordinate(cd, sdd, O.unnormalize_coord(cd, xn), t), g.t("u_d_warp"))
    vd, feat = g.t("u_viewdirs"), g.t("u_feat")
    close(O.render_head(cs, ss, xn, vd, feat, t[:, None]), g.t("u_s_head"))
    close(O.render_head(cd, sdd, xn, vd, feat, t[:, None]), g.t("u_d_head"))


def test_field_forward_and_compositing(golden_case):
    g = golden_case
    cs, cd = g.cfgs()
    rt = g.meta["ray_type"]
    rays, ts = g.t("rays"), g.t("ts")
    xyz, z, valid = g.t("xyz"), g.t("z_vals"), g.t("ray_valid")
    with torch.no_grad():
        fs = O.field_forward(cs, g.sd("static"), rays, ts, xyz, z, valid, rt)
        fd = O.field_forward(cd, g.sd("dynamic"), rays, ts, xyz, z, valid, rt)
    assert fs[0] is None and fs[1] is None and fs[2] is None and fs[5] is None
    for tag, f in (("s", fs), ("d", fd)):
        close(f[4], g.t(f"{tag}_weight"))
        close(f[6], g.t(f"{tag}_rgb"))
        close(f[7], g.t(f"{tag}_sigma"))
        close(f[9], g.t(f"{tag}_dists"))
    close(fd[2], g.t("d_blending"))
    close(fd[5], g.t("d_xyz_prime"))
    o13 = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, rt)
    assert len(o13) == 13
    for i, o in enumerate(o13):
        close(o, g.t(f"o13_{i}"), rtol=1e-5, atol=2e-6)


def test_gradients(golden_case):
    """Autograd through the oracle == autograd through the reference graph."""
    g = golden_case
    cs, cd = g.cfgs()
    rt = g.meta["ray_type"]
    s = int(g.raw["n_samples"])
    ss = {k: v.clone().requires_grad_(True) for k, v in g.sd("static").items()}
    sdd = {k: v.clone().requires_grad_(True) for k, v in g.sd("dynamic").items()}
    rays = g.t("rays").clone().requires_grad_(True)
    ts = g.t("ts")
    xyz, z, valid = O.sampleXYZ(cd, rays, s, rt, False)
    fs = O.field_forward(cs, ss, rays, ts, xyz, z, valid, rt)
    fd = O.field_forward(cd, sdd, rays, ts, xyz, z, valid, rt)
    o13 = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, rt)
    gen = torch.Generator().manual_seed(20211202 + 13)
    loss = 0.0
    for o in o13:
        loss = loss + (o * torch.randn(o.shape, generator=gen)).sum()
    loss = loss + (fd[5] * torch.randn(fd[5].shape, generator=gen)).sum() * 0.1
    loss.backward()
    close(loss.detach().reshape(-1), g.t("g_loss").reshape(-1), rtol=1e-5, atol=1e-5)
    close(rays.grad, g.t("g_rays"), rtol=1e-4, atol=1e-5)
    checked = 0
    for tag, sd in (("s", ss), ("d", sdd)):
        for n, p in sd.items():
            key = f"g_{tag}/{n}/head"
            if key not in g.raw:
                assert p.grad is None or float(p.grad.abs().sum()) == 0.0 or n.startswith("scene_flow"), n
                continue
            gfl = p.grad.reshape(-1)
            close(gfl[:256], g.t(key), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(float(gfl.double().abs().sum()), float(g.raw[f"g_{tag}/{n}/abssum"]), rtol=1e-4)
            checked += 1
    assert checked > 30
