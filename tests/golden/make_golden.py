#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden.py [/root/reference]

The reference checkout is imported read-only through the stub recipe of SURVEY.md
section 8c / appendix B: empty module stubs for the absent, arithmetic-irrelevant
packages (imageio, kornia, easydict, plyfile, skimage) and a CPU shim for
`Tensor.get_device` (the reference does `x.to(y.get_device())`, which raises on CPU).
Nothing from the reference is copied: only seeded inputs, the random parameters its
constructors produce, and the outputs it computes are stored.

The GPU box has no /root/reference; tests there read only the .npz files.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 20211202  # the reference's own seed, train.py:2663


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    sys.path.insert(0, REF)
    _stub("imageio")
    _stub("kornia", create_meshgrid=None)
    _stub("plyfile")

    class _ED(dict):
        __getattr__ = dict.__getitem__

    _stub("easydict", EasyDict=_ED)
    sk = _stub("skimage")
    sk.measure = _stub("skimage.measure")
    sk.morphology = _stub("skimage.morphology")
    torch.Tensor.get_device = lambda self: self.device
    with contextlib.redirect_stdout(io.StringIO()):
        import models.tensoRF as T  # noqa
        import models.tensorBase as B  # noqa
        import renderer as R  # noqa
    return T, B, R


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def np_(x):
    return x.detach().cpu().numpy()


def put_sd(dst, prefix, module):
    for k, v in module.state_dict().items():
        dst[f"{prefix}/{k}"] = np_(v)


COMMON = dict(density_n_comp=[16, 4, 4], appearance_n_comp=[48, 12, 12], app_dim=27, alphaMask_thres=1e-4,
              density_shift=-10, distance_scale=25, pos_pe=6, view_pe=0, featureC=128, step_ratio=2.0,
              fea2denseAct="relu")


def build_models(T, aabb, grid, near_far, static_mode):
    dev = torch.device("cpu")
    # train.py:874-893 (static, fea_pe=2) and train.py:903-922 (dynamic, fea_pe=0)
    static = quiet(T.TensorVMSplit, aabb, grid, 12, dev, near_far=near_far, shadingMode=static_mode, fea_pe=2, **COMMON)
    dyn = quiet(T.TensorVMSplit_TimeEmbedding, aabb, grid, 12, dev, near_far=near_far,
                shadingMode="MLP_Fea_late_view", fea_pe=0, **COMMON)
    # the reference initialises the last head bias to 0 and small planes; perturb a little so
    # every parameter matters in the vectors (still the reference code doing the arithmetic)
    with torch.no_grad():
        for m in (static, dyn):
            for n, p in m.named_parameters():
                if n.endswith("bias"):
                    p.add_(0.05 * torch.randn_like(p))
    return static, dyn


def run_case(T, B, R, name, aabb, grid, near_far, ray_type, static_mode, n_rays, n_samples, rays, ts):
    out = {}
    static, dyn = build_models(T, aabb, grid, near_far, static_mode)
    put_sd(out, "sd_static", static)
    put_sd(out, "sd_dynamic", dyn)
    out["aabb"] = np_(aabb)
    out["grid"] = np.array(grid)
    out["near_far"] = np.array(near_far, dtype=np.float32)
    out["rays"] = np_(rays)
    out["ts"] = np_(ts)
    out["n_samples"] = np.array(n_samples)
    meta = dict(ray_type=ray_type, static_mode=static_mode)
    out["meta"] = np.array(repr(meta))

    with torch.no_grad():
        xyz, z, valid = R.sampleXYZ(dyn, rays, N_samples=n_samples, ray_type=ray_type, is_train=False)
        out["xyz"], out["z_vals"], out["ray_valid"] = np_(xyz), np_(z), np_(valid)
        fs = static(rays, ts, None, xyz, z, valid, is_train=False, ray_type=ray_type, N_samples=n_samples)
        fd = dyn(rays, ts, None, xyz, z, valid, is_train=False, ray_type=ray_type, N_samples=n_samples)
        assert fs[0] is None and fs[1] is None and fs[2] is None and fs[5] is None
        for tag, f in (("s", fs), ("d", fd)):
            out[f"{tag}_weight"], out[f"{tag}_rgb"], out[f"{tag}_sigma"] = np_(f[4]), np_(f[6]), np_(f[7])
            out[f"{tag}_dists"] = np_(f[9])
        out["d_blending"], out["d_xyz_prime"] = np_(fd[2]), np_(fd[5])
        o13 = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, is_train=False, ray_type=ray_type)
        for i, o in enumerate(o13):
            out[f"o13_{i}"] = np_(o)

        # train-mode sampling: same seed -> same jitter (the reference draws rand_like inside)
        torch.manual_seed(SEED + 7)
        xyz_t, z_t, valid_t = R.sampleXYZ(dyn, rays, N_samples=n_samples, ray_type=ray_type, is_train=True)
        out["train_xyz"], out["train_z"], out["train_valid"] = np_(xyz_t), np_(z_t), np_(valid_t)

        # unit vectors: lookups with coordinates that leave [-1,1] (zeros padding), warp, heads
        g = torch.Generator().manual_seed(SEED + 11)
        P = 257
        xn = (torch.rand(P, 3, generator=g) * 2.5 - 1.25)
        tt = torch.rand(P, generator=g) * 2 - 1
        out["u_xn"], out["u_t"] = np_(xn), np_(tt)
        out["u_s_density"] = np_(static.compute_densityfeature(xn, tt, None))
        out["u_s_app"] = np_(static.compute_appfeature(xn, tt, None))
        out["u_d_density"] = np_(dyn.compute_densityfeature(xn, tt, None))
        out["u_d_blending"] = np_(dyn.compute_blendingfeature(xn, tt, None))
        out["u_d_app"] = np_(dyn.compute_appfeature(xn, tt, None))
        out["u_d_warp"] = np_(dyn.warp_coordinate(dyn.unnormalize_coord(xn), tt))
        vd = torch.nn.functional.normalize(torch.randn(P, 3, generator=g), dim=-1)
        feat = torch.randn(P, 27, generator=g)
        out["u_viewdirs"], out["u_feat"] = np_(vd), np_(feat)
        out["u_s_head"] = np_(static.renderModule(xn, vd, feat, tt[:, None]))
        out["u_d_head"] = np_(dyn.renderModule(xn, vd, feat, tt[:, None]))

    # gradients of a fixed scalar of the 13 outputs, autograd through the reference graph
    rays_g = rays.clone().requires_grad_(True)
    xyz, z, valid = R.sampleXYZ(dyn, rays_g, N_samples=n_samples, ray_type=ray_type, is_train=False)
    fs = static(rays_g, ts, None, xyz, z, valid, is_train=False, ray_type=ray_type, N_samples=n_samples)
    fd = dyn(rays_g, ts, None, xyz, z, valid, is_train=False, ray_type=ray_type, N_samples=n_samples)
    o13 = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays_g, is_train=False, ray_type=ray_type)
    g = torch.Generator().manual_seed(SEED + 13)
    loss = 0.0
    for o in o13:
        loss = loss + (o * torch.randn(o.shape, generator=g)).sum()
    loss = loss + (fd[5] * torch.randn(fd[5].shape, generator=g)).sum() * 0.1
    loss.backward()
    out["g_loss"] = np_(loss)
    out["g_rays"] = np_(rays_g.grad)
    for tag, m in (("s", static), ("d", dyn)):
        for n, p in m.named_parameters():
            if p.grad is None:
                continue
            gfl = p.grad.reshape(-1)
            out[f"g_{tag}/{n}/sum"] = np_(gfl.double().sum())
            out[f"g_{tag}/{n}/abssum"] = np_(gfl.double().abs().sum())
            out[f"g_{tag}/{n}/head"] = np_(gfl[:256])
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(name, "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "valid frac", float(valid.float().mean()),
          "app frac s/d", float((fs[4] > 1e-4).float().mean()), float((fd[4] > 1e-4).float().mean()))


def main():
    T, B, R = import_reference()
    torch.manual_seed(SEED)
    g = torch.Generator().manual_seed(SEED + 1)

    # case A: NDC (dataLoader/nvidia.py:246-251 constants), odd/even grid for the stride-2/4 views
    aabb = torch.tensor([[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]])
    n, s = 48, 40
    o = torch.cat((torch.rand(n, 2, generator=g) * 2 - 1, -torch.ones(n, 1)), -1)
    o[:6, 0] *= 1.7  # a few rays start/leave outside the aabb -> ray_valid False samples
    d = torch.cat((0.25 * torch.randn(n, 2, generator=g), 2 * torch.ones(n, 1)), -1)
    d[6:10, 2] = 2.3  # overshoot far plane -> invalid tail
    rays = torch.cat((o, d), -1)
    ts = torch.rand(n, generator=g) * 2 - 1
    run_case(T, B, R, "ndc_small", aabb, [21, 24, 18], [0.0, 1.0], "ndc", "MLP_Fea", n, s, rays, ts)

    # case B: contracted rays (DAVIS constants), static head MLP_Fea_TimeEmbedding (opt.py:143-148 default)
    aabb = torch.tensor([[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]])
    n, s = 40, 33
    o = 0.05 * torch.randn(n, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    rays = torch.cat((o, d), -1)
    ts = torch.rand(n, generator=g) * 2 - 1
    run_case(T, B, R, "contract_small", aabb, [16, 19, 20], [0.0, 256.0], "contract", "MLP_Fea_TimeEmbedding",
             n, s, rays, ts)

    # standalone known-answer vectors for the two pure functions
    g = torch.Generator().manual_seed(SEED + 3)
    x = torch.randn(33, 3, generator=g)
    sig = torch.relu(torch.randn(17, 29, generator=g)) * 3
    dist = torch.rand(17, 29, generator=g) * 0.3
    a, w, tl = B.raw2alpha(sig, dist)
    np.savez_compressed(os.path.join(OUT, "pure_fns.npz"), pe_x=np_(x), pe_6=np_(B.positional_encoding(x, 6)),
                        pe_10=np_(B.positional_encoding(x, 10)), r2a_sigma=np_(sig), r2a_dist=np_(dist),
                        r2a_alpha=np_(a), r2a_weight=np_(w), r2a_last=np_(tl))
    print("pure_fns done")


if __name__ == "__main__":
    main()
