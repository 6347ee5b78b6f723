"""Autograd glue between the drop-in modules (models/, renderer.py) and the C-ABI kernels.

Forward always runs the native kernels.  When a result must carry gradients the matching
torch.autograd.Function (backed by the *_bwd entry points) is used; if the backward kernels for
a requested gradient are not available the call fails loudly -- nothing silently detaches.
"""
from __future__ import annotations

import torch

from . import lib as L
from . import ops


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def _model_needs_grad(m) -> bool:
    return torch.is_grad_enabled() and any(p.requires_grad for p in m.parameters())


# ----------------------------------------------------------------------------- a1-a3


def sample_points(model, rays, n_samples, ray_type, is_train):
    """-> xyz [N,S,3], zvec [S], valid [N,S]."""
    ops._require_cuda(rays)
    cfg = model.native_cfg(ray_type)
    zvec = ops.make_zvec(cfg, n_samples, ray_type, is_train, rays.device)
    if _needs_grad(rays):
        from . import backward as B

        xyz, z, valid = B.SamplePointsFn.apply(rays, zvec, model, ray_type)
    else:
        xyz, z, valid = ops.sample_points(cfg, rays.detach(), zvec)
    return xyz, zvec, valid


# ----------------------------------------------------------------------------- a10


def field_forward(model, rays, ts, xyz, z_vals, valid, ray_type):
    ops._require_cuda(rays, xyz, z_vals, valid)
    if ray_type not in ("ndc", "contract"):
        raise NotImplementedError("ray_type must be 'ndc' or 'contract'")
    grad = _model_needs_grad(model) or _needs_grad(rays, xyz)
    if grad:
        from . import backward as B

        return B.field_forward_with_grad(model, rays, ts, xyz, z_vals, valid, ray_type)
    if model.is_dynamic:
        r = ops.dynamic_forward_raw(model, rays, ts, xyz, z_vals, valid, ray_type)
        return None, None, r["blend"], xyz, r["weight"], r["xyz_prime"], r["rgb"], r["sigma"], z_vals, r["dists"]
    r = ops.static_forward_raw(model, rays, xyz, z_vals, valid, ray_type)
    ops.static_app_raw(model, r, ray_type)
    return None, None, None, xyz, r["weight"], None, r["rgb"], r["sigma"], z_vals, r["dists"]


def warp_coordinate(model, xyz, t):
    """TensorVMSplit_TimeEmbedding.warp_coordinate for xyz [N,S,3] with per-ray time t[:,0]."""
    import ctypes as C

    ops._require_cuda(xyz)
    if _model_needs_grad(model) or _needs_grad(xyz):
        raise NotImplementedError("standalone warp_coordinate is forward-only; gradients flow through TensorBase.forward")
    lib = L.load()
    xyz = ops._f32c(xyz)
    n, s = xyz.shape[:2]
    ts = ops._f32c(t[:, 0] if t.dim() == 2 else t)
    ly = model.packed_layers()
    cfg = model.native_cfg("ndc")
    tfeat = torch.empty(n, L.TFEAT, device=xyz.device)
    L.check(lib.dynrf_time_embed_fwd(ops._p(ts), n, C.byref(ly["layer1"]), C.byref(ly["layer2"]), ops._p(tfeat), ops._stream()),
            "dynrf_time_embed_fwd")
    out = torch.empty_like(xyz)
    L.check(lib.dynrf_warp_fwd(C.byref(cfg), ops._p(xyz), n, s, ops._p(tfeat), C.byref(ly["layer3"]), C.byref(ly["layer4"]),
                               C.byref(ly["layer5"]), ops._p(out), ops._stream()), "dynrf_warp_fwd")
    return out


# ----------------------------------------------------------------------------- a19


def _split_ray_out(o):
    # out_ray [N,16]: rgb_full3 depth_full acc_full rgb_s3 depth_s acc_s rgb_d3 depth_d acc_d dynamicness
    return (o[:, 0:3], o[:, 3], o[:, 4], o[:, 5:8], o[:, 8], o[:, 9], o[:, 10:13], o[:, 13], o[:, 14], o[:, 15])


def raw2outputs(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk, is_train=False, ray_type="ndc"):
    """renderer.raw2outputs (renderer.py:173-315) -> the reference's 13-tuple."""
    # the reference draws a host-side coin for the white background when training (renderer.py:269)
    white = bool(is_train and (torch.rand((1,)) < 0.5))
    if _needs_grad(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk):
        from . import backward as B

        o, w_full, w_s, w_d = B.CompositeFn.apply(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk, white,
                                                  ray_type)
    else:
        o, w_full, w_s, w_d, _ = ops.composite_raw(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk, white,
                                                   ray_type)
    rf, df, af, rs, ds, as_, rd, dd, ad, dyn = _split_ray_out(o)
    return rf, df, af, w_full, rs, ds, as_, w_s, rd, dd, ad, w_d, dyn
