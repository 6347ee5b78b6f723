"""Torch-facing wrappers + autograd for the rows next to the ray-march path:

  sample_ray          a4  models/tensorBase.py:501-522
  generate_rays       f1  train.py:96-103, 1062-1077; camera.py:8-15; dataLoader/ray_utils.py:53-90, 115-140
  weighted_point      f2  the S-reduction of renderer.render_3d_point (renderer.py:1335-1338)
  tv_loss_vm          f3  utils.TVLoss over every plane / line of a VM set, one launch (utils.py:157-181, models/tensoRF.py:100-116)
  dense_l1            f3  density_L1 / blending_L1 without the dense volume (models/tensoRF.py:80-98, 378-416)

CUDA only, like the rest of the package: CPU tensors raise (no fallback)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import lib as L
from . import ops
from .ops import _f32c, _p, _require_cuda, _stream


# ----------------------------------------------------------------------------- a4


def sample_ray(model, rays_o, rays_d, is_train: bool, n_samples: int):
    """-> xyz [N,S,3], z_vals [N,S] (per ray), valid [N,S] bool."""
    _require_cuda(rays_o, rays_d)
    rays = _f32c(torch.cat((rays_o, rays_d), -1))
    n = rays.shape[0]
    cfg = model.native_cfg("ndc")        # aabb / near / far only
    jitter = torch.rand(n, device=rays.device) if is_train else None       # one draw per ray (rng[:, [0]])
    xyz = torch.empty(n, n_samples, 3, device=rays.device)
    z = torch.empty(n, n_samples, device=rays.device)
    valid = torch.empty(n, n_samples, device=rays.device, dtype=torch.bool)
    L.check(L.load().dynrf_sample_ray(C.byref(cfg), _p(rays), n, n_samples, float(model.stepSize), _p(jitter), _p(xyz), _p(z), _p(valid),
                                      _stream()), "dynrf_sample_ray")
    return xyz, z, valid


# ----------------------------------------------------------------------------- f1


def pose_to_mtx(pose: torch.Tensor) -> torch.Tensor:
    """camera.pose_to_mtx (camera.py:8-15) on the [F,9] pose table -- F is the number of frames (tens to hundreds), so the
    Gram-Schmidt stays a handful of tiny torch ops whose autograd closes the chain from the kernel's gmtx back to the 9-D poses."""
    b1 = pose[..., 0:3]
    b1 = b1 / torch.norm(b1, dim=-1)[:, None]
    b2 = pose[..., 3:6] - torch.sum(b1 * pose[..., 3:6], dim=-1)[:, None] * b1
    b2 = b2 / torch.norm(b2, dim=-1)[:, None]
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3, pose[..., 6:9]], dim=-1)


class _RayGenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, mtx, focal, H, W, ndc, near, uv=None, view=None):
        mtxc, fc = _f32c(mtx.detach()), _f32c(focal.detach()).reshape(1)
        if uv is not None:
            uv, view = _f32c(uv.detach()), view.long().contiguous()
            ids, b = None, uv.shape[0]
        else:
            ids = ids.contiguous()
            b = ids.shape[0]
        rays = torch.empty(b, 6, device=mtxc.device)
        L.check(L.load().dynrf_raygen_fwd(_p(ids), _p(uv), _p(view), b, _p(mtxc), mtxc.shape[0], _p(fc), H, W, int(ndc), float(near), _p(rays),
                                          _stream()), "dynrf_raygen_fwd")
        ctx.saved = (ids, uv, view, mtxc, fc)
        ctx.args = (b, H, W, int(ndc), float(near), focal.shape)
        return rays

    @staticmethod
    def backward(ctx, grays):
        ids, uv, view, mtx, fc = ctx.saved
        b, H, W, ndc, near, fshape = ctx.args
        gm = torch.zeros_like(mtx)
        gf = torch.zeros(1, device=mtx.device)
        L.check(L.load().dynrf_raygen_bwd(_p(ids), _p(uv), _p(view), b, _p(mtx), mtx.shape[0], _p(fc), H, W, ndc, near, _p(_f32c(grays)),
                                          _p(gm), _p(gf), _stream()), "dynrf_raygen_bwd")
        return None, gm, gf.reshape(fshape), None, None, None, None, None, None


def generate_rays(ray_idx: Optional[torch.Tensor], poses9: torch.Tensor, focal: torch.Tensor, H: int, W: int, ray_type: str = "ndc",
                  near: float = 1.0, uv: Optional[torch.Tensor] = None, view_ids: Optional[torch.Tensor] = None,
                  poses_mtx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ids2pixel -> get_ray_directions_lean -> get_rays_lean -> (ndc_rays_blender2) in one kernel: flat pixel ids [B] (int64),
    9-D poses [F,9], focal (0-d / 1-element tensor) -> rays [B,6], differentiable w.r.t. poses and focal.
    uv [B,2] (sub-pixel x, y) + view_ids [B] replace the ids for the flow-displaced rays (train.py:1433-1455); `poses_mtx`
    [F,3,4] may be passed instead of the 9-D table (e.g. the shifted neighbour-frame poses, train.py:1081-1085)."""
    _require_cuda(poses9 if poses_mtx is None else poses_mtx, focal)
    if ray_idx is not None and ray_idx.dtype != torch.int64:
        ray_idx = ray_idx.long()
    mtx = pose_to_mtx(poses9) if poses_mtx is None else poses_mtx
    return _RayGenFn.apply(ray_idx, mtx, focal, int(H), int(W), ray_type == "ndc", near, uv, view_ids)


# ----------------------------------------------------------------------------- f2


class _WeightedPointFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, pts, sf, clamp):
        wc, pc = _f32c(w.detach()), _f32c(pts.detach())
        sc = None if sf is None else _f32c(sf.detach())
        n, s = wc.shape
        pm = torch.empty(n, 3, device=wc.device)
        acc = torch.empty(n, device=wc.device)
        lo, hi = (clamp if clamp is not None else (0.0, 0.0))
        L.check(L.load().dynrf_weighted_point_fwd(_p(wc), _p(pc), _p(sc), int(clamp is not None), float(lo), float(hi), n, s, _p(pm), _p(acc),
                                                  _stream()), "dynrf_weighted_point_fwd")
        ctx.save_for_backward(wc, pc, sc)
        ctx.clamp = clamp
        ctx.set_materialize_grads(False)
        return pm, acc

    @staticmethod
    def backward(ctx, gpm, gacc):
        wc, pc, sc = ctx.saved_tensors
        n, s = wc.shape
        clamp = ctx.clamp
        lo, hi = (clamp if clamp is not None else (0.0, 0.0))
        need_w = ctx.needs_input_grad[0]
        need_p = ctx.needs_input_grad[1] or (sc is not None and ctx.needs_input_grad[2])
        if gpm is None:
            gpm = torch.zeros(n, 3, device=wc.device)
        gw = torch.empty(n, s, device=wc.device) if need_w else None
        gm = torch.empty(n, s, 3, device=wc.device) if need_p else None
        if need_w or need_p:
            L.check(L.load().dynrf_weighted_point_bwd(_p(wc), _p(pc), _p(sc), int(clamp is not None), float(lo), float(hi), n, s,
                                                      _p(_f32c(gpm)), _p(None if gacc is None else _f32c(gacc)), _p(gw), _p(gm), _stream()),
                    "dynrf_weighted_point_bwd")
        return gw, (gm if ctx.needs_input_grad[1] else None), (gm if (sc is not None and ctx.needs_input_grad[2]) else None), None


def weighted_point(weights, pts, scene_flow=None, clamp=None):
    """-> (sum_s w * moved [N,3], sum_s w [N]) with moved = pts (+ scene_flow) (clamped to `clamp` = (lo, hi) if given)."""
    _require_cuda(weights, pts)
    return _WeightedPointFn.apply(weights, pts, scene_flow, clamp)


# ----------------------------------------------------------------------------- f3: TV


def _is_channel_last(p):
    _, c, h, w = p.shape
    s = p.stride()
    return (c == 1 or s[1] == 1) and (h == 1 or s[2] == w * c) and (w == 1 or s[3] == c)


class _TvFn(torch.autograd.Function):
    """total = sum_i scale_i * TVLoss_weight * 2 * (h_tv_i / count_h_i + w_tv_i / count_w_i) / batch  over the given tensors,
    value and gradient in one pass.  A (1,C,L,1) line has count_w = 0: its value term is 0/0 = NaN exactly as utils.py:172-174
    computes it, while its gradient (only the h direction has terms) stays finite -- also as in the reference."""

    @staticmethod
    def forward(ctx, weight, scales, *tensors):
        lib = L.load()
        n = len(tensors)
        dev = tensors[0].device
        any_grad = any(ctx.needs_input_grad[2:])     # (grad mode is off inside Function.forward: ask the ctx, not the tensors)
        xs, gs, Cs, Hs, Ws, ch, cw, counts = [], [], [], [], [], [], [], []
        for t, sc in zip(tensors, scales):
            x = t.detach()
            hfast = (not _is_channel_last(x)) and ops.is_hfast(x)     # the z planes are stored [W][H][C]
            if not hfast and not _is_channel_last(x):
                x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
            b, c, h, w = x.shape
            if b != 1:
                raise NotImplementedError("fused TV expects (1,C,H,W) planes / (1,C,L,1) lines")
            count_h, count_w = c * (h - 1) * w, c * h * (w - 1)
            xs.append(x)
            gs.append(torch.empty_like(x, memory_format=torch.preserve_format) if any_grad else None)
            kh = sc * weight * 2.0 / count_h / b if count_h > 0 else 0.0
            kw = sc * weight * 2.0 / count_w / b if count_w > 0 else 0.0
            Cs.append(c)
            if hfast:      # the kernel sees the physical [rows = W][cols = H] tensor: its "h" direction is the logical w direction
                Hs.append(w), Ws.append(h), ch.append(kw), cw.append(kh)
            else:
                Hs.append(h), Ws.append(w), ch.append(kh), cw.append(kw)
            counts.append((count_h, count_w, b, sc, hfast))
        sums = torch.zeros(2 * n, device=dev, dtype=torch.float64)
        xa = (C.c_void_p * n)(*[x.data_ptr() for x in xs])
        ga = (C.c_void_p * n)(*[(g.data_ptr() if g is not None else None) for g in gs])
        ia = lambda v: (C.c_int32 * n)(*v)  # noqa: E731
        fa = lambda v: (C.c_float * n)(*v)  # noqa: E731
        L.check(lib.dynrf_tv_fused(n, xa, ga if any_grad else None, ia(Cs), ia(Hs), ia(Ws), fa(ch), fa(cw), _p(sums), _stream()),
                "dynrf_tv_fused")
        s32 = sums.float()
        total = torch.zeros((), device=dev)
        for i, (count_h, count_w, b, sc, hfast) in enumerate(counts):
            h_tv, w_tv = (s32[2 * i + 1], s32[2 * i]) if hfast else (s32[2 * i], s32[2 * i + 1])
            # same expression as TVLoss.forward: weight * 2 * (h_tv / count_h + w_tv / count_w) / batch  (0/0 -> NaN for lines)
            total = total + (weight * 2 * (h_tv / count_h + w_tv / count_w) / b) * sc
        ctx.gs = gs
        return total

    @staticmethod
    def backward(ctx, gout):
        return (None, None, *[None if (g is None or not need) else g * gout for g, need in zip(ctx.gs, ctx.needs_input_grad[2:])])


def tv_loss_vm(weight: float, planes: Sequence[torch.Tensor], lines: Sequence[torch.Tensor]) -> torch.Tensor:
    """sum_i TV(plane_i) * 1e-2 + TV(line_i) * 1e-3  (models/tensoRF.py:100-116, 419-444), one kernel launch."""
    ts = list(planes) + list(lines)
    _require_cuda(*ts)
    return _TvFn.apply(float(weight), [1e-2] * len(planes) + [1e-3] * len(lines), *ts)


# ----------------------------------------------------------------------------- f3: dense L1


class _DenseL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, shift, *tensors):
        planes, lines = tensors[:3], tensors[3:]
        vm = ops.vm_struct(planes, lines)
        dev = planes[0].device
        any_grad = any(ctx.needs_input_grad[2:])
        total = torch.zeros(1, device=dev, dtype=torch.float64)
        gp = gl = gstruct = None
        if any_grad:
            gp = [ops.zeros_like_cl(p) for p in planes]
            gl = [ops.zeros_like_cl(p) for p in lines]
            gstruct = ops.vm_grad_struct(gp, gl)
        L.check(L.load().dynrf_dense_l1(C.byref(vm), {"relu": L.ACT_RELU, "softplus": L.ACT_SOFTPLUS}[act], float(shift), _p(total),
                                        C.byref(gstruct) if gstruct is not None else None, _stream()), "dynrf_dense_l1")
        numel = planes[0].shape[3] * planes[0].shape[2] * lines[0].shape[2]
        ctx.grads = None if not any_grad else [g / numel for g in gp + gl]
        return (total[0] / numel).float()

    @staticmethod
    def backward(ctx, gout):
        return (None, None, *[(g * gout if need else None) for g, need in zip(ctx.grads, ctx.needs_input_grad[2:])])


def dense_l1(model, planes, lines) -> torch.Tensor:
    """mean |feature2density(sum_c plane (x) line)| over the dense grid, streamed (no [C,Gy,Gx,Gz] product)."""
    _require_cuda(*planes, *lines)
    return _DenseL1Fn.apply(model.fea2denseAct, model.density_shift, *planes, *lines)
