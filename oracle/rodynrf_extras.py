"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the rows next to the ray-march path.

Plain PyTorch restatements, each citing the reference lines it follows (paths relative to the reference checkout):
  a4  TensorBase.sample_ray                       models/tensorBase.py:501-522
  f1  ray generation (ids -> pixels -> camera dirs -> world rays -> NDC)   train.py:96-103, camera.py:8-15,
      dataLoader/ray_utils.py:53-90, 115-140
  f2  scene-flow MLP + induce_flow / render_3d_point chain   models/tensoRF.py:446-519, renderer.py:1266-1392
  f3  TVLoss, TV_loss_*, density_L1 / blending_L1, vector_comp_diffs   utils.py:157-181, models/tensoRF.py:63-116, 378-444

Pinned by tests/golden/utils_small.npz (tests/golden/make_golden_utils.py runs the UNMODIFIED reference and stores its
outputs and autograd gradients); tests/test_oracle_utils_golden.py checks every function here against those vectors.
Nothing in the product package imports this file.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

from .rodynrf_oracle import FieldCfg, _linear, _plist, feature2density, normalize_coord, positional_encoding

Tensor = torch.Tensor


# --------------------------------------------------------------------------- a4


def sample_ray(cfg: FieldCfg, step_size: float, rays_o: Tensor, rays_d: Tensor, n_samples: int, is_train: bool = False,
               jitter: Tensor | None = None) -> Tuple[Tensor, Tensor, Tensor]:
    """models/tensorBase.py:501-522 -- aabb entry distance + fixed step marching; `jitter` ([N,1] uniforms) replaces the
    reference's rand_like draw when given."""
    near, far = cfg.near_far
    vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
    rate_a = (cfg.aabb[1] - rays_o) / vec
    rate_b = (cfg.aabb[0] - rays_o) / vec
    t_min = torch.minimum(rate_a, rate_b).amax(-1).clamp(min=near, max=far)
    rng = torch.arange(n_samples)[None].float()
    if is_train:
        rng = rng.repeat(rays_d.shape[-2], 1)
        rng = rng + (torch.rand_like(rng[:, [0]]) if jitter is None else jitter)
    interpx = t_min[..., None] + step_size * rng.to(rays_o.device)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * interpx[..., None]
    outside = ((cfg.aabb[0] > pts) | (pts > cfg.aabb[1])).any(dim=-1)
    return pts, interpx, ~outside


# --------------------------------------------------------------------------- f1


def pose_to_mtx(pose: Tensor) -> Tensor:
    """camera.py:8-15 -- 9-D pose (two 3-vectors Gram-Schmidt'ed into a rotation + a translation) -> [F,3,4]."""
    b1 = pose[..., 0:3]
    b1 = b1 / torch.norm(b1, dim=-1)[:, None]
    b2 = pose[..., 3:6] - torch.sum(b1 * pose[..., 3:6], dim=-1)[:, None] * b1
    b2 = b2 / torch.norm(b2, dim=-1)[:, None]
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3, pose[..., 6:9]], dim=-1)


def ids2pixel(W: int, H: int, ids: Tensor):
    """train.py:96-103."""
    return ids % W, (ids // W) % H, ids // (W * H)


def generate_rays(ids: Tensor, pose9: Tensor, focal: Tensor, H: int, W: int, ray_type: str = "ndc", near: float = 1.0) -> Tensor:
    """train.py:1062-1077: ids2pixel -> get_ray_directions_lean (ray_utils.py:53-69) -> get_rays_lean (:72-90) ->
    ndc_rays_blender2 (:115-140) for NDC scenes.  -> rays [B,6]."""
    col, row, view = ids2pixel(W, H, ids)
    i, j = col.float() + 0.5, row.float() + 0.5
    dirs = torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)
    c2w = pose_to_mtx(pose9)[view]
    rays_d = torch.bmm(dirs[:, None, :], torch.transpose(c2w[:, :3, :3], 1, 2))[:, 0, :]
    rays_o = c2w[:, :3, 3]
    if ray_type == "ndc":
        t = -(near + rays_o[..., 2]) / rays_d[..., 2]
        rays_o = rays_o + t[..., None] * rays_d
        o0 = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
        o1 = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
        o2 = 1.0 + 2.0 * near / rays_o[..., 2]
        d0 = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
        d1 = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
        d2 = -2.0 * near / rays_o[..., 2]
        rays_o, rays_d = torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)
    return torch.cat([rays_o, rays_d], -1)


# --------------------------------------------------------------------------- f2


def scene_flow(cfg: FieldCfg, sd: Dict[str, Tensor], pts: Tensor, t: Tensor) -> Tuple[Tensor, Tensor]:
    """TensorVMSplit_TimeEmbedding.get_forward_backward_scene_flow, models/tensoRF.py:446-464.  pts [N,S,3], t [N]."""
    n, s = pts.shape[:2]
    flat = normalize_coord(cfg, pts.reshape(-1, 3))
    tt = t[:, None].repeat(1, s).view(-1, 1)
    x = torch.cat((flat, positional_encoding(flat, 4), tt, positional_encoding(tt, 4)), -1)
    h = x
    for i in (0, 2, 4):
        h = F.relu(_linear(sd, f"scene_flow_mlp.{i}", h))
    out = _linear(sd, "scene_flow_mlp.6", h).view(n, s, 6)
    return out[..., 0:3], out[..., 3:6]


def scene_flow_point_single(cfg: FieldCfg, sd, pts_map: Tensor, t: Tensor):
    """models/tensoRF.py:505-519."""
    xn = normalize_coord(cfg, pts_map)
    x = torch.cat((xn, positional_encoding(xn, 4), t[..., None], positional_encoding(t[..., None], 4)), -1)
    h = x
    for i in (0, 2, 4):
        h = F.relu(_linear(sd, f"scene_flow_mlp.{i}", h))
    out = _linear(sd, "scene_flow_mlp.6", h)
    f, b = out[..., 0:3], out[..., 3:6]
    return pts_map + f, pts_map + b, f, b


def ndc2world(pts: Tensor, H, W, f) -> Tensor:
    """renderer.py:1266-1274."""
    pz = 2 / (torch.clamp(pts[..., 2:], min=-1.0, max=1 - 1e-6) - 1)
    px = -pts[..., 0:1] * pz * W / 2 / f
    py = -pts[..., 1:2] * pz * H / 2 / f
    return torch.cat([px, py, pz], -1)


def world2ndc(p: Tensor, H, W, f) -> Tensor:
    """renderer.py:1277-1283."""
    o0 = -1.0 / (W / (2.0 * f)) * p[..., 0:1] / p[..., 2:]
    o1 = -1.0 / (H / (2.0 * f)) * p[..., 1:2] / p[..., 2:]
    o2 = 1.0 + 2.0 * 1 / p[..., 2:]
    return torch.cat([o0, o1, o2], -1)


def contract2world(p: Tensor) -> Tensor:
    """renderer.py:1286-1298 (functional form of the in-place masked update)."""
    nrm = p.abs().amax(-1, keepdim=True)
    return torch.where(nrm > 1.0, p / nrm * (-1 / (nrm - 2)), p)


def weighted_point(weights: Tensor, pts: Tensor, rays: Tensor, ray_type: str) -> Tensor:
    """First half of render_3d_point, renderer.py:1335-1351: expected 3-D point + background term."""
    acc = weights.sum(-1)[:, None]
    pm = (weights[..., None] * pts).sum(-2)
    if ray_type == "ndc":
        return pm + (1.0 - acc) * (rays[:, :3] + rays[:, 3:])
    far_pts = rays[:, :3] + rays[:, 3:] * 256.0
    nrm = far_pts.abs().amax(-1, keepdim=True)
    far_pts = torch.where(nrm > 1.0, (2 - 1 / nrm) * (far_pts / nrm), far_pts)
    return pm + (1.0 - acc) * far_pts


def project_point(H, W, f, c2w: Tensor, pts_map: Tensor, ray_type: str) -> Tuple[Tensor, Tensor]:
    """Second half of render_3d_point, renderer.py:1353-1376: NDC/contract -> world -> neighbour camera -> pixel + NDC depth."""
    w2c = c2w[:, :3, :3].transpose(1, 2)
    world = ndc2world(pts_map, H, W, f) if ray_type == "ndc" else contract2world(pts_map)
    world = world - c2w[..., 3]
    cam = torch.sum(world[..., None, :] * w2c[:, :3, :3], -1)
    plane = torch.cat([cam[..., 0:1] / (-cam[..., 2:]) * f + W * 0.5, -cam[..., 1:2] / (-cam[..., 2:]) * f + H * 0.5], -1)
    return plane, world2ndc(cam, H, W, f)[:, 2:]


def induce_flow(H, W, f, c2w: Tensor, weights: Tensor, pts: Tensor, pts_2d: Tensor, rays: Tensor, ray_type: str = "ndc"):
    """renderer.py:1380-1392 -> (induced_flow [N,2], induced_disp [N,1])."""
    plane, disp = project_point(H, W, f, c2w, weighted_point(weights, pts, rays, ray_type), ray_type)
    return plane - pts_2d, disp


# --------------------------------------------------------------------------- f3


def tv_loss(x: Tensor, ignore_axis: str | None = None) -> Tensor:
    """utils.TVLoss (weight 1), utils.py:157-181.  x: (1,C,H,W)."""
    b, c, h, w = x.shape
    count_h, count_w = c * (h - 1) * w, c * h * (w - 1)
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, : h - 1, :], 2).sum()
    w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, : w - 1], 2).sum()
    if ignore_axis is None:
        return 2 * (h_tv / count_h + w_tv / count_w) / b
    if ignore_axis == "h":
        return 2 * (w_tv / count_w) / b
    return 2 * (h_tv / count_h) / b


def tv_loss_vm(sd: Dict[str, Tensor], name: str) -> Tensor:
    """TV_loss_density / TV_loss_app / TV_loss_blending, models/tensoRF.py:100-116, 419-444 (lines have W = 1: 0/0 -> NaN
    exactly as the reference's own TODO at utils.py:174 notes; the train loop adds the result regardless)."""
    total = 0
    for p, l in zip(_plist(sd, f"{name}_plane"), _plist(sd, f"{name}_line")):
        total = total + tv_loss(p) * 1e-2 + tv_loss(l) * 1e-3
    return total


def dense_l1(cfg: FieldCfg, sd: Dict[str, Tensor], name: str = "density") -> Tensor:
    """density_L1 / blending_L1, models/tensoRF.py:80-98, 378-416: mean |act(sum_c plane (x) line)| over the dense grid."""
    p0, p1, p2 = (p[0] for p in _plist(sd, f"{name}_plane"))          # (C,Gy,Gx) (C,Gz,Gx) (C,Gz,Gy)
    l0, l1, l2 = (l[0, :, :, 0] for l in _plist(sd, f"{name}_line"))   # (C,Gz) (C,Gy) (C,Gx)
    vol = (torch.einsum("cyx,cz->yxz", p0, l0) + torch.einsum("czx,cy->yxz", p1, l1) + torch.einsum("czy,cx->yxz", p2, l2))
    return feature2density(cfg, vol).abs().mean()


def vector_comp_diffs(sd: Dict[str, Tensor]) -> Tensor:
    """models/tensoRF.py:63-78 (static class only)."""
    total = 0
    for name in ("density_line", "app_line"):
        for ln in _plist(sd, name):
            v = ln[0, :, :, 0]
            dotp = v @ v.t()
            c = dotp.shape[0]
            total = total + dotp.view(-1)[1:].view(c - 1, c + 1)[..., :-1].abs().mean()
    return total


def upsample_vm(sd: Dict[str, Tensor], names: Sequence[str], res: Sequence[int]) -> Dict[str, Tensor]:
    """up_sampling_VM, models/tensoRF.py:198-220 -> new state-dict entries."""
    from .rodynrf_oracle import MAT_MODE, VEC_MODE

    out = dict(sd)
    for name in names:
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            out[f"{name}_plane.{i}"] = F.interpolate(sd[f"{name}_plane.{i}"], size=(res[m1], res[m0]), mode="bilinear", align_corners=True)
            out[f"{name}_line.{i}"] = F.interpolate(sd[f"{name}_line.{i}"], size=(res[VEC_MODE[i]], 1), mode="bilinear", align_corners=True)
    return out
