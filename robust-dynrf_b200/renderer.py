"""Drop-in for the hot-path functions of the reference's renderer.py:

  sampleXYZ                   renderer.py:147-170
  raw2outputs                 renderer.py:173-315
  OctreeRender_trilinear_fast renderer.py:24-144   (dead code upstream: raises TypeError as shipped)
  render_rays                 the fused evaluation pass train.py / evaluation_path perform per chunk
                              (renderer.py:1061-1130), one call

Same names, argument order and return tuples.  The image/video loops of the reference's
renderer.py (render / evaluation / evaluation_path: host orchestration + file IO) are out of
scope; INTEGRATION.md shows how an unmodified checkout picks these three functions up.
"""
import torch

from dynrf_b200 import autograd as _ag
from dynrf_b200 import ops as _ops


def sampleXYZ(tensorf, rays_train, N_samples, ray_type="ndc", is_train=False):
    """-> xyz_sampled [N,S,3], z_vals [N,S] (the batch-wide depth vector tiled per ray), ray_valid [N,S] bool."""
    if ray_type not in ("ndc", "contract"):
        raise NotImplementedError("ray_type must be 'ndc' or 'contract' (sample_ray is unused by every config)")
    n_samples = N_samples if N_samples > 0 else tensorf.nSamples
    xyz, zvec, valid = _ag.sample_points(tensorf, rays_train[:, :6], n_samples, ray_type, is_train)
    return xyz, zvec[None].expand(xyz.shape[0], -1).contiguous(), valid


def raw2outputs(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk, is_train=False, ray_type="ndc"):
    """13-tuple: rgb_map_full, depth_map_full, acc_map_full, weights_full, rgb_map_s, depth_map_s, acc_map_s,
    weights_s, rgb_map_d, depth_map_d, acc_map_d, weights_d, dynamicness_map."""
    return _ag.raw2outputs(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays_chunk, is_train, ray_type)


def OctreeRender_trilinear_fast(rays, ts, timeembeddings, tensorf, xyz_sampled, z_vals_input, ray_valid, chunk=4096,
                                N_samples=-1, ray_type="ndc", white_bg=True, is_train=False, device="cuda"):
    """Chunked field forward with the reference's signature and 11-tuple layout.

    Upstream this function is never called and raises TypeError for both model classes
    (`None - Tensor` at renderer.py:94, `torch.cat([None])` at :133), so there is no reference
    behaviour to match ("parity unpinned").  Defined behaviour here: None-safe concatenation;
    rgb/depth maps (which TensorBase.forward returns as None) stay None."""
    cols = [[] for _ in range(8)]  # blending, pts_ref, weights, delta_xyz, rgb_point, sigma, z_val, dist
    n = rays.shape[0]
    for c0 in range(0, n, chunk):
        sl = slice(c0, c0 + chunk)
        out = tensorf(rays[sl].to(device), ts[sl].to(device), None if timeembeddings is None else timeembeddings[sl].to(device),
                      xyz_sampled[sl].to(device), z_vals_input[sl].to(device), ray_valid[sl].to(device), is_train=is_train,
                      white_bg=white_bg, ray_type=ray_type, N_samples=N_samples)
        _, _, blend, pts_ref, weights, xyz_prime, rgb_point, sigma, z_val, dist = out
        delta = None if xyz_prime is None else xyz_prime - pts_ref
        for lst, v in zip(cols, (blend, pts_ref, weights, delta, rgb_point, sigma, z_val, dist)):
            lst.append(v)
    cat = [None if any(v is None for v in lst) else torch.cat(lst) for lst in cols]
    blend, pts_ref, weights, delta, rgb_point, sigma, z_val, dist = cat
    return None, None, blend, pts_ref, weights, delta, None, rgb_point, sigma, z_val, dist


@torch.no_grad()
def render_rays(tensorf_static, tensorf, rays, ts, N_samples, ray_type="ndc"):
    """One no-grad evaluation pass: samples -> static field -> dynamic field -> two-field compositing.
    Returns (raw2outputs 13-tuple, static 10-tuple, dynamic 10-tuple)."""
    xyz, z_vals, valid = sampleXYZ(tensorf, rays, N_samples, ray_type, False)
    fs = tensorf_static(rays, ts, None, xyz, z_vals, valid, is_train=False, ray_type=ray_type, N_samples=N_samples)
    fd = tensorf(rays, ts, None, xyz, z_vals, valid, is_train=False, ray_type=ray_type, N_samples=N_samples)
    out = raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, ray_type)
    return out, fs, fd


# ----------------------------------------------------------------------------- f2: induced flow (renderer.py:1266-1392)


def NDC2world(pts, H, W, f):
    """reference renderer.py:1266-1274."""
    pts_z = 2 / (torch.clamp(pts[..., 2:], min=-1.0, max=1 - 1e-6) - 1)
    pts_x = -pts[..., 0:1] * pts_z * W / 2 / f
    pts_y = -pts[..., 1:2] * pts_z * H / 2 / f
    return torch.cat([pts_x, pts_y, pts_z], -1)


def world2NDC(pts_world, H, W, f):
    """reference renderer.py:1277-1283."""
    o0 = -1.0 / (W / (2.0 * f)) * pts_world[..., 0:1] / pts_world[..., 2:]
    o1 = -1.0 / (H / (2.0 * f)) * pts_world[..., 1:2] / pts_world[..., 2:]
    o2 = 1.0 + 2.0 * 1 / pts_world[..., 2:]
    return torch.cat([o0, o1, o2], -1)


def contract2world(pts_contract):
    """reference renderer.py:1286-1298 (functional form of its in-place masked update)."""
    nrm = pts_contract.abs().amax(-1, keepdim=True)
    return torch.where(nrm > 1.0, pts_contract / nrm * (-1 / (nrm - 2)), pts_contract)


def _project(H, W, f, c2w, pts_map, ray_type):
    """per-RAY tail of render_3d_point (renderer.py:1353-1376): [N,3] tensors only -- a few tiny torch ops whose autograd also
    carries the pose / focal gradients of the static pose passes (train.py:1897-1933)."""
    w2c = c2w[:, :3, :3].transpose(1, 2)
    world = NDC2world(pts_map, H, W, f) if ray_type == "ndc" else contract2world(pts_map)
    world = world - c2w[..., 3]
    cam = torch.sum(world[..., None, :] * w2c[:, :3, :3], -1)
    plane = torch.cat([cam[..., 0:1] / (-cam[..., 2:]) * f + W * 0.5, -cam[..., 1:2] / (-cam[..., 2:]) * f + H * 0.5], -1)
    return plane, world2NDC(cam, H, W, f)


def render_single_3d_point(H, W, f, c2w, pt_NDC):
    """reference renderer.py:1301-1327."""
    plane, cam_ndc = _project(H, W, f, c2w, pt_NDC, "ndc")
    return plane, (cam_ndc[:, 2:] + 1.0) / 2.0


def render_3d_point(H, W, f, c2w, weights, pts, rays, ray_type="ndc", scene_flow=None, clamp=None):
    """reference renderer.py:1330-1376.  The per-SAMPLE part -- sum_s w * pts and sum_s w over [N,S] / [N,S,3] -- is one
    kernel (forward and backward, dynrf_weighted_point_*); `scene_flow` / `clamp` (extensions, default off) fold the caller's
    `pts_ref + scene_flow` and the contract-mode clamp (train.py:1371-1378) into that kernel."""
    from dynrf_b200 import extras as _ex

    pm, acc = _ex.weighted_point(weights, pts, scene_flow, clamp)
    acc = acc[:, None]
    if ray_type == "ndc":
        pts_map = pm + (1.0 - acc) * (rays[:, :3] + rays[:, 3:])
    elif ray_type == "contract":
        far_pts = rays[:, :3] + rays[:, 3:] * 256.0
        nrm = far_pts.abs().amax(-1, keepdim=True)
        far_pts = torch.where(nrm > 1.0, (2 - 1 / nrm) * (far_pts / nrm), far_pts)
        pts_map = pm + (1.0 - acc) * far_pts
    else:
        pts_map = pm
    plane, cam_ndc = _project(H, W, f, c2w, pts_map, ray_type)
    return plane, cam_ndc[:, 2:]


def induce_flow_single(H, W, focal, pose_neighbor, pts_3d_neighbor, pts_2d):
    """reference renderer.py:1379-1386."""
    pts_2d_neighbor, _ = render_single_3d_point(H, W, focal, pose_neighbor, pts_3d_neighbor)
    return pts_2d_neighbor - pts_2d


def induce_flow(H, W, focal, pose_neighbor, weights, pts_3d_neighbor, pts_2d, rays, ray_type="ndc", scene_flow=None, clamp=None):
    """reference renderer.py:1380-1392 -> (induced_flow [N,2], induced_disp [N,1])."""
    pts_2d_neighbor, induced_disp = render_3d_point(H, W, focal, pose_neighbor, weights, pts_3d_neighbor, rays, ray_type,
                                                    scene_flow=scene_flow, clamp=clamp)
    return pts_2d_neighbor - pts_2d, induced_disp


# ----------------------------------------------------------------------------- f1: fused ray generation


def generate_rays(ray_idx, poses_refine9, focal_refine, H, W, ray_type="ndc", near=1.0, uv=None, view_ids=None, poses_mtx=None):
    """train.py:1062-1077 as ONE kernel (+ its backward): ids2pixel -> get_ray_directions_lean -> get_rays_lean ->
    ndc_rays_blender2 (NDC scenes), differentiable w.r.t. the 9-D poses and the focal length.  -> rays_train [B,6]."""
    from dynrf_b200 import extras as _ex

    return _ex.generate_rays(ray_idx, poses_refine9, focal_refine, H, W, ray_type, near, uv=uv, view_ids=view_ids, poses_mtx=poses_mtx)
