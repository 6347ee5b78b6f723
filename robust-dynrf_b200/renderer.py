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
