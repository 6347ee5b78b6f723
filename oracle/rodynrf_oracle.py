"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RoDynRF ray-march hot path.

A functional restatement (PyTorch fp32, CPU) of the one path this repo accelerates:

    renderer.sampleXYZ -> TensorBase.forward (static) -> TensorBase.forward (dynamic)
    -> renderer.raw2outputs

in facebookresearch/robust-dynrf.  Every function cites the reference file:line it
restates (paths relative to the reference checkout).  The third-party arithmetic on the
path is PyTorch itself (F.grid_sample bilinear / zeros padding / align_corners=True,
torch.cumprod, nn.Linear); it is used here exactly as the reference uses it so this
module is also the "port" CPU baseline timed by bench.py.

Pinning: tests/golden/make_golden.py imports the UNMODIFIED reference from
/root/reference (in the build container only), runs it on seeded inputs and commits the
input/output vectors under tests/golden/*.npz; tests/test_oracle_golden.py checks every
function of this module against those vectors.  Parity status: PINNED for a1-a19 and
a21; `OctreeRender_trilinear_fast` (a20) raises TypeError in the reference as shipped
(renderer.py:94,133) so that one symbol is "parity unpinned".

Nothing in the product package imports this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# models/tensorBase.py:326-327
MAT_MODE = ((0, 1), (0, 2), (1, 2))
VEC_MODE = (2, 1, 0)


@dataclass
class FieldCfg:
    """Hyper-parameters of one field (kwargs of TensorBase.__init__, models/tensorBase.py:282-304)."""

    aabb: Tensor  # [2,3]
    near_far: Tuple[float, float] = (0.0, 1.0)
    density_n_comp: Sequence[int] = (16, 4, 4)
    app_n_comp: Sequence[int] = (48, 12, 12)
    app_dim: int = 27
    shadingMode: str = "MLP_Fea"
    fea_pe: int = 2
    view_pe: int = 0
    featureC: int = 128
    distance_scale: float = 25.0
    rayMarch_weight_thres: float = 1e-4
    fea2denseAct: str = "relu"
    density_shift: float = -10.0
    dynamic: bool = False

    def inv_aabb_size(self) -> Tensor:
        # models/tensorBase.py:376-377
        return 2.0 / (self.aabb[1] - self.aabb[0])


# --------------------------------------------------------------------------- basics


def positional_encoding(positions: Tensor, freqs: int) -> Tensor:
    """models/tensorBase.py:13-19 -- [sin(x_d*2^f) (dim-major, freq-minor) ..., cos(...) ...]."""
    bands = (2.0 ** torch.arange(freqs, dtype=torch.float32)).to(positions.device)
    scaled = (positions.unsqueeze(-1) * bands).flatten(-2)
    return torch.cat((scaled.sin(), scaled.cos()), dim=-1)


def raw2alpha(sigma: Tensor, dist: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """models/tensorBase.py:22-34 -- alpha, weights = alpha * exclusive cumprod(1-alpha+1e-10), last T."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    ones = torch.ones(alpha.shape[0], 1, device=alpha.device)
    trans = torch.cumprod(torch.cat((ones, 1.0 - alpha + 1e-10), dim=-1), dim=-1)
    return alpha, alpha * trans[:, :-1], trans[:, -1:]


def normalize_coord(cfg: FieldCfg, xyz: Tensor) -> Tensor:
    """models/tensorBase.py:425-428."""
    return (xyz - cfg.aabb[0]) * cfg.inv_aabb_size() - 1


def unnormalize_coord(cfg: FieldCfg, xn: Tensor) -> Tensor:
    """models/tensorBase.py:430-433."""
    return (xn + 1) / cfg.inv_aabb_size() + cfg.aabb[0]


def feature2density(cfg: FieldCfg, feat: Tensor) -> Tensor:
    """models/tensorBase.py:678-682."""
    if cfg.fea2denseAct == "softplus":
        return F.softplus(feat + cfg.density_shift)
    if cfg.fea2denseAct == "relu":
        return F.relu(feat)
    raise ValueError(cfg.fea2denseAct)


# --------------------------------------------------------------------------- samplers


def sample_ray_ndc(
    cfg: FieldCfg, rays_o: Tensor, rays_d: Tensor, n_samples: int, is_train: bool = False,
    jitter: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """models/tensorBase.py:487-499.  `jitter` ([1,S] uniforms) replaces the reference's
    torch.rand_like draw when given; otherwise the RNG is consumed exactly as the
    reference does (one rand_like of shape [1,S])."""
    near, far = cfg.near_far
    z = torch.linspace(near, far, n_samples).unsqueeze(0).to(rays_o)
    if is_train:
        u = torch.rand_like(z) if jitter is None else jitter.to(z)
        z = z + u * ((far - near) / n_samples)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
    outside = ((cfg.aabb[0] > pts) | (pts > cfg.aabb[1])).any(dim=-1)
    return pts, z, ~outside


def sample_ray_contracted(
    cfg: FieldCfg, rays_o: Tensor, rays_d: Tensor, n_samples: int, is_train: bool = False,
    jitter_inner: Optional[Tensor] = None, jitter_outer: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Tensor]:
    """models/tensorBase.py:524-559.  Inner half: midpoints of linspace(near,2,Sin+1); outer
    half: inverse-depth midpoints out to `far`; then L-inf contraction of points with
    |p|_inf > 1.  All samples are reported valid."""
    near, far = cfg.near_far
    n_in = n_samples - n_samples // 2
    n_out = n_samples // 2
    edges = torch.linspace(near, 2.0, n_in + 1).unsqueeze(0).to(rays_o)
    if is_train:
        u = torch.rand_like(edges) if jitter_inner is None else jitter_inner.to(edges)
        edges = edges.clone()
        edges[:, :-1] += (u * ((2.0 - near) / n_in))[:, :-1]
    z_in = (edges[:, 1:] + edges[:, :-1]) * 0.5
    r = torch.arange(n_out + 1, device=rays_o.device)[None].float()
    if is_train:
        u = torch.rand_like(r) if jitter_outer is None else jitter_outer.to(r)
        r[:, :-1] += u[:, :-1]
    r = torch.flip(r, [1])
    r = (r[:, 1:] + r[:, :-1]) * 0.5
    z_out = 1.0 / (1 / far + (1 / 2.0 - 1 / far) * r / n_out).to(rays_o.device)
    z = torch.cat((z_in, z_out), -1)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
    linf = pts.abs().amax(dim=-1)
    far_mask = linf > 1.0
    pts = pts.clone()
    pts[far_mask] = (2 - 1 / linf[far_mask])[..., None] * (pts[far_mask] / linf[far_mask][..., None])
    return pts, z, torch.ones_like(linf, dtype=torch.bool)


def sampleXYZ(cfg: FieldCfg, rays: Tensor, n_samples: int, ray_type: str = "ndc", is_train: bool = False,
              **jit) -> Tuple[Tensor, Tensor, Tensor]:
    """renderer.py:147-170 -- dispatch + tile z_vals to [N,S]."""
    if ray_type == "ndc":
        xyz, z, valid = sample_ray_ndc(cfg, rays[:, :3], rays[:, 3:6], n_samples, is_train, **jit)
    elif ray_type == "contract":
        xyz, z, valid = sample_ray_contracted(cfg, rays[:, :3], rays[:, 3:6], n_samples, is_train, **jit)
    else:
        raise NotImplementedError("sample_ray (models/tensorBase.py:501-522) is unused by every config")
    return xyz, torch.tile(z, (xyz.shape[0], 1)), valid


# --------------------------------------------------------------------------- VM lookup


def _grid_coords(xn: Tensor) -> Tuple[Tensor, Tensor]:
    """Sampling grids the reference builds at models/tensoRF.py:120-136 (and :549-565 etc.)."""
    plane = torch.stack([xn[..., list(m)] for m in MAT_MODE]).view(3, -1, 1, 2)
    line = torch.stack([xn[..., v] for v in VEC_MODE])
    line = torch.stack((torch.zeros_like(line), line), dim=-1).view(3, -1, 1, 2)
    return plane, line


def vm_products(planes: Sequence[Tensor], lines: Sequence[Tensor], xn: Tensor,
                strides: Sequence[int] = (1,)) -> Tensor:
    """plane*line products for every component, [sum_i C_i * len(strides), P].

    Static: models/tensoRF.py:139-152 / :176-194 (stride 1).  Dynamic: the stride-1/2/4
    sub-sampled views `plane[:, :, ::s, ::s]`, `line[:, :, ::s]` of models/tensoRF.py:
    568-618, :671-721, :759-809, concatenated stride-major then plane-major."""
    cp, cl = _grid_coords(xn)
    n = xn.shape[0]
    pf, lf = [], []
    for s in strides:
        for i in range(3):
            pl = planes[i] if s == 1 else planes[i][:, :, ::s, ::s]
            ln = lines[i] if s == 1 else lines[i][:, :, ::s]
            pf.append(F.grid_sample(pl, cp[[i]], align_corners=True).view(-1, n))
            lf.append(F.grid_sample(ln, cl[[i]], align_corners=True).view(-1, n))
    return torch.cat(pf) * torch.cat(lf)


def _plist(sd: Dict[str, Tensor], name: str) -> List[Tensor]:
    return [sd[f"{name}.{i}"] for i in range(3)]


def _linear(sd: Dict[str, Tensor], name: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def static_density_feature(sd, xn: Tensor) -> Tensor:
    """TensorVMSplit.compute_densityfeature, models/tensoRF.py:118-154."""
    return vm_products(_plist(sd, "density_plane"), _plist(sd, "density_line"), xn).sum(0)


def static_app_feature(sd, xn: Tensor) -> Tensor:
    """TensorVMSplit.compute_appfeature, models/tensoRF.py:156-196."""
    prod = vm_products(_plist(sd, "app_plane"), _plist(sd, "app_line"), xn)
    return _linear(sd, "basis_mat", prod.T)


def warp_coordinate(cfg: FieldCfg, sd, xyz: Tensor, t: Tensor) -> Tensor:
    """TensorVMSplit_TimeEmbedding.warp_coordinate, models/tensoRF.py:521-541 (un-normalised in/out)."""
    tin = torch.cat((t[..., None], positional_encoding(t[..., None], 8)), dim=-1)
    t_out = _linear(sd, "layer2", F.relu(_linear(sd, "layer1", tin)))
    xn = normalize_coord(cfg, xyz)
    xin = torch.cat((xn, positional_encoding(xn, 10), t_out), dim=-1)
    h = F.relu(_linear(sd, "layer3", xin))
    h = F.relu(_linear(sd, "layer4", h))
    return xyz + _linear(sd, "layer5", h)


def _warped_normalized(cfg, sd, xn, t):
    # models/tensoRF.py:544-546, :647-649, :735-737
    return normalize_coord(cfg, warp_coordinate(cfg, sd, unnormalize_coord(cfg, xn), t))


def _dyn_mlp_feature(cfg, sd, which: str, xn: Tensor, t: Tensor) -> Tensor:
    xp = _warped_normalized(cfg, sd, xn, t)
    prod = vm_products(_plist(sd, f"{which}_plane"), _plist(sd, f"{which}_line"), xp, (1, 2, 4))
    x = torch.cat((prod.T, xn, positional_encoding(xn, 10), t[..., None],
                   positional_encoding(t[..., None], 8)), dim=-1)
    return _linear(sd, f"{which}_layer2", F.relu(_linear(sd, f"{which}_layer1", x)))[..., 0]


def dynamic_density_feature(cfg, sd, xn, t):
    """TensorVMSplit_TimeEmbedding.compute_densityfeature, models/tensoRF.py:646-732."""
    return _dyn_mlp_feature(cfg, sd, "density", xn, t)


def dynamic_blending_feature(cfg, sd, xn, t):
    """TensorVMSplit_TimeEmbedding.compute_blendingfeature, models/tensoRF.py:543-629."""
    return _dyn_mlp_feature(cfg, sd, "blending", xn, t)


def dynamic_app_feature(cfg, sd, xn, t):
    """TensorVMSplit_TimeEmbedding.compute_appfeature, models/tensoRF.py:734-811."""
    xp = _warped_normalized(cfg, sd, xn, t)
    prod = vm_products(_plist(sd, "app_plane"), _plist(sd, "app_line"), xp, (1, 2, 4))
    return _linear(sd, "basis_mat", prod.T)


# --------------------------------------------------------------------------- appearance heads


def render_head(cfg: FieldCfg, sd, pts: Tensor, viewdirs: Tensor, feat: Tensor, time: Tensor) -> Tensor:
    """The three live shading heads:
    MLPRender_Fea models/tensorBase.py:81-111, MLPRender_Fea_TimeEmbedding :114-145,
    MLPRender_Fea_late_view :148-183."""
    fpe = [positional_encoding(feat, cfg.fea_pe)] if cfg.fea_pe > 0 else []
    vpe = [positional_encoding(viewdirs, cfg.view_pe)] if cfg.view_pe > 0 else []
    mode = cfg.shadingMode
    if mode == "MLP_Fea":
        x = torch.cat([feat, viewdirs] + fpe + vpe, dim=-1)
        h = F.relu(_linear(sd, "renderModule.mlp.0", x))
        h = F.relu(_linear(sd, "renderModule.mlp.2", h))
        return torch.sigmoid(_linear(sd, "renderModule.mlp.4", h))
    if mode == "MLP_Fea_TimeEmbedding":
        x = torch.cat([feat] + fpe, dim=-1)
    elif mode == "MLP_Fea_late_view":
        x = torch.cat([feat] + fpe + [pts, positional_encoding(pts, 10), time, positional_encoding(time, 8)], dim=-1)
        viewdirs = viewdirs.detach()  # models/tensorBase.py:174-176: no gradient to the view direction
        vpe = [positional_encoding(viewdirs, cfg.view_pe)] if cfg.view_pe > 0 else []
    else:
        raise ValueError(f"shading mode {mode!r} is not live in the reference (SURVEY appendix A)")
    h = F.relu(_linear(sd, "renderModule.mlp.0", x))
    h = F.relu(_linear(sd, "renderModule.mlp.2", h))
    return torch.sigmoid(_linear(sd, "renderModule.mlp_view.0", torch.cat([h, viewdirs] + vpe, dim=-1)))


# --------------------------------------------------------------------------- field forward


def field_forward(cfg: FieldCfg, sd, rays: Tensor, ts: Tensor, xyz: Tensor, z_vals: Tensor, ray_valid: Tensor,
                  ray_type: str = "ndc"):
    """TensorBase.forward, models/tensorBase.py:704-850 (timeembeddings is always None, appendix A).

    Returns the reference's 10-tuple
    (None, None, blending|None, pts_ref, weight, xyz_prime|None, rgb, sigma, z_vals, dists*scale)."""
    n, s = xyz.shape[:2]
    viewdirs = rays[:, 3:6]
    dists = torch.cat((z_vals[:, 1:] - z_vals[:, :-1], torch.zeros_like(z_vals[:, :1])), dim=-1)
    if ray_type in ("ndc", "contract"):
        norm = torch.norm(viewdirs, dim=-1, keepdim=True)
        dists = dists * norm
        viewdirs = viewdirs / norm
    viewdirs = viewdirs.view(-1, 1, 3).expand(xyz.shape)
    t_tiled = torch.tile(ts.unsqueeze(-1), (1, s))

    dev = xyz.device   # CPU in the tests; bench.py's `ref_cuda_eager` leg runs the same restatement on the GPU
    sigma = torch.zeros(n, s, device=dev)
    rgb = torch.zeros(n, s, 3, device=dev)
    blending = torch.zeros(n, s, 1, device=dev)
    xn = normalize_coord(cfg, xyz)
    if ray_valid.any():
        if cfg.dynamic:
            f = dynamic_density_feature(cfg, sd, xn[ray_valid], t_tiled[ray_valid])
        else:
            f = static_density_feature(sd, xn[ray_valid])
        sigma[ray_valid] = feature2density(cfg, f)

    _, weight, _ = raw2alpha(sigma, dists * cfg.distance_scale)
    app_mask = weight > cfg.rayMarch_weight_thres
    if app_mask.any():
        if cfg.dynamic:
            af = dynamic_app_feature(cfg, sd, xn[app_mask], t_tiled[app_mask])
        else:
            af = static_app_feature(sd, xn[app_mask])
        rgb[app_mask] = render_head(cfg, sd, xn[app_mask], viewdirs[app_mask], af, t_tiled[app_mask][..., None])

    if not cfg.dynamic:
        return None, None, None, xyz, weight, None, rgb, sigma, z_vals, dists * cfg.distance_scale

    xyz_prime = warp_coordinate(cfg, sd, xyz, t_tiled)
    if ray_valid.any():
        bf = dynamic_blending_feature(cfg, sd, xn[ray_valid], t_tiled[ray_valid])
        blending[ray_valid] = torch.sigmoid(bf)[:, None]
    return None, None, blending[..., 0], xyz, weight, xyz_prime, rgb, sigma, z_vals, dists * cfg.distance_scale


# --------------------------------------------------------------------------- two-field compositing


def raw2outputs(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays, is_train=False, ray_type="ndc",
                white_bg_draw: Optional[bool] = None):
    """renderer.raw2outputs, renderer.py:173-315 -> 13-tuple.

    `white_bg_draw` replaces the host-side `torch.rand((1,)) < 0.5` coin of renderer.py:269
    when given (tests); otherwise the coin is drawn exactly as the reference does."""
    a_d = 1.0 - torch.exp(-sigma_d * dists)
    a_s = 1.0 - torch.exp(-sigma_s * dists)

    def excl_cumprod(x):
        ones = torch.ones((x.shape[0], 1), device=x.device)
        return torch.cumprod(torch.cat((ones, x), -1), -1)[:, :-1]

    t_d = excl_cumprod(1.0 - a_d + 1e-10)
    t_s = excl_cumprod(1.0 - a_s + 1e-10)
    t_f = excl_cumprod((1.0 - a_d * blending) * (1.0 - a_s * (1.0 - blending)) + 1e-10)

    w_d = a_d * t_d
    w_s = a_s * t_s
    w_d = w_d / (w_d.sum(-1, keepdim=True) + 1e-10)
    w_f = (a_d * blending + a_s * (1.0 - blending)) * t_f

    rgb_map_d = (w_d[..., None] * rgb_d).sum(-2)
    rgb_map_s = (w_s[..., None] * rgb_s).sum(-2)
    rgb_map_f = ((t_f * a_d * blending)[..., None] * rgb_d + (t_f * a_s * (1.0 - blending))[..., None] * rgb_s).sum(-2)

    acc_d, acc_s, acc_f = w_d.sum(-1), w_s.sum(-1), w_f.sum(-1)

    if is_train:
        coin = bool(torch.rand((1,)) < 0.5) if white_bg_draw is None else white_bg_draw
        if coin:
            rgb_map_d = rgb_map_d + (1.0 - acc_d[..., None])
            rgb_map_s = rgb_map_s + (1.0 - acc_s[..., None])
            rgb_map_f = rgb_map_f + torch.relu(1.0 - acc_f[..., None])

    depth_d = (w_d * z_vals).sum(-1)
    depth_s = (w_s * z_vals).sum(-1)
    depth_f = (w_f * z_vals).sum(-1)
    if ray_type == "ndc":
        bg = rays[..., 2] + rays[..., -1]
    elif ray_type == "contract":
        bg = 256.0
    else:
        bg = None
    if bg is not None:
        depth_d = depth_d + (1.0 - acc_d) * bg
        depth_s = depth_s + (1.0 - acc_s) * bg
        depth_f = depth_f + torch.relu(1.0 - acc_f) * bg

    dyn = (w_f * blending).sum(-1) + torch.relu(1.0 - acc_f) * 0.0
    return (rgb_map_f.clamp(0, 1), depth_f, acc_f, w_f, rgb_map_s.clamp(0, 1), depth_s, acc_s, w_s,
            rgb_map_d.clamp(0, 1), depth_d, acc_d, w_d, dyn)


# --------------------------------------------------------------------------- whole path + synthetic models


def render_rays(cfg_s: FieldCfg, sd_s, cfg_d: FieldCfg, sd_d, rays, ts, n_samples, ray_type="ndc"):
    """One evaluation pass of the hot path (renderer.py:1061-1130 / train.py:1092-1162 shape)."""
    xyz, z, valid = sampleXYZ(cfg_d, rays, n_samples, ray_type, False)
    fs = field_forward(cfg_s, sd_s, rays, ts, xyz, z, valid, ray_type)
    fd = field_forward(cfg_d, sd_d, rays, ts, xyz, z, valid, ray_type)
    out = raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, ray_type)
    return xyz, z, valid, fs, fd, out


def head_in_dim(cfg: FieldCfg) -> int:
    if cfg.shadingMode == "MLP_Fea":
        return 2 * cfg.view_pe * 3 + 2 * cfg.fea_pe * cfg.app_dim + 3 + cfg.app_dim
    if cfg.shadingMode == "MLP_Fea_TimeEmbedding":
        return 2 * cfg.fea_pe * cfg.app_dim + cfg.app_dim
    if cfg.shadingMode == "MLP_Fea_late_view":
        return 2 * cfg.fea_pe * cfg.app_dim + cfg.app_dim + 2 * 10 * 3 + 3 + 2 * 8 + 1
    raise ValueError(cfg.shadingMode)


def init_state_dict(cfg: FieldCfg, grid: Sequence[int], gen: torch.Generator, scale: float = 0.1,
                    include_scene_flow: bool = True) -> Dict[str, Tensor]:
    """Random parameters with the reference's shapes, key names and init distributions
    (models/tensoRF.py:26-47,283-327; nn.Linear default init; last head bias 0,
    models/tensorBase.py:99,130,164).  Not bit-identical to the reference's RNG order --
    used for synthetic workloads, never for golden vectors."""
    sd: Dict[str, Tensor] = {}

    def vm(name, comps):
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            sd[f"{name}_plane.{i}"] = scale * torch.randn((1, comps[i], grid[m1], grid[m0]), generator=gen)
            sd[f"{name}_line.{i}"] = scale * torch.randn((1, comps[i], grid[VEC_MODE[i]], 1), generator=gen)

    def lin(name, k, m, bias=True, zero_bias=False):
        bound = 1.0 / math.sqrt(k)
        sd[f"{name}.weight"] = (torch.rand((m, k), generator=gen) * 2 - 1) * bound
        if bias:
            sd[f"{name}.bias"] = torch.zeros(m) if zero_bias else (torch.rand((m,), generator=gen) * 2 - 1) * bound

    if cfg.dynamic:
        vm("blending", cfg.density_n_comp)
    vm("density", cfg.density_n_comp)
    vm("app", cfg.app_n_comp)
    lin("basis_mat", sum(cfg.app_n_comp) * (3 if cfg.dynamic else 1), cfg.app_dim, bias=False)
    k = head_in_dim(cfg)
    lin("renderModule.mlp.0", k, cfg.featureC)
    lin("renderModule.mlp.2", cfg.featureC, cfg.featureC)
    if cfg.shadingMode == "MLP_Fea":
        lin("renderModule.mlp.4", cfg.featureC, 3, zero_bias=True)
    else:
        lin("renderModule.mlp_view.0", cfg.featureC + 3 + 2 * cfg.view_pe * 3, 3, zero_bias=True)
    if cfg.dynamic:
        lin("layer1", 17, 64)
        lin("layer2", 64, 30)
        lin("layer3", 93, 64)
        lin("layer4", 64, 64)
        lin("layer5", 64, 3)
        kin = sum(cfg.density_n_comp) * 3 + 3 + 60 + 1 + 16
        lin("density_layer1", kin, 64)
        lin("density_layer2", 64, 1)
        lin("blending_layer1", kin, 64)
        lin("blending_layer2", 64, 1)
        if include_scene_flow:
            lin("scene_flow_mlp.0", 36, 64)
            lin("scene_flow_mlp.2", 64, 64)
            lin("scene_flow_mlp.4", 64, 64)
            lin("scene_flow_mlp.6", 64, 6)
    return sd
