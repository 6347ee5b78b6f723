"""Drop-in for the reference's models/tensorBase.py hot-path surface.

Same class / function names, constructor kwargs, attribute names and state_dict keys as
models/tensorBase.py in facebookresearch/robust-dynrf, but `TensorBase.forward` and the
samplers run the sm_100a kernels of libdynrf_b200.so instead of the grid_sample / cumprod /
Linear chain.  Planes and lines keep their logical (1,C,H,W)/(1,C,L,1) shapes and are stored
channel-last (one bilinear tap is one contiguous 16..192-byte read), the two z planes with z as the fastest texel dimension.

Out of scope here (never reached by train.py): AlphaGridMask, updateAlphaMask,
filtering_rays, shrink, compute_alpha, the dead shading modes (MLP_PE, MLP, MLP_Fea_woView, RGB).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from dynrf_b200 import lib as _L
from dynrf_b200 import ops as _ops
from dynrf_b200 import autograd as _ag


def positional_encoding(positions, freqs):
    """reference models/tensorBase.py:13-19 (utility; the kernels encode in shared memory)."""
    bands = (2.0 ** torch.arange(freqs, dtype=torch.float32, device=positions.device))
    scaled = (positions.unsqueeze(-1) * bands).flatten(-2)
    return torch.cat((scaled.sin(), scaled.cos()), dim=-1)


def raw2alpha(sigma, dist):
    """reference models/tensorBase.py:22-34 (utility; the fused kernels do this with a warp scan)."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    trans = torch.cumprod(torch.cat((torch.ones_like(alpha[:, :1]), 1.0 - alpha + 1e-10), -1), -1)
    return alpha, alpha * trans[:, :-1], trans[:, -1:]


class _Head(nn.Module):
    """The three live appearance heads share this container; `mlp` / `mlp_view` index the Linear
    layers exactly like the reference so state_dict keys match (models/tensorBase.py:81-183)."""

    late_view = False
    split_view = False

    def __init__(self, inChanel, viewpe=6, feape=6, featureC=128):
        super().__init__()
        self.viewpe, self.feape = viewpe, feape
        self.in_view = 2 * viewpe * 3 + 3
        k = 2 * feape * inChanel + inChanel
        if self.late_view:
            k += 2 * 10 * 3 + 3 + 2 * 8 + 1
        if not self.split_view:
            k += self.in_view
        self.in_mlpC = k
        body = [nn.Linear(k, featureC), nn.ReLU(inplace=True), nn.Linear(featureC, featureC), nn.ReLU(inplace=True)]
        if self.split_view:
            self.mlp = nn.Sequential(*body)
            self.mlp_view = nn.Sequential(nn.Linear(featureC + self.in_view, 3))
            last = self.mlp_view[-1]
        else:
            self.mlp = nn.Sequential(*body, nn.Linear(featureC, 3))
            last = self.mlp[-1]
        nn.init.constant_(last.bias, 0)

    def _enc(self, x, f):
        return [positional_encoding(x, f)] if f > 0 else []

    def forward(self, pts, viewdirs, features, time):
        """Eager utility with the reference's call signature; TensorBase.forward never calls it."""
        x = [features] + self._enc(features, self.feape)
        if self.late_view:
            x += [pts, positional_encoding(pts, 10), time, positional_encoding(time, 8)]
            viewdirs = viewdirs.detach()
        v = [viewdirs] + self._enc(viewdirs, self.viewpe)
        if not self.split_view:
            x = [features, viewdirs] + self._enc(features, self.feape) + self._enc(viewdirs, self.viewpe)
            return torch.sigmoid(self.mlp(torch.cat(x, -1)))
        h = self.mlp(torch.cat(x, -1))
        return torch.sigmoid(self.mlp_view(torch.cat([h] + v, -1)))


class MLPRender_Fea(_Head):
    pass


class MLPRender_Fea_TimeEmbedding(_Head):
    split_view = True


class MLPRender_Fea_late_view(_Head):
    split_view = True
    late_view = True


_HEADS = {
    "MLP_Fea": (MLPRender_Fea, _L.HEAD_FEA),
    "MLP_Fea_TimeEmbedding": (MLPRender_Fea_TimeEmbedding, _L.HEAD_FEA_TIMEEMB),
    "MLP_Fea_late_view": (MLPRender_Fea_late_view, _L.HEAD_FEA_LATE_VIEW),
}


class TensorBase(nn.Module):
    def __init__(self, aabb, gridSize, tSize, device, density_n_comp=8, appearance_n_comp=24, app_dim=27,
                 shadingMode="MLP_PE", alphaMask=None, near_far=[2.0, 6.0], density_shift=-10, alphaMask_thres=0.001,
                 distance_scale=25, rayMarch_weight_thres=0.0001, pos_pe=6, view_pe=6, fea_pe=6, featureC=128,
                 step_ratio=2.0, fea2denseAct="softplus"):
        super().__init__()
        self.density_n_comp, self.app_n_comp, self.app_dim = density_n_comp, appearance_n_comp, app_dim
        self.aabb, self.alphaMask, self.device = aabb, alphaMask, device
        self.density_shift, self.alphaMask_thres = density_shift, alphaMask_thres
        self.distance_scale, self.rayMarch_weight_thres = distance_scale, rayMarch_weight_thres
        self.fea2denseAct, self.near_far, self.step_ratio, self.tSizeFixed = fea2denseAct, near_far, step_ratio, tSize
        self.update_stepSize(gridSize, tSize)
        self.matMode = [[0, 1], [0, 2], [1, 2]]
        self.vecMode = [2, 1, 0]
        self.comp_w = [1, 1, 1]
        self._packed = {}
        self._cfg_cache = {}
        self.init_svd_volume(gridSize[0], device)
        self.shadingMode, self.pos_pe, self.view_pe, self.fea_pe, self.featureC = shadingMode, pos_pe, view_pe, fea_pe, featureC
        self.init_render_func(shadingMode, pos_pe, view_pe, fea_pe, featureC, device)

    # ------------------------------------------------------------------ construction helpers
    def init_render_func(self, shadingMode, pos_pe, view_pe, fea_pe, featureC, device):
        if shadingMode not in _HEADS:
            # the reference print()s and exit()s (models/tensorBase.py:367-369); its other modes
            # cannot run through its own forward (appendix A) -- raise instead of killing the process
            raise ValueError(f"Unrecognized / unsupported shading module {shadingMode!r}; live modes: {sorted(_HEADS)}")
        cls, self.head_mode = _HEADS[shadingMode]
        self.renderModule = cls(self.app_dim, view_pe, fea_pe, featureC).to(device)

    def update_stepSize(self, gridSize, tSize):
        # reference models/tensorBase.py:373-389
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invaabbSize = 2.0 / self.aabbSize
        self.gridSize = torch.LongTensor(list(gridSize)).to(self.device)
        self.units = self.aabbSize / (self.gridSize - 1)
        self.stepSize = torch.mean(self.units) * self.step_ratio
        self.aabbDiag = torch.sqrt(torch.sum(torch.square(self.aabbSize)))
        self.nSamples = int((self.aabbDiag / self.stepSize).item()) + 1
        self.tSize = torch.unsqueeze(torch.tensor(tSize), 0).to(self.device)
        self._cfg_cache = {}

    def init_svd_volume(self, res, device):
        pass

    def init_one_svd(self, n_component, gridSize, scale, device):
        """planes (1,C,G[m1],G[m0]), lines (1,C,G[v],1), 0.1*randn (reference models/tensoRF.py:26-47), channel-last."""
        planes, lines = [], []
        for (m0, m1), v, c in zip(self.matMode, self.vecMode, n_component):
            planes.append(nn.Parameter(scale * torch.randn((1, c, int(gridSize[m1]), int(gridSize[m0])))))
            lines.append(nn.Parameter(scale * torch.randn((1, c, int(gridSize[v]), 1))))
        planes, lines = nn.ParameterList(planes).to(device), nn.ParameterList(lines).to(device)
        for i, p in enumerate(planes):
            _ops.to_plane_layout_(p, i)     # plane 0 [H][W][C]; planes 1, 2 [W][H][C] (z fastest), see ops.to_plane_layout_
        for p in lines:
            _ops.to_channel_last_(p)
        return planes, lines

    # ------------------------------------------------------------------ native plumbing
    def native_cfg(self, ray_type):
        if ray_type not in self._cfg_cache:
            self._cfg_cache[ray_type] = _ops.field_cfg(self.aabb, self.near_far, self.distance_scale,
                                                       self.rayMarch_weight_thres, self.fea2denseAct, self.density_shift,
                                                       ray_type)
        return self._cfg_cache[ray_type]

    def _layer_table(self):
        """name -> nn.Linear for every layer the kernels consume."""
        rm = self.renderModule
        last = rm.mlp_view[0] if rm.split_view else rm.mlp[4]
        return {"basis": self.basis_mat, "h1": rm.mlp[0], "h2": rm.mlp[2], "h3": last}

    def packed_layers(self):
        """name -> dynrf_layer_t, repacked only when a parameter changed.  One cheap key over all layers
        (data_ptr + in-place version counter) guards the whole table so the steady state costs ~20 us."""
        tbl = self._layer_table()
        key = (_ops.PARAM_EPOCH[0],) + tuple((lin.weight.data_ptr(), lin.weight._version, -1 if lin.bias is None else lin.bias._version)
                                             for lin in tbl.values())
        if getattr(self, "_packed_key", None) == key:
            return self._packed_out
        out = {}
        for name, lin in tbl.items():
            pl = self._packed.get(name)
            if pl is None or pl.weight is not lin.weight or pl.bias is not lin.bias:
                pl = self._packed[name] = _ops.PackedLayer(lin.weight, lin.bias)
            out[name] = pl.refresh()
        self._packed_key, self._packed_out = key, out
        return out

    # ------------------------------------------------------------------ small utilities kept for callers
    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb.to(xyz_sampled.device)[0]) * self.invaabbSize.to(xyz_sampled.device) - 1

    def unnormalize_coord(self, xyz_sampled):
        return (xyz_sampled + 1) / self.invaabbSize.to(xyz_sampled.device) + self.aabb.to(xyz_sampled.device)[0]

    def feature2density(self, density_features):
        if self.fea2denseAct == "softplus":
            return F.softplus(density_features + self.density_shift)
        if self.fea2denseAct == "relu":
            return F.relu(density_features)

    def get_kwargs(self):
        return {
            "aabb": self.aabb, "gridSize": self.gridSize.tolist(), "tSize": self.tSize.item(),
            "density_n_comp": self.density_n_comp, "appearance_n_comp": self.app_n_comp, "app_dim": self.app_dim,
            "density_shift": self.density_shift, "alphaMask_thres": self.alphaMask_thres,
            "distance_scale": self.distance_scale, "rayMarch_weight_thres": self.rayMarch_weight_thres,
            "fea2denseAct": self.fea2denseAct, "near_far": self.near_far, "step_ratio": self.step_ratio,
            "shadingMode": self.shadingMode, "pos_pe": self.pos_pe, "view_pe": self.view_pe, "fea_pe": self.fea_pe,
            "featureC": self.featureC,
        }

    def save(self, se3_poses, focal_ratio_refine, path):
        """Checkpoint format of the reference (models/tensorBase.py:460-470): shapes (1,C,H,W), fp32."""
        kwargs = self.get_kwargs()
        kwargs["se3_poses"] = se3_poses
        kwargs["focal_ratio_refine"] = focal_ratio_refine
        sd = {k: v.contiguous() for k, v in self.state_dict().items()}  # NCHW-contiguous on disk, like the reference
        torch.save({"kwargs": kwargs, "state_dict": sd}, path)

    def load(self, ckpt):
        if "alphaMask.aabb" in ckpt.keys():
            raise NotImplementedError("alpha-mask checkpoints are outside the hot-path scope (never written by train.py)")
        self.load_state_dict(ckpt["state_dict"])  # copy_ keeps the channel-last strides of the live parameters

    # ------------------------------------------------------------------ a2 / a3 samplers
    def _sample(self, rays_o, rays_d, is_train, N_samples, ray_type):
        n_samples = N_samples if N_samples > 0 else self.nSamples
        rays = torch.cat((rays_o, rays_d), -1)
        return _ag.sample_points(self, rays, n_samples, ray_type, is_train)

    def sample_ray_ndc(self, rays_o, rays_d, is_train=True, N_samples=-1):
        xyz, zvec, valid = self._sample(rays_o, rays_d, is_train, N_samples, "ndc")
        return xyz, zvec[None], valid

    def sample_ray_contracted(self, rays_o, rays_d, is_train=True, N_samples=-1):
        xyz, zvec, valid = self._sample(rays_o, rays_d, is_train, N_samples, "contract")
        return xyz, zvec[None], valid

    def sample_ray(self, rays_o, rays_d, is_train=True, N_samples=-1):
        """aabb-entry + fixed-step marching (reference models/tensorBase.py:501-522; unused by the shipped configs).
        -> rays_pts [N,S,3], interpx [N,S], ~mask_outbbox [N,S]."""
        from dynrf_b200 import extras as _ex

        return _ex.sample_ray(self, rays_o, rays_d, is_train, N_samples if N_samples > 0 else self.nSamples)

    # ------------------------------------------------------------------ a10: the field forward
    def forward(self, rays_chunk, ts_chunk, timeembeddings_chunk, xyz_sampled, z_vals, ray_valid, white_bg=True,
                is_train=False, ray_type="ndc", N_samples=-1):
        """Same 10-tuple as the reference (models/tensorBase.py:813-850):
        (None, None, blending|None, pts_ref, weight, xyz_prime|None, rgb, sigma, z_vals, dists*distance_scale)."""
        if timeembeddings_chunk is not None:
            raise NotImplementedError("timeembeddings is always None in the reference (use_time_embedding is never read)")
        return _ag.field_forward(self, rays_chunk, ts_chunk, xyz_sampled, z_vals, ray_valid, ray_type)
