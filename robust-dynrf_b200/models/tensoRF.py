"""Drop-in for the reference's models/tensoRF.py: TensorVMSplit (static) and
TensorVMSplit_TimeEmbedding (dynamic) parameter containers with the reference's constructor
signature, attribute names, state_dict keys and optimizer-group order (models/tensoRF.py:11-61,
277-376).  Feature lookups / MLPs on the ray-march path run in libdynrf_b200.so via
TensorBase.forward; the TV / L1 regularisers run as streaming kernels (row f3, dynrf_b200/extras.py); coarse-to-fine
upsampling, vector_comp_diffs and the scene-flow MLP are plain PyTorch on the same channel-last parameters.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from dynrf_b200 import ops as _ops

from .tensorBase import TensorBase, positional_encoding


def _offdiag_mean_abs(lines):
    """mean |off-diagonal| of the component Gram matrix of each line set (reference vectorDiffs :63-75)."""
    total = 0
    for ln in lines:
        v = ln[0, :, :, 0]
        gram = v @ v.t()
        c = gram.shape[0]
        off = gram.flatten()[1:].view(c - 1, c + 1)[:, :-1]
        total = total + off.abs().mean()
    return total


def _tv(reg, planes, lines):
    """TV_loss_* (reference models/tensoRF.py:100-116, 419-444).  `reg` is the caller's regulariser; the reference always passes
    utils.TVLoss (train.py:1024), recognised here by its `TVLoss_weight` attribute and evaluated -- value and gradient of all
    six tensors -- by ONE streaming kernel (row f3).  Any other callable is applied tensor by tensor as the reference does."""
    if hasattr(reg, "TVLoss_weight"):
        from dynrf_b200 import extras as _ex

        return _ex.tv_loss_vm(float(reg.TVLoss_weight), list(planes), list(lines))
    total = 0
    for p, l in zip(planes, lines):
        total = total + reg(p) * 1e-2 + reg(l) * 1e-3
    return total


class _VMMixin:
    """Shared by both fields."""

    def _resize_vm(self, planes, lines, res_target):
        for i, ((m0, m1), v) in enumerate(zip(self.matMode, self.vecMode)):
            planes[i] = nn.Parameter(F.interpolate(planes[i].data, size=(res_target[m1], res_target[m0]), mode="bilinear",
                                                   align_corners=True))
            lines[i] = nn.Parameter(F.interpolate(lines[i].data, size=(res_target[v], 1), mode="bilinear",
                                                  align_corners=True))
            _ops.to_plane_layout_(planes[i], i)
            _ops.to_channel_last_(lines[i])
        return planes, lines

    def vector_comp_diffs(self):
        return _offdiag_mean_abs(self.density_line) + _offdiag_mean_abs(self.app_line)

    def density_L1(self):
        """reference models/tensoRF.py:80-98 / 378-396 materialises the dense C x Gy x Gx x Gz product (1.7 GB at 300^3);
        here one kernel streams the tables and returns value + gradient (row f3)."""
        from dynrf_b200 import extras as _ex

        return _ex.dense_l1(self, list(self.density_plane), list(self.density_line))

    def TV_loss_density(self, reg):
        return _tv(reg, self.density_plane, self.density_line)

    def TV_loss_app(self, reg):
        return _tv(reg, self.app_plane, self.app_line)

    def shrink(self, new_aabb):
        raise NotImplementedError("shrink needs the alpha-mask machinery, which train.py never invokes (SURVEY appendix A)")


class TensorVMSplit(_VMMixin, TensorBase):
    def __init__(self, aabb, gridSize, tSize, device, **kargs):
        super().__init__(aabb, gridSize, tSize, device, **kargs)

    is_dynamic = False

    def init_svd_volume(self, res, device):
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self.gridSize, 0.1, device)
        self.basis_mat = nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        groups = [(self.density_line, lr_init_spatialxyz), (self.density_plane, lr_init_spatialxyz),
                  (self.app_line, lr_init_spatialxyz), (self.app_plane, lr_init_spatialxyz),
                  (self.basis_mat.parameters(), lr_init_network), (self.renderModule.parameters(), lr_init_network)]
        return [{"params": p, "lr": lr} for p, lr in groups]

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        self.app_plane, self.app_line = self._resize_vm(self.app_plane, self.app_line, res_target)
        self.density_plane, self.density_line = self._resize_vm(self.density_plane, self.density_line, res_target)
        self.update_stepSize(res_target, 1)


class TensorVMSplit_TimeEmbedding(_VMMixin, TensorBase):
    is_dynamic = True

    def __init__(self, aabb, gridSize, tSize, device, **kargs):
        super().__init__(aabb, gridSize, tSize, device, **kargs)
        # coordinate warp (reference :283-287), density / blending MLPs (:289-297), scene flow (:299-313)
        self.layer1 = nn.Linear(1 + 8 * 2, 64).to(device)
        self.layer2 = nn.Linear(64, 30).to(device)
        self.layer3 = nn.Linear((3 + 10 * 2 * 3) + 30, 64).to(device)
        self.layer4 = nn.Linear(64, 64).to(device)
        self.layer5 = nn.Linear(64, 3).to(device)
        kin = sum(self.density_n_comp) * 3 + 3 + 10 * 2 * 3 + 1 + 8 * 2
        self.density_layer1 = nn.Linear(kin, 64).to(device)
        self.density_layer2 = nn.Linear(64, 1).to(device)
        self.blending_layer1 = nn.Linear(kin, 64).to(device)
        self.blending_layer2 = nn.Linear(64, 1).to(device)
        self.scene_flow_mlp = nn.Sequential(nn.Linear(4 * 2 * 4 + 4, 64), nn.ReLU(inplace=True), nn.Linear(64, 64),
                                            nn.ReLU(inplace=True), nn.Linear(64, 64), nn.ReLU(inplace=True),
                                            nn.Linear(64, 6)).to(device)

    def init_svd_volume(self, res, device):
        self.blending_plane, self.blending_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self.gridSize, 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self.gridSize, 0.1, device)
        self.basis_mat = nn.Linear(sum(self.app_n_comp) * 3, self.app_dim, bias=False).to(device)

    def _layer_table(self):
        t = super()._layer_table()
        t.update(layer1=self.layer1, layer2=self.layer2, layer3=self.layer3, layer4=self.layer4, layer5=self.layer5,
                 d1=self.density_layer1, d2=self.density_layer2, b1=self.blending_layer1, b2=self.blending_layer2)
        return t

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        # order matters: train.py:2331,2335 read groups 0 and 4
        sp, nw = lr_init_spatialxyz, lr_init_network
        groups = [(self.density_line, sp), (self.density_plane, sp), (self.blending_line, sp), (self.blending_plane, sp),
                  (self.app_line, sp), (self.app_plane, sp), (self.basis_mat.parameters(), nw),
                  (self.scene_flow_mlp.parameters(), nw)]
        groups += [(getattr(self, f"layer{i}").parameters(), nw) for i in range(1, 6)]
        groups += [(self.density_layer1.parameters(), nw), (self.density_layer2.parameters(), nw),
                   (self.blending_layer1.parameters(), nw), (self.blending_layer2.parameters(), nw),
                   (self.renderModule.parameters(), nw)]
        return [{"params": p, "lr": lr} for p, lr in groups]

    def blending_L1(self):
        from dynrf_b200 import extras as _ex

        return _ex.dense_l1(self, list(self.blending_plane), list(self.blending_line))

    def TV_loss_blending(self, reg):
        return _tv(reg, self.blending_plane, self.blending_line)

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        self.app_plane, self.app_line = self._resize_vm(self.app_plane, self.app_line, res_target)
        self.density_plane, self.density_line = self._resize_vm(self.density_plane, self.density_line, res_target)
        self.blending_plane, self.blending_line = self._resize_vm(self.blending_plane, self.blending_line, res_target)
        self.update_stepSize(res_target, 1)

    # ---- scene-flow MLP (reference :446-519).  "next" row f2: plain PyTorch, not on the accelerated path.
    def _sf_in(self, pts, t):
        xn = self.normalize_coord(pts)
        return torch.cat((xn, positional_encoding(xn, 4), t, positional_encoding(t, 4)), -1)

    def get_forward_backward_scene_flow(self, unnormalized_pts, t_sampled):
        n, s = unnormalized_pts.shape[:2]
        t = t_sampled[:, None].repeat(1, s).view(-1, 1)
        sf = self.scene_flow_mlp(self._sf_in(unnormalized_pts.view(-1, 3), t)).view(n, s, 6)
        return sf[..., 0:3], sf[..., 3:6]

    def get_forward_backward_scene_flow_point(self, unnormalized_pts, t_sampled, weights, rays):
        n, s = unnormalized_pts.shape[:2]
        t = t_sampled[:, None].repeat(1, s).view(-1, 1)
        sf_f = self.scene_flow_mlp(self._sf_in(unnormalized_pts.view(-1, 3), t)).view(n, s, 6)
        sf_b = -sf_f
        acc = weights.sum(-1)[:, None]
        bg = (1.0 - acc) * (rays[:, :3] + rays[:, 3:])
        w = weights[..., None]
        f_pt = (w * (unnormalized_pts + sf_f)).sum(-2) + bg
        b_pt = (w * (unnormalized_pts + sf_b)).sum(-2) + bg
        ref = (w * unnormalized_pts).sum(-2) + bg
        return f_pt, b_pt, ((f_pt - ref).abs() + (b_pt - ref).abs()) / 2.0

    def get_forward_backward_scene_flow_point_single(self, pts_map_NDC, t_sampled):
        sf = self.scene_flow_mlp(self._sf_in(pts_map_NDC, t_sampled[..., None]))
        f, b = sf[..., 0:3], sf[..., 3:6]
        return pts_map_NDC + f, pts_map_NDC + b, f, b

    def warp_coordinate(self, unnormalized_xyz_sampled, t_sampled):
        """xyz [N,S,3], t [N,S] (constant along S, as every caller passes it) -> xyz + delta (reference :521-541)."""
        from dynrf_b200 import autograd as _ag

        return _ag.warp_coordinate(self, unnormalized_xyz_sampled, t_sampled)
