"""Synthetic workloads for bench.py's native arm and the tools: seeded rays with the value distributions the survey
states and the two drop-in models at reference initialisation (planes / lines 0.1 * randn, nn.Linear default init, last head bias 0
-- what the constructors in models/ do).  Nothing from the test infrastructure is involved."""
from __future__ import annotations

import torch

NDC_AABB = [[-1.5, -1.67, -1.0], [1.5, 1.67, 1.0]]      # dataLoader/nvidia.py:246-251
CONTRACT_AABB = [[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]]
COMMON = dict(density_n_comp=[16, 4, 4], appearance_n_comp=[48, 12, 12], app_dim=27, alphaMask_thres=1e-4, density_shift=-10,
              distance_scale=25, pos_pe=6, view_pe=0, featureC=128, step_ratio=2.0, fea2denseAct="relu")


def synth_rays(n, ray_type, gen, spread=0.1):
    """train-order rays: spatially random origins (train.py:88-93 permutes the pixels of all frames)."""
    if ray_type == "ndc":
        o = torch.cat((torch.rand(n, 2, generator=gen) * 2 - 1, -torch.ones(n, 1)), -1)
        d = torch.cat((spread * torch.randn(n, 2, generator=gen), 2 * torch.ones(n, 1)), -1)
    else:
        o = 0.05 * torch.randn(n, 3, generator=gen)
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1)
    ts = torch.rand(n, generator=gen) * 2 - 1
    return torch.cat((o, d), -1), ts


def _head_in_dim(static_mode: str, dynamic: bool) -> int:
    fea_pe, app_dim = (0 if dynamic else 2), 27
    base = 2 * fea_pe * app_dim + app_dim
    if dynamic:                       # MLP_Fea_late_view: + pts, PE(pts, 10), t, PE(t, 8)
        return base + 60 + 3 + 16 + 1
    return base + 3 if static_mode == "MLP_Fea" else base


def init_state_dict(dynamic: bool, static_mode: str, grid, gen: torch.Generator, scale: float = 0.1):
    """Random parameters with the reference's shapes, key names and init distributions (planes / lines 0.1 * randn,
    models/tensoRF.py:26-47; nn.Linear default init U(+-1/sqrt(k)); last head bias 0, models/tensorBase.py:99,130,164), drawn
    from ONE generator in a fixed order so that every rank, every run and every round of this repo benches the same model."""
    import math

    mat_mode, vec_mode = ((0, 1), (0, 2), (1, 2)), (2, 1, 0)
    sd = {}

    def vm(name, comps):
        for i in range(3):
            m0, m1 = mat_mode[i]
            sd[f"{name}_plane.{i}"] = scale * torch.randn((1, comps[i], grid[m1], grid[m0]), generator=gen)
            sd[f"{name}_line.{i}"] = scale * torch.randn((1, comps[i], grid[vec_mode[i]], 1), generator=gen)

    def lin(name, k, m, bias=True, zero_bias=False):
        bound = 1.0 / math.sqrt(k)
        sd[f"{name}.weight"] = (torch.rand((m, k), generator=gen) * 2 - 1) * bound
        if bias:
            sd[f"{name}.bias"] = torch.zeros(m) if zero_bias else (torch.rand((m,), generator=gen) * 2 - 1) * bound

    den, app = COMMON["density_n_comp"], COMMON["appearance_n_comp"]
    if dynamic:
        vm("blending", den)
    vm("density", den)
    vm("app", app)
    lin("basis_mat", sum(app) * (3 if dynamic else 1), 27, bias=False)
    lin("renderModule.mlp.0", _head_in_dim(static_mode, dynamic), 128)
    lin("renderModule.mlp.2", 128, 128)
    if not dynamic and static_mode == "MLP_Fea":
        lin("renderModule.mlp.4", 128, 3, zero_bias=True)
    else:
        lin("renderModule.mlp_view.0", 131, 3, zero_bias=True)
    if dynamic:
        lin("layer1", 17, 64)
        lin("layer2", 64, 30)
        lin("layer3", 93, 64)
        lin("layer4", 64, 64)
        lin("layer5", 64, 3)
        kin = sum(den) * 3 + 3 + 60 + 1 + 16
        lin("density_layer1", kin, 64)
        lin("density_layer2", 64, 1)
        lin("blending_layer1", kin, 64)
        lin("blending_layer2", 64, 1)
        lin("scene_flow_mlp.0", 36, 64)
        lin("scene_flow_mlp.2", 64, 64)
        lin("scene_flow_mlp.4", 64, 64)
        lin("scene_flow_mlp.6", 64, 6)
    return sd


def make_models(grid, ray_type, static_mode, device, seed=1):
    """(static TensorVMSplit, dynamic TensorVMSplit_TimeEmbedding) with the hyper-parameters train.py:874-922 passes; parameters
    come from init_state_dict (static first, then dynamic, one generator) -- identical on every rank for the same seed, and the
    same model round 1 of this repo benched."""
    import models.tensoRF as T

    dev = torch.device(device)
    aabb = torch.tensor(NDC_AABB if ray_type == "ndc" else CONTRACT_AABB, device=dev)
    near_far = [0.0, 1.0] if ray_type == "ndc" else [0.0, 256.0]
    kw = dict(COMMON, near_far=near_far)
    gen = torch.Generator().manual_seed(seed)
    sds = init_state_dict(False, static_mode, list(grid), gen)
    sdd = init_state_dict(True, static_mode, list(grid), gen)
    static = T.TensorVMSplit(aabb, list(grid), 12, dev, shadingMode=static_mode, fea_pe=2, **kw)
    dyn = T.TensorVMSplit_TimeEmbedding(aabb, list(grid), 12, dev, shadingMode="MLP_Fea_late_view", fea_pe=0, **kw)
    static.load_state_dict(sds)
    dyn.load_state_dict(sdd)
    return static, dyn
