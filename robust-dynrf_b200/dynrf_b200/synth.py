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


def make_models(grid, ray_type, static_mode, device, seed=1):
    """(static TensorVMSplit, dynamic TensorVMSplit_TimeEmbedding) built as train.py:874-922 builds them; identical on every
    rank for the same seed (parameters are drawn on the CPU generator, then moved)."""
    import models.tensoRF as T

    dev = torch.device(device)
    aabb = torch.tensor(NDC_AABB if ray_type == "ndc" else CONTRACT_AABB, device=dev)
    near_far = [0.0, 1.0] if ray_type == "ndc" else [0.0, 256.0]
    kw = dict(COMMON, near_far=near_far)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        cpu = torch.device("cpu")
        static = T.TensorVMSplit(aabb.cpu(), list(grid), 12, cpu, shadingMode=static_mode, fea_pe=2, **kw)
        dyn = T.TensorVMSplit_TimeEmbedding(aabb.cpu(), list(grid), 12, cpu, shadingMode="MLP_Fea_late_view", fea_pe=0, **kw)
    finally:
        torch.random.set_rng_state(state)
    sds, sdd = static.state_dict(), dyn.state_dict()
    static = T.TensorVMSplit(aabb, list(grid), 12, dev, shadingMode=static_mode, fea_pe=2, **kw)
    dyn = T.TensorVMSplit_TimeEmbedding(aabb, list(grid), 12, dev, shadingMode="MLP_Fea_late_view", fea_pe=0, **kw)
    static.load_state_dict(sds)
    dyn.load_state_dict(sdd)
    return static, dyn
