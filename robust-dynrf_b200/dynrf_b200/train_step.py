"""The step bench.py times: one pass of the hot path over a ray batch (PASS 1).

  mode "fwd"   : sampleXYZ -> static field -> dynamic field -> raw2outputs, no autograd.
  mode "train" : the same pass with autograd, an MSE loss on the composited colours, backward to every
                 parameter of both fields (+ one gradient all-reduce when world > 1).

`fused=True` (bench `value`) drives the kernels directly and lets the static density kernel generate its
samples in-kernel; `fused=False` (bench `e2e`) goes through the reference-shaped public API
(renderer.sampleXYZ, TensorBase.forward, renderer.raw2outputs).
"""
from __future__ import annotations

import torch

from . import ops


def backward_available() -> bool:
    try:
        from . import backward  # noqa: F401

        return backward.available()
    except ImportError:
        return False


class Step:
    def __init__(self, static, dyn, n_samples, ray_type, mode, world=1):
        if mode == "train" and not backward_available():
            raise RuntimeError("train mode needs the backward kernels (dynrf_b200.backward)")
        self.static, self.dyn, self.s, self.rt, self.mode, self.world = static, dyn, n_samples, ray_type, mode, world
        self.gather_buf = None

    def _gather(self, out):
        from .parallel import gather_ray_maps

        return gather_ray_maps(out, self.world)

    @torch.no_grad()
    def forward_fused(self, rays, ts):
        cfg = self.dyn.native_cfg(self.rt)
        zvec = ops.make_zvec(cfg, self.s, self.rt, True, rays.device)
        xyz, z, valid = ops.sample_points(cfg, rays, zvec)
        st = ops.static_forward_raw(self.static, rays, None, None, None, self.rt, zvec=zvec)
        st["xyz"] = xyz
        ops.static_app_raw(self.static, st, self.rt)
        d = ops.dynamic_forward_raw(self.dyn, rays, ts, xyz, z, valid, self.rt)
        out, _, _, _, _ = ops.composite_raw(st["rgb"], st["sigma"], d["rgb"], d["sigma"], d["dists"], d["blend"], z, rays,
                                            False, self.rt)
        return out

    @torch.no_grad()
    def forward_api(self, rays, ts):
        import renderer as R

        xyz, z, valid = R.sampleXYZ(self.dyn, rays, self.s, self.rt, True)
        fs = self.static(rays, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        fd = self.dyn(rays, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        o = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, self.rt)
        return torch.cat((o[0], o[1][:, None], o[2][:, None], o[4], o[5][:, None], o[6][:, None], o[8], o[9][:, None],
                          o[10][:, None], o[12][:, None]), -1)

    def __call__(self, rays, ts, target, fused=True):
        if self.mode == "fwd":
            out = self.forward_fused(rays, ts) if fused else self.forward_api(rays, ts)
            return self._gather(out)
        from . import backward as B

        return self._gather(B.train_step(self, rays, ts, target))
