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
    """graph=True captures the whole step (public-API calls, autograd backward, Adam) into CUDA graphs after three eager
    warm-up steps and replays them: the step is ~120 kernel launches and ~300 small tensor ops, and at ~20 ms of GPU
    work per step the Python / launch overhead is of the same size.  With world > 1 the collectives stay eager between
    two graphs (forward+backward | optimizer)."""

    def __init__(self, static, dyn, n_samples, ray_type, mode, world=1, graph=False, jitter=True, rays_grad=False):
        if mode == "train" and not backward_available():
            raise RuntimeError("train mode needs the backward kernels (dynrf_b200.backward)")
        self.static, self.dyn, self.s, self.rt, self.mode, self.world = static, dyn, n_samples, ray_type, mode, world
        self.gather_buf = None
        self.use_graph = graph
        self.jitter = jitter    # stratified-sampling jitter (is_train of renderer.sampleXYZ)
        self.rays_grad = rays_grad   # train mode: rays carry requires_grad (pose / focal gradient through the static field)
        self.rays_grad_out = None
        self._g = None          # (inputs, graphs, outputs) once captured
        self.launches_per_step = None

    def _gather(self, out):
        from .parallel import gather_ray_maps

        return gather_ray_maps(out, self.world)

    @torch.no_grad()
    def forward_fused(self, rays, ts):
        cfg = self.dyn.native_cfg(self.rt)
        zvec = ops.make_zvec(cfg, self.s, self.rt, self.jitter, rays.device)
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

        xyz, z, valid = R.sampleXYZ(self.dyn, rays, self.s, self.rt, self.jitter)
        fs = self.static(rays, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        fd = self.dyn(rays, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        o = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, self.rt)
        return torch.cat((o[0], o[1][:, None], o[2][:, None], o[4], o[5][:, None], o[6][:, None], o[8], o[9][:, None],
                          o[10][:, None], o[12][:, None]), -1)

    # ------------------------------------------------------------------ eager pieces
    def _compute(self, rays, ts, target, fused):
        """Local (this rank's) part up to -- excluding -- the collectives and the optimizer update."""
        if self.mode == "fwd":
            return self.forward_fused(rays, ts) if fused else self.forward_api(rays, ts)
        from . import backward as B

        return B.train_fwd_bwd(self, rays, ts, target)

    def _finish(self):
        if self.mode == "train":
            from . import backward as B

            B.train_update(self)

    def _sync_grads(self):
        if self.mode == "train" and self.world > 1:
            from .parallel import allreduce_flat_mean

            allreduce_flat_mean(self.flat_grad, self.world)

    def _eager(self, rays, ts, target, fused):
        from . import lib as L

        n0 = L.launch_count
        out = self._compute(rays, ts, target, fused)
        self._sync_grads()
        self._finish()
        self.launches_per_step = L.launch_count - n0
        return out

    # ------------------------------------------------------------------ CUDA-graph path
    def _capture(self, rays, ts, target, fused):
        import gc

        gc.collect()   # drop autograd graphs (and their AccumulateGrad nodes, which remember a stream) of earlier eager steps
        # p.grad lives in the flat buffer created on the caller's stream while the warm-up / capture run on a side stream: the
        # mismatch is intended (both are joined before and after), so the per-backward warning about it is noise
        _warn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if _warn is not None:
            _warn(False)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        ins = (rays.clone(), ts.clone(), None if target is None else target.clone())
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager(*ins, fused)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        gc.collect()
        g1 = torch.cuda.CUDAGraph()
        g2 = None
        if self.world == 1 or self.mode == "fwd":
            with torch.cuda.graph(g1, stream=side):
                out = self._compute(*ins, fused)
                self._finish()
        else:
            with torch.cuda.graph(g1, stream=side):
                out = self._compute(*ins, fused)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool(), stream=side):
                self._finish()
        self._g = (ins, g1, g2, out, fused)
        if _warn is not None:
            _warn(True)   # restore the process-wide default

    def _replay(self, rays, ts, target):
        ins, g1, g2, out, _ = self._g
        ins[0].copy_(rays, non_blocking=True)
        ins[1].copy_(ts, non_blocking=True)
        if target is not None and ins[2] is not None:
            ins[2].copy_(target, non_blocking=True)
        g1.replay()
        if g2 is not None:
            self._sync_grads()
            g2.replay()
        if self.mode == "train":
            ops.PARAM_EPOCH[0] += 1   # weights moved on the device without their version counters changing
        return out

    def __call__(self, rays, ts, target, fused=True):
        """-> per-ray maps [N*world, 16].  With graph=True and world == 1 the returned tensor is the graph's static output
        buffer: the next call overwrites it (clone it to keep it)."""
        if self.mode == "train":
            fused = False   # the train step always goes through the public API
        if self.use_graph:
            if self._g is None or self._g[4] != fused:
                self._capture(rays, ts, target, fused)
            return self._gather(self._replay(rays, ts, target))
        return self._gather(self._eager(rays, ts, target, fused))
