"""The reference's FULL train step replayed through the drop-in API (BASELINE cfg 4): the nine hot-path passes of
train.py:1092-2292 with the same field-call / detach structure, joint pose + focal optimisation, synthetic targets.

  PASS 1  main rays (detached)            both fields, static outputs detached        train.py:1092-1162
  PASS 2  same rays, random timestamps    both fields, static outputs detached        train.py:1166-1246
  PASS 3/4  flow-displaced rays, t +- dt  both fields, only weights_d used            train.py:1433-1510, 1530-1607
  PASS 5  main rays WITH grad             static field (-> poses, focal); the dynamic field's outputs are not used by any
                                          loss of this pass (train.py:1812) -- it runs under no_grad here (its gradient is
                                          exactly zero in the reference, which spends a full dynamic backward on it)
  PASS 6/7  flow-displaced rays WITH grad static field only (weights, pts_ref)        train.py:1982-2000, 2049-2067
  PASS 8/9  +1-pixel neighbours           both fields (dynamic under no_grad, as PASS 5), only depth_map_s used  train.py:2163-2292
Loss terms restated from the same lines: photometric (full / dynamic / masked static), mask + skewed-mask + mask-L1, novel-time
order loss, induced-flow losses x (2 dynamic + 2 static) through renderer.induce_flow, small-scene-flow, disparity consistency on
the flow-displaced passes, normalised depth losses, patch depth smoothness on the neighbour passes, optional TV regularisers.
NOT restated: `flatten_eff_distloss` (third-party torch_efficient_distloss, absent here, unpinned -- SURVEY 8c) and the
tensorboard scalar logging (`.item()` host syncs).  One Adam over the fields, the scene-flow MLP, the 9-D poses and the focal.
"""
from __future__ import annotations

import gc
from typing import Dict

import torch

from . import extras, ops


def _depth_loss(pred, target):
    """compute_depth_loss (train.py:797-807): both maps shifted by their median and scaled by their mean abs deviation."""
    def nrm(x):
        t = torch.sort(x).values[(x.numel() - 1) // 2]      # == torch.median(x) (lower median); sort is CUDA-graph capturable
        s = torch.mean(torch.abs(x - t))
        return (x - t) / (s + 1e-8)

    return torch.mean((nrm(pred) - nrm(target)) ** 2)


def _skewed(mask):
    m = torch.clamp(mask, min=1e-6, max=1.0 - 1e-6) ** 2
    return torch.mean(-(m * torch.log(m) + (1 - m) * torch.log(1 - m)))


class Train9:
    def __init__(self, static, dyn, n_samples: int, ray_type: str, H: int, W: int, n_frames: int, focal: float, world: int = 1,
                 graph: bool = True, tv_weight: float = 0.0):
        self.static, self.dyn, self.s, self.rt = static, dyn, n_samples, ray_type
        self.H, self.W, self.F, self.world, self.use_graph, self.tv_weight = H, W, n_frames, world, graph, tv_weight
        dev = next(static.parameters()).device
        self.dev = dev
        eye = torch.eye(3, device=dev)[:2].reshape(1, 6).repeat(n_frames, 1)
        self.poses = torch.nn.Parameter(torch.cat((eye, torch.zeros(n_frames, 3, device=dev)), -1))     # poses_refine (9-D)
        self.focal = torch.nn.Parameter(torch.tensor([float(focal)], device=dev))                       # focal_refine
        self.params = list(static.parameters()) + list(dyn.parameters()) + [self.poses, self.focal]
        from . import optim

        self.flat_grad = optim.alias_grads(self.params)
        self.opt = optim.FlatAdam(self.params, self.flat_grad, lr=1e-4, betas=(0.9, 0.99))
        self._g = None
        self.launches_per_step = None
        self.t_interval = 2.0 / (n_frames - 1)

    # ------------------------------------------------------------------ one local step (no collectives, no update)
    def _fields(self, rays_s, rays_d, ts, dyn_grad=True, static_grad=True, static_only=False):
        """sampleXYZ -> static -> dynamic through the public API.  rays_s: rays for the sampler + static field (may carry grad);
        rays_d: detached rays for the dynamic field."""
        import renderer as R

        xyz, z, valid = R.sampleXYZ(self.dyn, rays_s, self.s, self.rt, True)
        if static_grad:
            fs = self.static(rays_s, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        else:
            with torch.no_grad():
                fs = self.static(rays_s, ts, None, xyz, z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        if static_only:
            return fs, None, z
        if dyn_grad:
            fd = self.dyn(rays_d, ts, None, xyz.detach(), z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        else:
            with torch.no_grad():
                fd = self.dyn(rays_d, ts, None, xyz.detach(), z, valid, is_train=True, ray_type=self.rt, N_samples=self.s)
        return fs, fd, z

    def _compute(self, b: Dict[str, torch.Tensor]):
        import renderer as R

        H, W, rt = self.H, self.W, self.rt
        ids, rgb, ts, ts_rand = b["ids"], b["rgb"], b["ts"], b["ts_rand"]
        flow_f, flow_b, mf, mb, fg, disp = b["flow_f"], b["flow_b"], b["mask_f"], b["mask_b"], b["fg"], b["disp"]
        self.flat_grad.zero_()
        focal = self.focal
        mtx = extras.pose_to_mtx(self.poses)
        mtx_f = torch.cat((mtx[1:], mtx[-1:]), 0)            # train.py:1081-1085
        mtx_b = torch.cat((mtx[0:1], mtx[:-1]), 0)
        col, row, view = ids % W, (ids // W) % H, ids // (W * H)
        rays = R.generate_rays(ids, None, focal, H, W, rt, poses_mtx=mtx)
        rays_det = rays.detach()
        pts2d = torch.stack((col.float() + 0.5, row.float() + 0.5), -1)     # grid_train
        pose_f, pose_b = mtx_f[view], mtx_b[view]
        total = 0.0

        # ---- PASS 1
        fs, fd, z = self._fields(rays_det, rays_det, ts, static_grad=False)
        o = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays_det, True, rt)
        rgb_full, rgb_s, depth_s, w_s, rgb_d, depth_d, w_d, dynm = o[0], o[4], o[5], o[7], o[8], o[9], o[11], o[12]
        pts_ref = fd[3]
        total = total + 3.0 * torch.mean((rgb_full - rgb) ** 2) + torch.mean((rgb_d - rgb) ** 2)
        total = total + 0.1 * torch.mean(torch.abs(dynm - fg[:, 0])) + 0.01 * _skewed(dynm) + 0.01 * torch.mean(torch.abs(dynm))
        sf_f, sf_b = self.dyn.get_forward_backward_scene_flow(pts_ref, ts)
        clamp = None if rt == "ndc" else (-2.0 + 1e-6, 2.0 - 1e-6)
        if_f, disp_f = R.induce_flow(H, W, focal.detach(), pose_f.detach(), w_d, pts_ref, pts2d, rays_det, rt, scene_flow=sf_f, clamp=clamp)
        if_b, disp_b = R.induce_flow(H, W, focal.detach(), pose_b.detach(), w_d, pts_ref, pts2d, rays_det, rt, scene_flow=sf_b, clamp=clamp)
        total = total + 0.02 * (torch.sum(torch.abs(if_f - flow_f) * mf) / (torch.sum(mf) + 1e-8) / 2)
        total = total + 0.02 * (torch.sum(torch.abs(if_b - flow_b) * mb) / (torch.sum(mb) + 1e-8) / 2)
        total = total + 0.01 * (torch.mean(torch.abs(sf_f)) + torch.mean(torch.abs(sf_b)))
        total = total + 0.04 * _depth_loss(depth_d, disp)

        # ---- PASS 2: same rays, random timestamps
        fs2, fd2, _ = self._fields(rays_det, rays_det, ts_rand, static_grad=False)
        o2 = R.raw2outputs(fs2[6], fs2[7], fd2[6], fd2[7], fd2[9], fd2[2], fd2[8], rays_det, True, rt)
        depth_s2, depth_d2, dyn2 = o2[5], o2[9], o2[12]
        total = total + 0.01 * _skewed(dyn2) + 0.01 * torch.mean(torch.abs(dyn2))
        total = total + 10.0 * torch.sum(((depth_d2 - depth_s2.detach()) ** 2) * (1.0 - dyn2.detach())) / (torch.sum(1.0 - dyn2.detach()) + 1e-8)

        # ---- PASS 3 / 4: flow-displaced rays at t +- dt, only weights_d used (disparity consistency)
        for flow, pose_n, mtx_n, dt, disp_1, sf_back in ((flow_f, pose_f, mtx_f, self.t_interval, disp_f, 1), (flow_b, pose_b, mtx_b, -self.t_interval, disp_b, 0)):
            uv = pts2d + flow
            rays_n = R.generate_rays(None, None, focal.detach(), H, W, rt, uv=uv, view_ids=view, poses_mtx=mtx_n.detach()).detach()
            fs3, fd3, _ = self._fields(rays_n, rays_n, ts + dt, static_grad=False)
            o3 = R.raw2outputs(fs3[6], fs3[7], fd3[6], fd3[7], fd3[9], fd3[2], fd3[8], rays_n, True, rt)
            w_d3 = o3[11]
            sf3 = self.dyn.get_forward_backward_scene_flow(fd3[3], ts + dt)[sf_back]
            _, disp_2 = R.induce_flow(H, W, focal.detach(), pose_n.detach(), w_d3, fd3[3], uv, rays_n, rt, scene_flow=sf3, clamp=clamp)
            total = total + 0.04 * torch.mean(torch.abs(disp_2 - disp_1))

        # ---- PASS 5: rays WITH grad -> static field (pose / focal); dynamic field unused by the losses -> no_grad
        fs5, fd5, _ = self._fields(rays, rays_det, ts, dyn_grad=False)
        o5 = R.raw2outputs(fs5[6], fs5[7], fd5[6], fd5[7], fd5[9], fd5[2], fd5[8], rays_det, True, rt)
        rgb_s5, depth_s5, w_s5 = o5[4], o5[5], o5[7]
        bgm = 1.0 - fg
        total = total + torch.sum((rgb_s5 - rgb) ** 2 * bgm) / (torch.sum(bgm) + 1e-8) / 3
        total = total + 0.04 * _depth_loss(depth_s5, disp)
        if_fs, disp_fs = R.induce_flow(H, W, focal, pose_f, w_s5, fs5[3], pts2d, rays, rt)
        if_bs, disp_bs = R.induce_flow(H, W, focal, pose_b, w_s5, fs5[3], pts2d, rays, rt)
        total = total + 0.02 * (torch.sum(torch.abs(if_fs - flow_f) * mf * bgm) / (torch.sum(mf * bgm) + 1e-8) / 2)
        total = total + 0.02 * (torch.sum(torch.abs(if_bs - flow_b) * mb * bgm) / (torch.sum(mb * bgm) + 1e-8) / 2)

        # ---- PASS 6 / 7: static only on the flow-displaced rays, WITH grad to poses / focal
        for flow, pose_n, mtx_n, disp_1 in ((flow_f, pose_f, mtx_f, disp_fs), (flow_b, pose_b, mtx_b, disp_bs)):
            uv = pts2d + flow
            rays_n = R.generate_rays(None, None, focal, H, W, rt, uv=uv, view_ids=view, poses_mtx=mtx_n)
            fs6, _, _ = self._fields(rays_n, None, ts, static_only=True)
            _, disp_2 = R.induce_flow(H, W, focal, pose_n, fs6[4], fs6[3], uv, rays_n, rt)
            total = total + 0.04 * torch.sum(torch.abs(disp_2 - disp_1) * bgm) / (torch.sum(bgm) + 1e-8)

        # ---- PASS 8 / 9: +1-pixel neighbours, only depth_map_s used (patch smoothness, train.py:2120-2292)
        for dcol, drow in ((1, 0), (0, 1)):
            uvn = torch.stack((torch.clamp(col + dcol, max=W - 1).float() + 0.5, torch.clamp(row + drow, max=H - 1).float() + 0.5), -1)
            rays_n = R.generate_rays(None, None, focal, H, W, rt, uv=uvn, view_ids=view, poses_mtx=mtx)
            fs8, fd8, _ = self._fields(rays_n, rays_n.detach(), ts, dyn_grad=False)
            o8 = R.raw2outputs(fs8[6], fs8[7], fd8[6], fd8[7], fd8[9], fd8[2], fd8[8], rays_n.detach(), True, rt)
            total = total + 0.01 * torch.mean(torch.abs(o8[5] - depth_s5) * bgm[:, 0])

        if self.tv_weight > 0:
            reg = _TV(1.0)
            total = total + self.tv_weight * (self.dyn.TV_loss_density(reg) + self.dyn.TV_loss_blending(reg) + self.dyn.TV_loss_app(reg)
                                              + self.static.TV_loss_density(reg) + self.static.TV_loss_app(reg))
        from . import backward as B

        B.DIRECT_PLANE_GRADS[0] = True        # p.grad are live views of the (zeroed) flat buffer
        try:
            total.backward()
        finally:
            B.DIRECT_PLANE_GRADS[0] = False
        return torch.cat((rgb_full, depth_d[:, None], dynm[:, None]), -1).detach()

    # ------------------------------------------------------------------ eager / graph drivers
    def _sync(self):
        if self.world > 1:
            from .parallel import allreduce_flat_mean

            allreduce_flat_mean(self.flat_grad, self.world)

    def _eager(self, b):
        from . import lib as L

        n0 = L.launch_count
        out = self._compute(b)
        self._sync()
        self.opt.step()
        self.launches_per_step = L.launch_count - n0
        return out

    def _capture(self, b):
        gc.collect()
        _warn = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if _warn is not None:
            _warn(False)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        ins = {k: v.clone() for k, v in b.items()}
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(3):
                self._eager(ins)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        gc.collect()
        g1, g2 = torch.cuda.CUDAGraph(), None
        if self.world == 1:
            with torch.cuda.graph(g1, stream=side):
                out = self._compute(ins)
                self.opt.step()
        else:
            with torch.cuda.graph(g1, stream=side):
                out = self._compute(ins)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool(), stream=side):
                self.opt.step()
        self._g = (ins, g1, g2, out)
        if _warn is not None:
            _warn(True)

    def __call__(self, b: Dict[str, torch.Tensor]):
        if not self.use_graph:
            return self._eager(b)
        if self._g is None:
            self._capture(b)
        ins, g1, g2, out = self._g
        for k, v in b.items():
            ins[k].copy_(v, non_blocking=True)
        g1.replay()
        if g2 is not None:
            self._sync()
            g2.replay()
        ops.PARAM_EPOCH[0] += 1
        return out


class _TV:
    def __init__(self, w):
        self.TVLoss_weight = w


def synth_batch(n_rays: int, H: int, W: int, n_frames: int, gen: torch.Generator, dev) -> Dict[str, torch.Tensor]:
    """Synthetic targets with the shapes train.py:1045-1060 gathers per iteration (no dataset exists here)."""
    ids = torch.randint(0, n_frames * H * W, (n_rays,), generator=gen)
    t_ref = ids // (H * W)
    ts = t_ref.float() / (n_frames - 1) * 2.0 - 1.0
    b = dict(ids=ids, rgb=torch.rand(n_rays, 3, generator=gen), ts=ts, ts_rand=torch.rand(n_rays, generator=gen) * 2 - 1,
             flow_f=torch.randn(n_rays, 2, generator=gen), flow_b=torch.randn(n_rays, 2, generator=gen),
             mask_f=(torch.rand(n_rays, 1, generator=gen) > 0.2).float(), mask_b=(torch.rand(n_rays, 1, generator=gen) > 0.2).float(),
             fg=(torch.rand(n_rays, 1, generator=gen) > 0.7).float(), disp=torch.rand(n_rays, generator=gen))
    return {k: v.to(dev) for k, v in b.items()}
