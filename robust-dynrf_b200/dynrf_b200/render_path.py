"""Frame-sharded novel-view rendering (SURVEY.md 8f row f4 / BASELINE cfg 5): the chunk loop of
renderer.evaluation_path (reference renderer.py:970-1263: per pose -> all H*W rays -> 8192-ray chunks -> sampleXYZ ->
static field -> dynamic field -> raw2outputs -> cat) driven on the device:

  * rays of a frame are generated on the GPU by the ray-generation kernel (f1; the reference builds them on the CPU with
    get_ray_directions_blender / get_rays / ndc_rays_blender, renderer.py:1021-1032, and copies every chunk H2D);
  * every chunk is one CUDA-graph replay of the no-grad forward (samples generated in-kernel for the static density,
    active-sample lists, tensor-core MLP tiles, two-field compositing);
  * the 16 per-ray maps of a finished frame go to pinned host memory on a copy stream while the next frame renders
    (double buffered); image encoding / video writing (renderer.py:1140-1250) is the caller's business;
  * frames are independent: rank r renders frames r, r + world, ... (no collective in the loop).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch

from . import extras
from .train_step import Step


def change_view_time_poses(c2w: np.ndarray, max_trans: float, length: int) -> np.ndarray:
    """The `change_view_time` camera path of train.generate_path (reference train.py:282-317): a circular translation of period
    30 frames around the reference pose.  c2w: [3,4] -> [length,3,4]."""
    ref = np.concatenate([c2w[:3, :4], np.array([[0.0, 0.0, 0.0, 1.0]])], 0)
    out = []
    for i in range(length):
        x = max_trans * 1.5 * np.sin(2.0 * np.pi * float(i) / 30.0) * 2.0
        y = max_trans * 1.5 * (np.cos(2.0 * np.pi * float(i) / 30.0) - 1.0) * 2.0 / 3.0
        ip = np.eye(4)
        ip[:3, 3] = [x, y, 0.0]
        out.append(np.dot(ref, np.linalg.inv(ip))[:3, :])
    return np.stack(out, 0)


def frame_time(idx: int, n_frames: int) -> float:
    """`change_time == "change"` branch of evaluation_path (renderer.py:1034-1040): discrete time in [-1, 1]."""
    return round(idx / (n_frames - 1) * (n_frames - 1)) / (n_frames - 1) * 2.0 - 1.0


class RenderDriver:
    """render(frames) -> list of pinned host tensors [H*W,16] (rgb_full3 depth_full acc_full rgb_s3 depth_s acc_s rgb_d3 depth_d
    acc_d dynamicness), one per frame of this rank."""

    def __init__(self, static, dyn, H: int, W: int, focal: float, n_samples: int, ray_type: str = "ndc", chunk: int = 8192,
                 rank: int = 0, world: int = 1, graph: bool = True):
        self.static, self.dyn, self.H, self.W, self.focal = static, dyn, H, W, float(focal)
        self.s, self.rt, self.chunk, self.rank, self.world = n_samples, ray_type, chunk, rank, world
        self.dev = next(static.parameters()).device
        n = H * W
        self.n_full, self.tail = n // chunk, n % chunk
        self.step_full = Step(static, dyn, n_samples, ray_type, "fwd", 1, graph=graph, jitter=False)
        self.step_tail = Step(static, dyn, n_samples, ray_type, "fwd", 1, graph=graph, jitter=False) if self.tail else None
        self.ids = torch.arange(n, device=self.dev, dtype=torch.int64)
        self.focal_t = torch.tensor([self.focal], device=self.dev)
        self.frame = [torch.empty(n, 16, device=self.dev) for _ in range(2)]
        self.host = [torch.empty(n, 16, pin_memory=True) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.launches = 0

    def my_frames(self, n_frames: int) -> List[int]:
        return list(range(self.rank, n_frames, self.world))

    @torch.no_grad()
    def render_frame(self, c2w: torch.Tensor, t: float, slot: int) -> torch.Tensor:
        """c2w [3,4] (device) -> device frame buffer [H*W,16] (slot 0/1)."""
        n = self.H * self.W
        rays = extras._RayGenFn.apply(self.ids, c2w.reshape(1, 3, 4).contiguous(), self.focal_t, self.H, self.W, self.rt == "ndc", 1.0, None, None)
        ts = torch.full((self.chunk,), float(t), device=self.dev)
        buf = self.frame[slot]
        for c in range(self.n_full):
            lo = c * self.chunk
            buf[lo:lo + self.chunk].copy_(self.step_full(rays[lo:lo + self.chunk], ts, None, fused=True))
        if self.tail:
            lo = self.n_full * self.chunk
            buf[lo:].copy_(self.step_tail(rays[lo:], ts[: self.tail], None, fused=True))
        # a graph replay launches the kernels its eager warm-up step counted (Step.launches_per_step); + 1 ray-generation kernel
        self.launches += 1 + self.n_full * (self.step_full.launches_per_step or 0) + ((self.step_tail.launches_per_step or 0) if self.tail else 0)
        return buf

    def render(self, poses: np.ndarray, times: Optional[List[float]] = None, frames: Optional[List[int]] = None):
        """poses [F,3,4]; renders this rank's frames (or `frames`) and returns [(frame index, pinned host maps)]."""
        n_frames = poses.shape[0]
        frames = self.my_frames(n_frames) if frames is None else frames
        out = []
        cur = torch.cuda.current_stream(self.dev)
        for k, idx in enumerate(frames):
            slot = k & 1
            self.done[slot].synchronize()                   # the D2H copy that last used this slot has finished
            t = frame_time(idx, n_frames) if times is None else times[idx]
            buf = self.render_frame(torch.as_tensor(poses[idx], dtype=torch.float32, device=self.dev), t, slot)
            self.copy_stream.wait_stream(cur)
            with torch.cuda.stream(self.copy_stream):
                self.host[slot].copy_(buf, non_blocking=True)
                self.done[slot].record(self.copy_stream)
            out.append((idx, self.host[slot]))
        for e in self.done:
            e.synchronize()
        return out
