"""Torch-facing wrappers around the C ABI: allocate outputs with torch, pass raw device
pointers + the current CUDA stream to libdynrf_b200.so.  PyTorch is plumbing here (memory,
streams, autograd bookkeeping); every FLOP of the hot path runs in the library.

Reference functions replaced (paths relative to the reference checkout):
  sample_xyz            renderer.py:147-170, models/tensorBase.py:487-499,524-559
  static_field_forward  models/tensorBase.py:704-825 + models/tensoRF.py:118-196
  dynamic_field_forward models/tensorBase.py:704-850 + models/tensoRF.py:521-811
  composite             renderer.py:173-315
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "dynrf_b200: the ray-march hot path runs only on a CUDA device (sm_100a); got a "
                f"{t.device} tensor. There is no CPU fallback.")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------- parameter layout


def to_channel_last_(p: torch.Tensor) -> None:
    """Make a (1,C,H,W) / (1,C,L,1) parameter physically [H][W][C] while keeping its logical shape
    (state_dict shape, Adam state and checkpoints are unaffected -- SURVEY.md 7, hard part 1)."""
    _, c, h, w = p.shape
    want = (c * h * w, 1, w * c, c)
    s = p.stride()
    ok = (c == 1 or s[1] == 1) and (h == 1 or s[2] == want[2]) and (w == 1 or s[3] == want[3])
    if not ok:
        with torch.no_grad():
            p.data = p.data.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def is_hfast(p: torch.Tensor) -> bool:
    """True if a (1,C,H,W) tensor is physically [W][H][C] (H fastest among the two texel dimensions)."""
    _, c, h, w = p.shape
    s = p.stride()
    return (c == 1 or s[1] == 1) and (h == 1 or s[2] == c) and (w == 1 or s[3] == h * c)


# "zfast" (default) or "xfast" (every plane [H][W][C], the round-1 layout; kept selectable for A/B measurements)
PLANE_LAYOUT = os.environ.get("DYNRF_PLANE_LAYOUT", "zfast")


def to_plane_layout_(p: torch.Tensor, plane_index: int) -> None:
    """Native layout of VM plane `plane_index` (logical shape (1,C,H,W) unchanged):
      plane 0 (x, y)           : [H][W][C]  -- the samples of a ray share their (x, y) texel, nothing to coalesce;
      planes 1, 2 (x|y, z = H) : [W][H][C]  -- z fastest: consecutive samples of a ray (= consecutive lanes of a warp) advance
                                  mostly in z, so their taps fall into a few contiguous 128-byte lines instead of one row each."""
    if plane_index == 0 or PLANE_LAYOUT == "xfast":
        to_channel_last_(p)
    elif not is_hfast(p):
        with torch.no_grad():
            p.data = p.data.permute(0, 3, 2, 1).contiguous().permute(0, 3, 2, 1)


_VM_CACHE: Dict[tuple, "L.VmT"] = {}


def _vm_key(planes, lines):
    return tuple((t.data_ptr(), t.shape[2], t.shape[3], t.stride(1), t.stride(2), t.stride(3)) for t in (*planes, *lines))


def vm_struct(planes: Sequence[torch.Tensor], lines: Sequence[torch.Tensor]) -> L.VmT:
    """dynrf_vm_t for a plane/line set; cached on (data_ptr, shape, strides) of the six tensors (pointers only change when a
    parameter is replaced, e.g. by upsample_volume_grid; values may change in place)."""
    key = _vm_key(planes, lines)
    hit = _VM_CACHE.get(key)
    if hit is not None:
        return hit
    if len(_VM_CACHE) > 256:
        _VM_CACHE.clear()
    vm = _vm_struct_build(planes, lines)
    # re-key after a possible in-place layout conversion moved the storage
    _VM_CACHE[_vm_key(planes, lines)] = vm
    return vm


def _vm_struct_build(planes: Sequence[torch.Tensor], lines: Sequence[torch.Tensor]) -> L.VmT:
    vm = L.VmT()
    for i in range(3):
        to_plane_layout_(planes[i], i)
        to_channel_last_(lines[i])
        _, c, h, w = planes[i].shape
        vm.plane[i] = planes[i].data_ptr()
        vm.line[i] = lines[i].data_ptr()
        vm.C[i], vm.H[i], vm.W[i] = c, h, w
        vm.L[i] = lines[i].shape[2]
        vm.sv[i], vm.su[i] = (1, h) if (i > 0 and PLANE_LAYOUT != "xfast") else (w, 1)
        if lines[i].shape[1] != c:
            raise ValueError("plane/line component counts differ")
    return vm


def vm_grad_struct(gplanes: Sequence[torch.Tensor], glines: Sequence[torch.Tensor]) -> L.VmGradT:
    g = L.VmGradT()
    for i in range(3):
        g.plane[i] = gplanes[i].data_ptr()
        g.line[i] = glines[i].data_ptr()
    return g


def zeros_like_cl(p: torch.Tensor) -> torch.Tensor:
    """zeros with the same (channel-last) strides as parameter p."""
    return torch.zeros_like(p, memory_format=torch.preserve_format)


class PackedLayer:
    """nn.Linear weight packed to the library's [Kpad][Npad] input-major layout; repacked when the
    parameter's version counter moves (optimizer steps are in-place)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        self.weight, self.bias = weight, bias
        lib = L.load()
        self.N, self.K = weight.shape
        self.Kpad, self.Npad = lib.dynrf_pad_k(self.K), lib.dynrf_pad_n(self.N)
        self.wt = torch.empty(self.Kpad * self.Npad, device=weight.device, dtype=torch.float32)
        self.b = torch.empty(self.Npad, device=weight.device, dtype=torch.float32)
        self._key = None
        self.struct = L.LayerT(self.wt.data_ptr(), self.b.data_ptr(), self.K, self.N, self.Kpad, self.Npad)

    def refresh(self) -> L.LayerT:
        key = (PARAM_EPOCH[0], self.weight.data_ptr(), self.weight._version, None if self.bias is None else self.bias._version)
        if key != self._key:
            w = _f32c(self.weight.detach())
            b = None if self.bias is None else _f32c(self.bias.detach())
            L.check(L.load().dynrf_pack_linear(_p(w), _p(b), self.N, self.K, _p(self.wt), _p(self.b), _stream()),
                    "dynrf_pack_linear")
            self._key = key
        return self.struct


class PackedUmmaLayer:
    """nn.Linear weight pre-split (tf32 hi/lo) and laid out for the tcgen05 path; repacked on version change."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor]):
        self.weight, self.bias = weight, bias
        self.N, self.K = weight.shape
        self.Npad = (self.N + 7) // 8 * 8
        self.nchunks = (self.K + 31) // 32
        self.wq = torch.empty(self.nchunks * 2 * self.Npad * 32, device=weight.device, dtype=torch.float32)
        self.b = torch.empty(self.Npad, device=weight.device, dtype=torch.float32)
        self._key = None
        self.struct = L.UmmaLayerT(self.wq.data_ptr(), self.b.data_ptr(), self.K, self.N, self.Npad, self.nchunks)

    def refresh(self) -> L.UmmaLayerT:
        key = (PARAM_EPOCH[0], self.weight.data_ptr(), self.weight._version, None if self.bias is None else self.bias._version)
        if key != self._key:
            w = _f32c(self.weight.detach())
            b = None if self.bias is None else _f32c(self.bias.detach())
            L.check(L.load().dynrf_pack_linear_umma(_p(w), _p(b), self.N, self.K, _p(self.wq), _p(self.b), _stream()),
                    "dynrf_pack_linear_umma")
            self._key = key
        return self.struct


# tensor-core (tcgen05) kernels for the MLP stages that have them; the SIMT fp32 kernels stay in the library
# as a cross-check (tests flip this flag to compare the two)
USE_TC = True
# Bumped whenever parameters may have changed without their autograd version counters moving (a CUDA-graph replay of a
# train step updates the weights in place on the device): every packed-operand cache includes it in its key.
PARAM_EPOCH = [0]
# keep layer inputs / hidden activations of the tensor-core forward kernels for their backward instead of recomputing them
SAVE_ACTIVATIONS = True
TC_SB_VERSION = 2
TC_WARP_VERSION = int(os.environ.get("DYNRF_TC_WARP", "2"))   # 2: warp-specialised two-tile kernel (tc_warp_fwd.cu); 1: first-generation k_warp_tc (kept as a cross-check)


def umma_layers(m, names):
    cache = m.__dict__.setdefault("_packed_umma", {})
    tbl = m._layer_table()
    out = {}
    for n in names:
        lin = tbl[n]
        pl = cache.get(n)
        if pl is None or pl.weight is not lin.weight or pl.bias is not lin.bias:
            pl = cache[n] = PackedUmmaLayer(lin.weight, lin.bias)
        out[n] = pl.refresh()
    return out


class PackedSigmaBlend:
    """Operands of dynrf_dyn_sigma_blend_fwd_tc2, rebuilt when any of the four density / blending layers changes."""

    def __init__(self, m):
        self.m = m
        dev = m.density_layer1.weight.device
        self.wq = torch.empty(L.load().dynrf_sb_umma_floats(), device=dev)
        self.b1cat = torch.empty(128, device=dev)
        self.w2cat = torch.empty(128, device=dev)
        self.b2cat = torch.empty(2, device=dev)
        self._key = None

    def refresh(self):
        m = self.m
        ps = [m.density_layer1.weight, m.density_layer1.bias, m.density_layer2.weight, m.density_layer2.bias,
              m.blending_layer1.weight, m.blending_layer1.bias, m.blending_layer2.weight, m.blending_layer2.bias]
        key = (PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._key:
            with torch.no_grad():
                wd, wb = _f32c(ps[0].detach()), _f32c(ps[4].detach())
                L.check(L.load().dynrf_pack_sb_umma(_p(wd), _p(wb), _p(self.wq), _stream()), "dynrf_pack_sb_umma")
                torch.cat((ps[1].detach(), ps[5].detach()), out=self.b1cat)
                torch.cat((ps[2].detach().reshape(-1), ps[6].detach().reshape(-1)), out=self.w2cat)
                torch.cat((ps[3].detach().reshape(-1), ps[7].detach().reshape(-1)), out=self.b2cat)
            self._key = key
        return self


class PackedWarpFwd:
    """Operands of dynrf_warp_fwd_tc2 (coordinate-warp MLP layer3 / layer4 / layer5), rebuilt when one of the layers changes."""

    def __init__(self, m):
        self.m = m
        lib = L.load()
        dev = m.layer3.weight.device
        if tuple(m.layer3.weight.shape) != (64, 93) or tuple(m.layer4.weight.shape) != (64, 64) or tuple(m.layer5.weight.shape) != (3, 64):
            raise NotImplementedError("tensor-core warp MLP is compiled for 93 -> 64 -> 64 -> 3 (models/tensoRF.py:521-541); got "
                                      f"{tuple(m.layer3.weight.shape)}, {tuple(m.layer4.weight.shape)}, {tuple(m.layer5.weight.shape)}")
        self.wq = torch.empty(lib.dynrf_warp_fwd_wfloats(), device=dev)
        self.cst = torch.empty(lib.dynrf_warp_fwd_cst_floats(), device=dev)
        self._key = None

    def refresh(self):
        m = self.m
        ps = [m.layer3.weight, m.layer3.bias, m.layer4.weight, m.layer4.bias, m.layer5.weight, m.layer5.bias]
        key = (PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._key:
            with torch.no_grad():
                L.check(L.load().dynrf_pack_warp_fwd(_p(_f32c(ps[0].detach())), _p(_f32c(ps[2].detach())), _p(self.wq), _stream()),
                        "dynrf_pack_warp_fwd")
                torch.cat((ps[1].detach(), ps[3].detach(), ps[4].detach().t().reshape(-1), ps[5].detach()), out=self.cst)   # b3 | b4 | W5 [64][3] | b5
            self._key = key
        return self


_HEAD_K1 = {L.HEAD_FEA: 138, L.HEAD_FEA_TIMEEMB: 135, L.HEAD_FEA_LATE_VIEW: 107}


def check_tc_shapes(m) -> None:
    """The tensor-core kernels are compiled for the shapes every shipped config uses (density / blending C=[16,4,4],
    appearance C=[48,12,12], app_dim 27, featureC 128, view_pe 0, fea_pe 2 static / 0 dynamic).  The pack entry points take raw
    pointers, so anything else must fail HERE, loudly, instead of reading weights with the wrong row stride."""
    tbl = m._layer_table()
    l1, l2, l3 = tbl["h1"], tbl["h2"], tbl["h3"]
    k1 = _HEAD_K1[m.head_mode]
    nstr = 3 if m.is_dynamic else 1
    want = {
        "appearance_n_comp": (list(m.app_n_comp), [48, 12, 12]),
        "density_n_comp": (list(m.density_n_comp) if m.is_dynamic else [16, 4, 4], [16, 4, 4]),   # static density is shape-generic
        "basis_mat.weight": (tuple(m.basis_mat.weight.shape), (27, 72 * nstr)),
        "renderModule layer 1": (tuple(l1.weight.shape), (128, k1)),
        "renderModule layer 2": (tuple(l2.weight.shape), (128, 128)),
        "renderModule output layer": (tuple(l3.weight.shape), (3, 128 if m.head_mode == L.HEAD_FEA else 131)),
    }
    bad = {k: v for k, v in want.items() if v[0] != v[1]}
    if bad:
        raise NotImplementedError(
            "dynrf_b200: the tensor-core kernels support the hyper-parameters of the shipped configs only "
            "(app_dim=27, featureC=128, view_pe=0, fea_pe=2 static / 0 dynamic, C=[16,4,4]/[48,12,12]); got "
            + ", ".join(f"{k}={v[0]} (need {v[1]})" for k, v in bad.items())
            + ".  Set dynrf_b200.ops.USE_TC = False to run the shape-generic fp32 SIMT kernels.")


class PackedAppFwd:
    """Operands of dynrf_app_fwd_tc for one field: packed split-TF32 weights (basis_mat, head layers 1-2) and the
    constant block [b1 | b2 | W3[:, :128]^T | W3[:, 128:131]^T (zeros without a late view input) | b3]."""

    def __init__(self, m):
        self.m = m
        lib = L.load()
        check_tc_shapes(m)
        dev = m.basis_mat.weight.device
        self.dyn = 1 if m.is_dynamic else 0
        self.wq = torch.empty(lib.dynrf_app_fwd_wfloats(self.dyn), device=dev)
        self.cst = torch.empty(lib.dynrf_app_fwd_cst_floats(), device=dev)
        self._key = None

    def refresh(self):
        m = self.m
        tbl = m._layer_table()
        l1, l2, l3 = tbl["h1"], tbl["h2"], tbl["h3"]
        ps = [m.basis_mat.weight, l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias]
        key = (PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._key:
            check_tc_shapes(m)
            with torch.no_grad():
                L.check(L.load().dynrf_pack_app_fwd(self.dyn, m.head_mode, _p(_f32c(ps[0].detach())), _p(_f32c(ps[1].detach())),
                                                    _p(_f32c(ps[3].detach())), _p(self.wq), _stream()), "dynrf_pack_app_fwd")
                w3 = l3.weight.detach()
                w3v = w3[:, 128:131].t().reshape(-1) if w3.shape[1] == 131 else torch.zeros(9, device=w3.device)
                torch.cat((l1.bias.detach(), l2.bias.detach(), w3[:, :128].t().reshape(-1), w3v, l3.bias.detach()), out=self.cst)
            self._key = key
        return self


def packed_app_fwd(m) -> PackedAppFwd:
    p = m.__dict__.get("_packed_app_fwd")
    if p is None:
        p = m.__dict__["_packed_app_fwd"] = PackedAppFwd(m)
    return p.refresh()


def field_cfg(aabb: torch.Tensor, near_far, distance_scale, weight_thres, act: str, density_shift, ray_type: str) -> L.FieldCfgT:
    a = aabb.detach().cpu().tolist()
    cfg = L.FieldCfgT()
    for i in range(3):
        cfg.aabb_min[i] = a[0][i]
        cfg.aabb_max[i] = a[1][i]
    cfg.near_, cfg.far_ = float(near_far[0]), float(near_far[1])
    cfg.distance_scale = float(distance_scale)
    cfg.weight_thres = float(weight_thres)
    cfg.density_shift = float(density_shift)
    cfg.act = {"relu": L.ACT_RELU, "softplus": L.ACT_SOFTPLUS}[act]
    cfg.ray_type = {"ndc": L.RAY_NDC, "contract": L.RAY_CONTRACT}[ray_type]
    return cfg


# ----------------------------------------------------------------------------- a1-a3 sampling


_ZVEC_EVAL: Dict[tuple, torch.Tensor] = {}


def make_zvec(cfg: L.FieldCfgT, n_samples: int, ray_type: str, is_train: bool, device) -> torch.Tensor:
    """The batch-wide depth vector (models/tensorBase.py:490-492, 530-548).  Evaluation mode is deterministic, so its vector is
    computed once per (ray type, near, far, S, device) -- one launch less on every no-grad pass; callers never write to it."""
    lib = L.load()
    key = None
    if not is_train and not torch.cuda.is_current_stream_capturing():
        key = (ray_type, float(cfg.near_), float(cfg.far_), int(n_samples), str(device))
        hit = _ZVEC_EVAL.get(key)
        if hit is not None:
            return hit
    zvec = torch.empty(n_samples, device=device, dtype=torch.float32)
    if ray_type == "ndc":
        jit = torch.rand(n_samples, device=device) if is_train else None  # one draw shared by all rays
        L.check(lib.dynrf_zvec_ndc(cfg.near_, cfg.far_, n_samples, _p(jit), _p(zvec), _stream()), "dynrf_zvec_ndc")
    elif ray_type == "contract":
        n_in, n_out = n_samples - n_samples // 2, n_samples // 2
        ji = torch.rand(n_in + 1, device=device) if is_train else None
        jo = torch.rand(n_out + 1, device=device) if is_train else None
        L.check(lib.dynrf_zvec_contract(cfg.near_, cfg.far_, n_samples, _p(ji), _p(jo), _p(zvec), _stream()),
                "dynrf_zvec_contract")
    else:
        raise NotImplementedError(
            "ray_type must be 'ndc' or 'contract' (the aabb-marching sampler is TensorBase.sample_ray; renderer.sampleXYZ's own "
            "dispatch to it tiles the per-ray depths N times, renderer.py:168, and no config uses it)")
    if key is not None:
        if len(_ZVEC_EVAL) > 64:
            _ZVEC_EVAL.clear()
        _ZVEC_EVAL[key] = zvec
    return zvec


def sample_points(cfg: L.FieldCfgT, rays: torch.Tensor, zvec: torch.Tensor):
    _require_cuda(rays)
    rays = _f32c(rays)
    n, s = rays.shape[0], zvec.shape[0]
    xyz = torch.empty(n, s, 3, device=rays.device, dtype=torch.float32)
    z = torch.empty(n, s, device=rays.device, dtype=torch.float32)
    valid = torch.empty(n, s, device=rays.device, dtype=torch.bool)
    L.check(L.load().dynrf_sample_points(C.byref(cfg), _p(rays), n, s, _p(zvec), _p(xyz), _p(z), _p(valid), _stream()),
            "dynrf_sample_points")
    return xyz, z, valid


# ----------------------------------------------------------------------------- field forwards (no autograd)


def _active_buffers(n, s, device):
    idx = torch.empty(n * s, device=device, dtype=torch.int32)
    cnt = torch.zeros(1, device=device, dtype=torch.int32)
    return idx, cnt


def static_forward_raw(m, rays, xyz, z_vals, valid, ray_type, zvec=None):
    """-> dict(sigma, weight, dists, rgb, idx, cnt).  xyz=None generates the samples in-kernel from zvec."""
    lib = L.load()
    _require_cuda(rays)
    rays = _f32c(rays)
    n = rays.shape[0]
    s = zvec.shape[0] if xyz is None else xyz.shape[1]
    cfg = m.native_cfg(ray_type)
    dev = rays.device
    sigma = torch.empty(n, s, device=dev)
    weight = torch.empty(n, s, device=dev)
    dists = torch.empty(n, s, device=dev)
    rgb = torch.empty(n, s, 3, device=dev)
    idx, cnt = _active_buffers(n, s, dev)
    vden = vm_struct(m.density_plane, m.density_line)
    if xyz is not None:
        xyz, z_vals = _f32c(xyz), _f32c(z_vals)
        valid_u8 = valid.contiguous().view(torch.uint8)
    else:
        valid_u8 = None
    L.check(lib.dynrf_static_sigma_weight_fwd(C.byref(cfg), C.byref(vden), _p(rays), n, s, _p(xyz), _p(z_vals), _p(valid_u8),
                                              _p(zvec), _p(sigma), _p(weight), _p(dists), _p(idx), _p(cnt), _p(rgb), _stream()),
            "dynrf_static_sigma_weight_fwd")
    return dict(sigma=sigma, weight=weight, dists=dists, rgb=rgb, idx=idx, cnt=cnt, rays=rays, xyz=xyz, z_vals=z_vals,
                valid=valid_u8)


def static_app_raw(m, st, ray_type):
    lib = L.load()
    cfg = m.native_cfg(ray_type)
    n, s = st["sigma"].shape
    vapp = vm_struct(m.app_plane, m.app_line)
    if USE_TC:
        pa = packed_app_fwd(m)
        L.check(lib.dynrf_app_fwd_tc(C.byref(cfg), 0, C.byref(vapp), _p(pa.wq), _p(pa.cst), _p(st["rays"]), _p(st["xyz"]), _p(st["xyz"]),
                                     None, n, s, _p(st["idx"]), _p(st["cnt"]), _p(st["rgb"]), _p(st.get("featT")), st.get("ldf", 0),
                                     _stream()), "dynrf_app_fwd_tc")
        return
    ly = m.packed_layers()
    L.check(lib.dynrf_static_app_fwd(C.byref(cfg), C.byref(vapp), C.byref(ly["basis"]), C.byref(ly["h1"]), C.byref(ly["h2"]),
                                     C.byref(ly["h3"]), m.head_mode, _p(st["rays"]), _p(st["xyz"]), n, s, _p(st["idx"]),
                                     _p(st["cnt"]), _p(st["rgb"]), _stream()),
            "dynrf_static_app_fwd")


def round_up(x, k):
    return (x + k - 1) // k * k


def dynamic_forward_raw(m, rays, ts, xyz, z_vals, valid, ray_type, save_feat=False):
    lib = L.load()
    _require_cuda(rays, ts, xyz)
    rays, ts, xyz, z_vals = _f32c(rays), _f32c(ts), _f32c(xyz), _f32c(z_vals)
    valid_u8 = valid.contiguous().view(torch.uint8)
    n, s = z_vals.shape
    dev = rays.device
    cfg = m.native_cfg(ray_type)
    ly = m.packed_layers()
    tfeat = torch.empty(n, L.TFEAT, device=dev)
    L.check(lib.dynrf_time_embed_fwd(_p(ts), n, C.byref(ly["layer1"]), C.byref(ly["layer2"]), _p(tfeat), _stream()),
            "dynrf_time_embed_fwd")
    xyz_prime = torch.empty(n, s, 3, device=dev)
    sigma = torch.empty(n, s, device=dev)
    blend = torch.empty(n, s, device=dev)
    vden = vm_struct(m.density_plane, m.density_line)
    vbl = vm_struct(m.blending_plane, m.blending_line)
    sb_saved, sb_ld = None, 0
    warp_saved, warp_ld = None, 0
    if USE_TC:
        ul = umma_layers(m, ("layer3", "layer4", "d1", "b1"))
        if save_feat and SAVE_ACTIVATIONS:
            # hidden activations of the warp MLP, kept for k_warp_bwd (their recomputation was 10 of its 26 kMAC per sample);
            # one [128][ld] block per forward call, owned by that call's record like the other saved activations
            warp_ld = round_up(n * s, 128)
            warp_saved = torch.empty(128, warp_ld, device=dev)
        h1t = _p(warp_saved[:64]) if warp_saved is not None else None
        h2t = _p(warp_saved[64:]) if warp_saved is not None else None
        if TC_WARP_VERSION == 2:
            pw = m.__dict__.get("_packed_warp")
            if pw is None:
                pw = m.__dict__["_packed_warp"] = PackedWarpFwd(m)
            pw.refresh()
            L.check(lib.dynrf_warp_fwd_tc2(C.byref(cfg), _p(pw.wq), _p(pw.cst), _p(xyz), n, s, _p(tfeat), _p(xyz_prime), h1t, h2t, warp_ld,
                                           _stream()), "dynrf_warp_fwd_tc2")
        else:
            L.check(lib.dynrf_warp_fwd_tc(C.byref(cfg), _p(xyz), n, s, _p(tfeat), C.byref(ul["layer3"]), C.byref(ul["layer4"]),
                                          C.byref(ly["layer5"]), _p(xyz_prime), h1t, h2t, warp_ld, _stream()), "dynrf_warp_fwd_tc")
        if TC_SB_VERSION == 2:
            psb = m.__dict__.get("_packed_sb")
            if psb is None:
                psb = m.__dict__["_packed_sb"] = PackedSigmaBlend(m)
            psb.refresh()
            if save_feat and SAVE_ACTIVATIONS:
                # layer-1 inputs and hidden activations of the density / blending MLPs, kept for the backward: the step is
                # bound by instruction issue, not by HBM, so 1.4 GB of extra stores is cheaper than recomputing layer 1
                # allocated per forward call (not cached on the model): several forwards may be alive before their backwards run
                sb_ld = round_up(n * s, 128)
                sb_saved = dict(XT=torch.empty(225, sb_ld, device=dev), HT=torch.empty(128, sb_ld, device=dev))
            L.check(lib.dynrf_dyn_sigma_blend_fwd_tc2(C.byref(cfg), C.byref(vden), C.byref(vbl), _p(psb.wq), _p(psb.b1cat),
                                                      _p(psb.w2cat), _p(psb.b2cat), _p(xyz), _p(xyz_prime), _p(valid_u8),
                                                      _p(tfeat), n, s, _p(sigma), _p(blend),
                                                      _p(sb_saved["XT"]) if sb_saved else None,
                                                      _p(sb_saved["HT"]) if sb_saved else None, sb_ld, _stream()),
                    "dynrf_dyn_sigma_blend_fwd_tc2")
        else:
            L.check(lib.dynrf_dyn_sigma_blend_fwd_tc(C.byref(cfg), C.byref(vden), C.byref(vbl), C.byref(ul["d1"]),
                                                     C.byref(ly["d2"]), C.byref(ul["b1"]), C.byref(ly["b2"]), _p(xyz),
                                                     _p(xyz_prime), _p(valid_u8), _p(tfeat), n, s, _p(sigma), _p(blend),
                                                     _stream()), "dynrf_dyn_sigma_blend_fwd_tc")
    else:
        L.check(lib.dynrf_warp_fwd(C.byref(cfg), _p(xyz), n, s, _p(tfeat), C.byref(ly["layer3"]), C.byref(ly["layer4"]),
                                   C.byref(ly["layer5"]), _p(xyz_prime), _stream()), "dynrf_warp_fwd")
        L.check(lib.dynrf_dyn_sigma_blend_fwd(C.byref(cfg), C.byref(vden), C.byref(vbl), C.byref(ly["d1"]), C.byref(ly["d2"]),
                                              C.byref(ly["b1"]), C.byref(ly["b2"]), _p(xyz), _p(xyz_prime), _p(valid_u8),
                                              _p(tfeat), n, s, _p(sigma), _p(blend), _stream()), "dynrf_dyn_sigma_blend_fwd")
    weight = torch.empty(n, s, device=dev)
    dists = torch.empty(n, s, device=dev)
    rgb = torch.empty(n, s, 3, device=dev)
    idx, cnt = _active_buffers(n, s, dev)
    L.check(lib.dynrf_alpha_weight_fwd(C.byref(cfg), _p(rays), n, s, _p(sigma), _p(z_vals), _p(weight), _p(dists), _p(idx),
                                       _p(cnt), _p(rgb), _stream()), "dynrf_alpha_weight_fwd")
    vapp = vm_struct(m.app_plane, m.app_line)
    ldf = round_up(n * s, 128)
    featT = torch.empty(32, ldf, device=dev) if save_feat else None  # basis outputs, kept for the tensor-core backward
    if USE_TC:
        pa = packed_app_fwd(m)
        L.check(lib.dynrf_app_fwd_tc(C.byref(cfg), 1, C.byref(vapp), _p(pa.wq), _p(pa.cst), _p(rays), _p(xyz), _p(xyz_prime), _p(tfeat),
                                     n, s, _p(idx), _p(cnt), _p(rgb), _p(featT), ldf, _stream()), "dynrf_app_fwd_tc")
    else:
        L.check(lib.dynrf_dyn_app_fwd(C.byref(cfg), C.byref(vapp), C.byref(ly["basis"]), C.byref(ly["h1"]), C.byref(ly["h2"]),
                                      C.byref(ly["h3"]), _p(rays), _p(xyz), _p(xyz_prime), _p(tfeat), n, s, _p(idx), _p(cnt),
                                      _p(rgb), _p(featT), ldf, _stream()), "dynrf_dyn_app_fwd")
    return dict(sigma=sigma, blend=blend, weight=weight, dists=dists, rgb=rgb, xyz_prime=xyz_prime, tfeat=tfeat, idx=idx,
                cnt=cnt, rays=rays, ts=ts, xyz=xyz, z_vals=z_vals, valid=valid_u8, featT=featT, ldf=ldf, sb_saved=sb_saved, sb_ld=sb_ld,
                warp_saved=warp_saved, warp_ld=warp_ld)


def composite_raw(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays, white_bg: bool, ray_type: str):
    lib = L.load()
    _require_cuda(rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays)
    if rays.dim() != 2 or rays.shape[1] != 6:
        # the kernels index o_z / d_z as rays[n*6+2], rays[n*6+5] (the reference reads rays_chunk[..., 2] + rays_chunk[..., -1])
        raise ValueError(f"raw2outputs: rays_chunk must be [N,6] (origin, direction); got {tuple(rays.shape)} -- pass rays[:, :6]")
    args = [_f32c(t) for t in (rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays)]
    n, s = args[1].shape
    dev = args[1].device
    out_ray = torch.empty(n, L.RAY_OUT, device=dev)
    w_full = torch.empty(n, s, device=dev)
    w_s = torch.empty(n, s, device=dev)
    w_d = torch.empty(n, s, device=dev)
    rt = {"ndc": L.RAY_NDC, "contract": L.RAY_CONTRACT}.get(ray_type, 2)
    L.check(lib.dynrf_composite_fwd(rt, int(bool(white_bg)), *[_p(a) for a in args], n, s, _p(out_ray), _p(w_full), _p(w_s),
                                    _p(w_d), _stream()), "dynrf_composite_fwd")
    return out_ray, w_full, w_s, w_d, args
