"""torch.autograd.Function wrappers over the *_bwd entry points of libdynrf_b200.so.

Gradient coverage (SURVEY.md 8a "autograd requirements per pass"):
  * every parameter of both fields (planes, lines, basis, heads, warp / density / blending MLPs, time MLP);
  * static field: d/dxyz (-> rays, poses, focal; PASS 5-9), d/dviewdir, d/ddists;
  * compositing: every tensor input and the NDC background term of `rays`;
  * dynamic field: d/ddists only.  Gradients of the dynamic outputs w.r.t. xyz / view direction are not
    produced: in the reference train loop they only ever reach the loss multiplied by zero (only the *_s
    outputs of the passes that feed rays with grad are used, train.py:1812,2216,2281) and
    MLPRender_Fea_late_view detaches the view direction.
Small per-ray glue (norm / normalisation chain rule on [N,3] tensors) is done with torch ops inside backward.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as L
from . import ops
from .ops import _p, _stream, _f32c


def available() -> bool:
    try:
        lib = L.load()
        return hasattr(lib, "dynrf_dyn_mlp_bwd")
    except Exception:
        return False


class ZeroArena:
    """One zero-filled buffer per backward call from which the many small accumulators (packed layer gradients, bias / last-layer
    partials, weight-gradient GEMM outputs) are carved: one memset instead of ~40 fill launches."""

    def __init__(self, device, floats: int = 1 << 20):
        self.buf = torch.zeros(floats, device=device)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        a = (n + 15) // 16 * 16        # keep every carve 64-byte aligned (vector reductions need 16)
        if self.off + a > self.buf.numel():
            return torch.zeros(*shape, device=self.buf.device)
        t = self.buf[self.off:self.off + n].view(*shape)
        self.off += a
        return t


class GradLayer:
    """Packed gradient buffers of one nn.Linear."""

    def __init__(self, pl: ops.PackedLayer, arena: "ZeroArena" = None):
        self.pl = pl
        if arena is not None:
            self.dwt, self.db = arena.take(pl.Kpad * pl.Npad), arena.take(pl.Npad)
        else:
            self.dwt = torch.zeros(pl.Kpad * pl.Npad, device=pl.wt.device)
            self.db = torch.zeros(pl.Npad, device=pl.wt.device)
        self.struct = L.LayerGradT(self.dwt.data_ptr(), self.db.data_ptr())

    def unpack(self):
        dw = torch.empty(self.pl.N, self.pl.K, device=self.dwt.device)
        L.check(L.load().dynrf_unpack_linear_grad(_p(self.dwt), self.pl.N, self.pl.K, _p(dw), _stream()), "unpack_linear_grad")
        return dw, (self.db[: self.pl.N].clone() if self.pl.bias is not None else None)


def pack_bwd(pl: ops.PackedLayer, k0: int, k1: int, nrows: int) -> torch.Tensor:
    """wb[n][j] = W[n][k0+j]; cached on the PackedLayer per parameter version."""
    cache = pl.__dict__.setdefault("_bwd_cache", {})
    key = (k0, k1, nrows)
    ver = (ops.PARAM_EPOCH[0], pl.weight.data_ptr(), pl.weight._version)
    hit = cache.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    jpad = (k1 - k0 + 15) // 16 * 16
    wb = torch.empty(nrows * jpad, device=pl.wt.device)
    w = _f32c(pl.weight.detach())
    L.check(L.load().dynrf_pack_linear_bwd(_p(w), pl.N, pl.K, k0, k1, nrows, _p(wb), _stream()), "dynrf_pack_linear_bwd")
    cache[key] = (ver, wb)
    return wb


# When set (the train step sets it: every p.grad is a live view of the flat gradient buffer), the plane / line gradients -- 93 MB
# at cfg 2 -- are accumulated by the scatter kernels STRAIGHT into p.grad and the Functions return None for those inputs,
# instead of filling 30 temporaries with zeros and letting AccumulateGrad add them (the kernels add with atomics either way).
DIRECT_PLANE_GRADS = [False]


def _vm_grads(planes, lines):
    """-> (plane grads, line grads, dynrf_vm_grad_t, returned-to-autograd list)."""
    if DIRECT_PLANE_GRADS[0] and all(p.grad is not None and p.grad.stride() == p.stride() for p in list(planes) + list(lines)):
        gp, gl = [p.grad for p in planes], [p.grad for p in lines]
        return gp, gl, ops.vm_grad_struct(gp, gl), [None] * (len(gp) + len(gl))
    gp = [ops.zeros_like_cl(p) for p in planes]
    gl = [ops.zeros_like_cl(p) for p in lines]
    return gp, gl, ops.vm_grad_struct(gp, gl), gp + gl


def _rays_grad_from(rays, z_vals, scale, gdists, gvd):
    """Chain rule of dists = dz*|d|*scale and viewdir = d/|d| back to rays[:, 3:6]."""
    d = rays[:, 3:6]
    nrm = d.norm(dim=-1, keepdim=True)
    v = d / nrm
    g = torch.zeros_like(rays)
    if gdists is not None:
        dz = torch.cat((z_vals[:, 1:] - z_vals[:, :-1], torch.zeros_like(z_vals[:, :1])), -1)
        gn = (gdists * dz).sum(-1, keepdim=True) * scale
        g[:, 3:6] += gn * v
    if gvd is not None:
        g[:, 3:6] += (gvd - (gvd * v).sum(-1, keepdim=True) * v) / nrm
    return g


# ----------------------------------------------------------------------------- sampling


class SamplePointsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, zvec, model, ray_type):
        cfg = model.native_cfg(ray_type)
        xyz, z, valid = ops.sample_points(cfg, rays.detach(), zvec)
        ctx.save_for_backward(rays.detach(), zvec)
        ctx.ray_type = ray_type
        ctx.mark_non_differentiable(z, valid)
        return xyz, z, valid

    @staticmethod
    def backward(ctx, gxyz, _gz, _gv):
        rays, zvec = ctx.saved_tensors
        if ctx.ray_type == "ndc":
            g = torch.cat((gxyz.sum(1), (gxyz * zvec[None, :, None]).sum(1)), -1)
            return g, None, None, None
        # contraction Jacobian (models/tensorBase.py:552-556): recompute the tiny formula under autograd
        with torch.enable_grad():
            r = rays.clone().requires_grad_(True)
            p = r[:, None, :3] + r[:, None, 3:6] * zvec[None, :, None]
            linf = p.abs().amax(-1, keepdim=True)
            pc = torch.where(linf > 1.0, (2 - 1 / linf) * (p / linf), p)
            (g,) = torch.autograd.grad(pc, r, gxyz)
        return g, None, None, None


# ----------------------------------------------------------------------------- fields


def _static_params(m):
    rm = m.renderModule
    last = rm.mlp_view[0] if rm.split_view else rm.mlp[4]
    return (list(m.density_plane) + list(m.density_line) + list(m.app_plane) + list(m.app_line)
            + [m.basis_mat.weight, rm.mlp[0].weight, rm.mlp[0].bias, rm.mlp[2].weight, rm.mlp[2].bias, last.weight, last.bias])


class StaticFieldFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ray_type, rays, xyz, z_vals, valid, *params):
        st = ops.static_forward_raw(model, rays.detach(), xyz.detach(), z_vals.detach(), valid, ray_type)
        if ops.USE_TC:   # basis outputs, kept for the tensor-core head backward
            st["ldf"] = ops.round_up(st["sigma"].numel(), 128)
            st["featT"] = torch.empty(32, st["ldf"], device=st["sigma"].device)
        ops.static_app_raw(model, st, ray_type)
        ctx.model, ctx.ray_type, ctx.st = model, ray_type, st
        ctx.set_materialize_grads(False)
        return st["sigma"], st["weight"], st["rgb"], st["dists"]

    @staticmethod
    def backward(ctx, gsigma, gweight, grgb, gdists):
        lib = L.load()
        m, rt, st = ctx.model, ctx.ray_type, ctx.st
        cfg = m.native_cfg(rt)
        n, s = st["sigma"].shape
        dev = st["sigma"].device
        gsig = _f32c(gsigma).clone() if gsigma is not None else torch.zeros(n, s, device=dev)
        gd = _f32c(gdists).clone() if gdists is not None else torch.zeros(n, s, device=dev)
        if gweight is not None:
            L.check(lib.dynrf_alpha_weight_bwd(n, s, _p(st["sigma"]), _p(st["dists"]), _p(_f32c(gweight)), _p(gsig), _p(gd),
                                               _stream()), "dynrf_alpha_weight_bwd")
        need_xyz = ctx.needs_input_grad[3]
        need_rays = ctx.needs_input_grad[2]
        gxyz = torch.zeros(n, s, 3, device=dev) if need_xyz else None
        gvd = torch.zeros(n, 3, device=dev) if need_rays else None
        ly = m.packed_layers()
        pk = m._packed
        arena = ZeroArena(dev)
        gden_p, gden_l, gden, ret_den = _vm_grads(m.density_plane, m.density_line)
        gapp_p, gapp_l, gapp, ret_app = _vm_grads(m.app_plane, m.app_line)
        gl = {k: GradLayer(pk[k], arena) for k in ("basis", "h1", "h2", "h3")}
        tc_head = None
        if grgb is not None and ops.USE_TC and st.get("featT") is not None:
            tc_head = _head_bwd_tc(m, cfg, st, st["xyz"], None, _f32c(grgb), gl, gapp, gxyz, gvd, n, s, arena)
        elif grgb is not None:
            vapp = ops.vm_struct(m.app_plane, m.app_line)
            wb2 = pack_bwd(pk["h2"], 0, 128, 128)
            wb1 = pack_bwd(pk["h1"], 0, pk["h1"].K, 128)
            wbB = pack_bwd(pk["basis"], 0, 72, 32)
            L.check(lib.dynrf_static_app_bwd(C.byref(cfg), C.byref(vapp), C.byref(gapp), C.byref(ly["basis"]), C.byref(ly["h1"]),
                                             C.byref(ly["h2"]), C.byref(ly["h3"]), m.head_mode, _p(wb2), _p(wb1), _p(wbB),
                                             _p(st["rays"]), _p(st["xyz"]), n, s, _p(st["idx"]), _p(st["cnt"]), _p(_f32c(grgb)),
                                             C.byref(gl["basis"].struct), C.byref(gl["h1"].struct), C.byref(gl["h2"].struct),
                                             C.byref(gl["h3"].struct), _p(gxyz), _p(gvd), _stream()), "dynrf_static_app_bwd")
        vden = ops.vm_struct(m.density_plane, m.density_line)
        L.check(lib.dynrf_static_density_bwd(C.byref(cfg), C.byref(vden), C.byref(gden), _p(st["xyz"]), _p(st["valid"]),
                                             _p(st["sigma"]), _p(gsig), n, s, _p(gxyz), _stream()), "dynrf_static_density_bwd")
        grays = _rays_grad_from(st["rays"], st["z_vals"], float(m.distance_scale), gd, gvd) if need_rays else None
        gb, _ = gl["basis"].unpack()
        if tc_head is not None:
            (g1w, g1b), (g2w, g2b), (g3w, g3b) = tc_head["h1"], tc_head["h2"], tc_head["h3"]
        else:
            g1w, g1b = gl["h1"].unpack()
            g2w, g2b = gl["h2"].unpack()
            g3w, g3b = gl["h3"].unpack()
        pg = ret_den + ret_app + [gb, g1w, g1b, g2w, g2b, g3w, g3b]
        return (None, None, grays, gxyz, None, None, *pg)


def _dynamic_params(m):
    rm = m.renderModule
    lin = [m.basis_mat, rm.mlp[0], rm.mlp[2], rm.mlp_view[0], m.layer1, m.layer2, m.layer3, m.layer4, m.layer5,
           m.density_layer1, m.density_layer2, m.blending_layer1, m.blending_layer2]
    ps = (list(m.blending_plane) + list(m.blending_line) + list(m.density_plane) + list(m.density_line)
          + list(m.app_plane) + list(m.app_line))
    for l in lin:
        ps.append(l.weight)
        if l.bias is not None:
            ps.append(l.bias)
    return ps


_DYN_LAYER_ORDER = ["basis", "h1", "h2", "h3", "layer1", "layer2", "layer3", "layer4", "layer5", "d1", "d2", "b1", "b2"]


class DynamicFieldFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ray_type, rays, ts, xyz, z_vals, valid, *params):
        r = ops.dynamic_forward_raw(model, rays.detach(), ts.detach(), xyz.detach(), z_vals.detach(), valid, ray_type,
                                    save_feat=ops.USE_TC)
        ctx.model, ctx.ray_type, ctx.r = model, ray_type, r
        ctx.set_materialize_grads(False)
        return r["blend"], r["weight"], r["xyz_prime"], r["rgb"], r["sigma"], r["dists"]

    @staticmethod
    def backward(ctx, gblend, gweight, gxyzp, grgb, gsigma, gdists):
        lib = L.load()
        m, rt, r = ctx.model, ctx.ray_type, ctx.r
        if ctx.needs_input_grad[4] and any(g is not None for g in (gblend, gweight, gxyzp, grgb, gsigma)):
            # nothing silently detaches: d(dynamic outputs)/d(xyz) is not produced by the backward kernels (the reference train
            # loop only ever multiplies it by zero: train.py:1094-1159 detaches the rays of pass 1, :1812,2216,2281 use *_s only)
            raise NotImplementedError(
                "dynrf_b200: gradients of the DYNAMIC field's outputs with respect to the sample positions (xyz_sampled) are not "
                "implemented; pass xyz_sampled.detach() to the dynamic model (the static field provides d/dxyz for pose / focal "
                "optimisation).")
        cfg = m.native_cfg(rt)
        n, s = r["sigma"].shape
        dev = r["sigma"].device
        gsig = _f32c(gsigma).clone() if gsigma is not None else torch.zeros(n, s, device=dev)
        gd = _f32c(gdists).clone() if gdists is not None else torch.zeros(n, s, device=dev)
        if gweight is not None:
            L.check(lib.dynrf_alpha_weight_bwd(n, s, _p(r["sigma"]), _p(r["dists"]), _p(_f32c(gweight)), _p(gsig), _p(gd),
                                               _stream()), "dynrf_alpha_weight_bwd")
        gxp = _f32c(gxyzp).clone() if gxyzp is not None else torch.zeros(n, s, 3, device=dev)
        ly = m.packed_layers()
        pk = m._packed
        arena = ZeroArena(dev)
        gl = {k: GradLayer(pk[k], arena) for k in _DYN_LAYER_ORDER}
        gbl_p, gbl_l, gbl, ret_bl = _vm_grads(m.blending_plane, m.blending_line)
        gden_p, gden_l, gden, ret_den = _vm_grads(m.density_plane, m.density_line)
        gapp_p, gapp_l, gapp, ret_app = _vm_grads(m.app_plane, m.app_line)
        tc_head = None
        if grgb is not None and ops.USE_TC and r.get("featT") is not None:
            tc_head = _head_bwd_tc(m, cfg, r, r["xyz_prime"], r["tfeat"], _f32c(grgb), gl, gapp, gxp, None, n, s, arena)
        elif grgb is not None:
            vapp = ops.vm_struct(m.app_plane, m.app_line)
            wb2 = pack_bwd(pk["h2"], 0, 128, 128)
            wb1f = pack_bwd(pk["h1"], 0, 27, 128)
            wbB = torch.cat([pack_bwd(pk["basis"], 72 * i, 72 * i + 72, 32) for i in range(3)])
            L.check(lib.dynrf_dyn_app_bwd(C.byref(cfg), C.byref(vapp), C.byref(gapp), C.byref(ly["basis"]), C.byref(ly["h1"]),
                                          C.byref(ly["h2"]), C.byref(ly["h3"]), _p(wb2), _p(wb1f), _p(wbB), _p(r["rays"]),
                                          _p(r["xyz"]), _p(r["xyz_prime"]), _p(r["tfeat"]), n, s, _p(r["idx"]), _p(r["cnt"]),
                                          _p(_f32c(grgb)), C.byref(gl["basis"].struct), C.byref(gl["h1"].struct),
                                          C.byref(gl["h2"].struct), C.byref(gl["h3"].struct), _p(gxp), _stream()),
                    "dynrf_dyn_app_bwd")
        vden = ops.vm_struct(m.density_plane, m.density_line)
        vbl = ops.vm_struct(m.blending_plane, m.blending_line)
        tc_sb = None
        if ops.USE_TC:
            tc_sb = _dyn_mlp_bwd_tc(m, cfg, r, vden, vbl, gden, gbl, gsig, _f32c(gblend) if gblend is not None else None, gxp, n, s, arena)
        for mode, vm, gvm, k1, k2, gout in ((0, vden, gden, "d1", "d2", gsig), (1, vbl, gbl, "b1", "b2", gblend)):
            if gout is None or tc_sb is not None:
                continue
            wbf = pack_bwd(pk[k1], 0, 72, 64)
            L.check(lib.dynrf_dyn_mlp_bwd(C.byref(cfg), mode, C.byref(vm), C.byref(gvm), C.byref(ly[k1]), C.byref(ly[k2]), _p(wbf),
                                          _p(r["xyz"]), _p(r["xyz_prime"]), _p(r["valid"]), _p(r["tfeat"]), n, s, _p(_f32c(gout)),
                                          C.byref(gl[k1].struct), C.byref(gl[k2].struct), _p(gxp), _stream()), "dynrf_dyn_mlp_bwd")
        gtout = arena.take(n, 32)
        wb4 = pack_bwd(pk["layer4"], 0, 64, 64)
        wb3t = pack_bwd(pk["layer3"], 63, 93, 64)
        wsv = r.get("warp_saved")
        L.check(lib.dynrf_warp_bwd(C.byref(cfg), _p(r["xyz"]), n, s, _p(r["tfeat"]), C.byref(ly["layer3"]), C.byref(ly["layer4"]),
                                   C.byref(ly["layer5"]), _p(wb4), _p(wb3t), _p(gxp), C.byref(gl["layer3"].struct),
                                   C.byref(gl["layer4"].struct), C.byref(gl["layer5"].struct), _p(gtout),
                                   _p(wsv[:64]) if wsv is not None else None, _p(wsv[64:]) if wsv is not None else None,
                                   r.get("warp_ld", 0), _stream()),
                "dynrf_warp_bwd")
        L.check(lib.dynrf_time_embed_bwd(_p(r["ts"]), n, C.byref(ly["layer1"]), C.byref(ly["layer2"]), _p(gtout),
                                         C.byref(gl["layer1"].struct), C.byref(gl["layer2"].struct), _stream()),
                "dynrf_time_embed_bwd")
        grays = _rays_grad_from(r["rays"], r["z_vals"], float(m.distance_scale), gd, None) if ctx.needs_input_grad[2] else None
        pg = ret_bl + ret_den + ret_app
        for k in _DYN_LAYER_ORDER:
            if tc_head is not None and k in tc_head:
                dw, db = tc_head[k]
            elif tc_sb is not None and k in tc_sb:
                dw, db = tc_sb[k]
            else:
                dw, db = gl[k].unpack()
            pg.append(dw)
            if db is not None:
                pg.append(db)
        return (None, None, grays, None, None, None, None, *pg)


class _HeadBwdState:
    """Per-model packed weights / constants / scratch of the tensor-core head backward (either field)."""

    def __init__(self, m):
        lib = L.load()
        dev = m.basis_mat.weight.device
        self.dyn = 1 if m.is_dynamic else 0
        ops.check_tc_shapes(m)
        self.xrows = lib.dynrf_head_bwd_xrows(self.dyn)
        self.wq = torch.empty(lib.dynrf_head_bwd_wfloats(self.dyn), device=dev)
        self.cst = torch.empty(lib.dynrf_head_bwd_cst_floats(), device=dev)
        self.colmap = torch.empty(self.xrows, dtype=torch.int32, device=dev)
        self.key = None
        self.src_rows = None
        self.bufs = {}

    def refresh(self, m):
        tbl = m._layer_table()
        l1, l2, l3 = tbl["h1"], tbl["h2"], tbl["h3"]
        key = (ops.PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias))
        if key != self.key:
            ops.check_tc_shapes(m)
            with torch.no_grad():
                L.check(L.load().dynrf_pack_head_bwd(self.dyn, m.head_mode, _p(_f32c(l1.weight.detach())), _p(_f32c(l2.weight.detach())),
                                                     _p(self.wq), _p(self.colmap), _stream()), "dynrf_pack_head_bwd")
                w3 = l3.weight.detach()
                w3v = w3[:, 128:131].t().reshape(-1) if w3.shape[1] == 131 else torch.zeros(9, device=w3.device)
                torch.cat((l1.bias.detach(), l2.bias.detach(), w3[:, :128].t().reshape(-1), w3v, l3.bias.detach()), out=self.cst)
                if self.src_rows is None:                                   # constant map, resolved once (host sync)
                    cm = self.colmap.long()
                    self.src_rows = torch.nonzero(cm >= 0).reshape(-1)     # X1T rows that map to a layer-1 input column
                    self.dst_cols = cm[self.src_rows]
            self.key = key
        return self

    def scratch(self, ld, dev):
        b = self.bufs.get(ld)
        if b is None:
            b = self.bufs[ld] = dict(X1T=torch.empty(self.xrows, ld, device=dev), H1T=torch.empty(128, ld, device=dev),
                                     DH2T=torch.empty(128, ld, device=dev), DH1T=torch.empty(128, ld, device=dev),
                                     DFT=torch.empty(32, ld, device=dev))
        return b


def _wgrad(XT, K, DYT, N, ld, cnt, out, out_ld):
    L.check(L.load().dynrf_wgrad_gemm(_p(XT), ld, K, _p(DYT), ld, N, ld, _p(cnt), _p(out), out_ld, _stream()), "dynrf_wgrad_gemm")


def _head_bwd_tc(m, cfg, r, coords, tfeat, grgb, gl, gapp, gcoord, gvd, n, s, arena=None):
    """Tensor-core appearance-head backward: fused tile kernel (recompute + data gradients) -> wgrad GEMMs -> basis / lookups.
    r: forward record (rays, xyz, featT, ldf, idx, cnt); coords: lookup coordinates; gcoord: their gradient (+=) or None."""
    lib = L.load()
    st = m.__dict__.get("_head_bwd_state")
    if st is None:
        st = m.__dict__["_head_bwd_state"] = _HeadBwdState(m)
    st.refresh(m)
    dev = grgb.device
    ld = r["ldf"]
    b = st.scratch(ld, dev)
    xr = st.xrows
    l1 = m._layer_table()["h1"]
    k1 = l1.weight.shape[1]
    arena = arena if arena is not None else ZeroArena(dev, 1 << 17)
    small = arena.take(3 * 128 + 9 + 3 + 128)    # gw3h [3][128] | gw3v [3][3] | gb3 [3] | gb2 [128]
    gw3h, gw3v, gb3, gb2 = small[:384].view(3, 128), small[384:393].view(3, 3), small[393:396], small[396:524]
    L.check(lib.dynrf_head_bwd_tc(C.byref(cfg), st.dyn, _p(st.wq), _p(st.cst), _p(r["rays"]), _p(r["xyz"]), _p(tfeat), _p(r["featT"]),
                                  _p(grgb), _p(r["idx"]), _p(r["cnt"]), n, s, ld, _p(b["X1T"]), _p(b["H1T"]), _p(b["DH2T"]),
                                  _p(b["DH1T"]), _p(b["DFT"]), _p(gw3h), _p(gw3v), _p(gb3), _p(gb2), _p(gvd), _stream()),
            "dynrf_head_bwd_tc")
    cnt = r["cnt"]
    dw1m = arena.take(128, xr)
    _wgrad(b["X1T"], xr, b["DH1T"], 128, ld, cnt, dw1m, xr)
    dw1 = arena.take(128, k1)
    dw1[:, st.dst_cols] = dw1m[:, st.src_rows]
    db1 = dw1m[:, xr - 1].contiguous()
    dw2 = arena.take(128, 128)
    _wgrad(b["H1T"], 128, b["DH2T"], 128, ld, cnt, dw2, 128)
    k3 = m._layer_table()["h3"].weight.shape[1]
    dw3 = torch.cat((gw3h, gw3v), 1) if k3 == 131 else gw3h
    # basis_mat + lookups (SIMT tile kernel)
    pk = m._packed
    vapp = ops.vm_struct(m.app_plane, m.app_line)
    nstr = 3 if st.dyn else 1
    wbB = torch.cat([pack_bwd(pk["basis"], 72 * i, 72 * i + 72, 32) for i in range(nstr)])
    L.check(lib.dynrf_basis_bwd(C.byref(cfg), st.dyn, C.byref(vapp), C.byref(gapp), _p(wbB), _p(coords), _p(b["DFT"]), ld, n, s,
                                _p(r["idx"]), _p(cnt), _p(gl["basis"].dwt), _p(gcoord), _stream()), "dynrf_basis_bwd")
    return {"h1": (dw1, db1), "h2": (dw2, gb2), "h3": (dw3, gb3)}


class _SbBwdState:
    """Per-model packed weights / constants / scratch of the tensor-core density + blending backward."""

    def __init__(self, m):
        lib = L.load()
        dev = m.density_layer1.weight.device
        self.wq = torch.empty(lib.dynrf_sb_bwd_wfloats(), device=dev)
        self.cst = torch.empty(258, device=dev)
        self.colmap = torch.empty(80, dtype=torch.int32, device=dev)
        self.key = None
        self.cols = None
        self.bufs = {}

    def refresh(self, m):
        ps = [m.density_layer1.weight, m.density_layer1.bias, m.density_layer2.weight, m.density_layer2.bias,
              m.blending_layer1.weight, m.blending_layer1.bias, m.blending_layer2.weight, m.blending_layer2.bias]
        key = (ops.PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in ps)
        if key != self.key:
            with torch.no_grad():
                L.check(L.load().dynrf_pack_sb_bwd(_p(_f32c(ps[0].detach())), _p(_f32c(ps[4].detach())), _p(self.wq), _p(self.colmap),
                                                   _stream()), "dynrf_pack_sb_bwd")
                torch.cat((ps[1].detach(), ps[5].detach(), ps[2].detach().reshape(-1), ps[6].detach().reshape(-1),
                           ps[3].detach().reshape(-1), ps[7].detach().reshape(-1)), out=self.cst)
                if self.cols is None:
                    self.cols = self.colmap.long()
            self.key = key
        return self

    def scratch(self, ld, dev, need_xt=True):
        b = self.bufs.get(ld)
        if b is None:
            b = self.bufs[ld] = dict(DHT=torch.empty(128, ld, device=dev))
        if need_xt and "XT" not in b:
            b["XT"] = torch.empty(225, ld, device=dev)
        return b


def _dyn_mlp_bwd_tc(m, cfg, r, vden, vbl, gden, gbl, gsig, gblend, gxp, n, s, arena=None):
    """Tensor-core backward of the density + blending MLPs (a16 + a17): one fused tile kernel + one wgrad GEMM."""
    lib = L.load()
    st = m.__dict__.get("_sb_bwd_state")
    if st is None:
        st = m.__dict__["_sb_bwd_state"] = _SbBwdState(m)
    st.refresh(m)
    dev = gxp.device
    P = n * s
    ld = ops.round_up(P, 128)
    saved = r.get("sb_saved")
    if saved is not None and r.get("sb_ld") != ld:
        saved = None
    b = st.scratch(ld, dev, need_xt=saved is None)
    xt = saved["XT"] if saved is not None else b["XT"]
    arena = arena if arena is not None else ZeroArena(dev, 1 << 16)
    gsmall = arena.take(258)      # gw2 (128) | gb1 (128) | gb2 (2)
    L.check(lib.dynrf_dyn_mlp_bwd_tc(C.byref(cfg), C.byref(vden), C.byref(vbl), C.byref(gden), C.byref(gbl), _p(st.wq), _p(st.cst),
                                     _p(r["xyz"]), _p(r["xyz_prime"]), _p(r["valid"]), _p(r["tfeat"]), _p(gsig), _p(gblend), n, s, ld,
                                     _p(xt), _p(saved["HT"]) if saved is not None else None, _p(b["DHT"]), _p(gxp), _p(gsmall),
                                     _p(gsmall[128:]), _p(gsmall[256:]), _stream()),
            "dynrf_dyn_mlp_bwd_tc")
    dwf = arena.take(128, 224)
    L.check(lib.dynrf_wgrad_gemm(_p(xt), ld, 224, _p(b["DHT"]), ld, 128, P, None, _p(dwf), 224, _stream()), "dynrf_wgrad_gemm")
    dw_d = torch.empty(64, 152, device=dev)
    dw_b = torch.empty(64, 152, device=dev)
    dw_d[:, :72] = dwf[:64, 80:152]
    dw_b[:, :72] = dwf[64:, 152:224]
    dw_d[:, st.cols] = dwf[:64, :80]
    dw_b[:, st.cols] = dwf[64:, :80]
    return {"d1": (dw_d, gsmall[128:192]), "d2": (gsmall[:64].reshape(1, 64), gsmall[256:257]),
            "b1": (dw_b, gsmall[192:256]), "b2": (gsmall[64:128].reshape(1, 64), gsmall[257:258])}


def field_forward_with_grad(model, rays, ts, xyz, z_vals, valid, ray_type):
    if model.is_dynamic:
        blend, weight, xyzp, rgb, sigma, dists = DynamicFieldFn.apply(model, ray_type, rays, ts, xyz, z_vals, valid,
                                                                       *_dynamic_params(model))
        return None, None, blend, xyz, weight, xyzp, rgb, sigma, z_vals, dists
    sigma, weight, rgb, dists = StaticFieldFn.apply(model, ray_type, rays, xyz, z_vals, valid, *_static_params(model))
    return None, None, None, xyz, weight, None, rgb, sigma, z_vals, dists


# ----------------------------------------------------------------------------- compositing


class CompositeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays, white, ray_type):
        o, w_full, w_s, w_d, args = ops.composite_raw(rgb_s.detach(), sigma_s.detach(), rgb_d.detach(), sigma_d.detach(),
                                                      dists.detach(), blending.detach(), z_vals.detach(), rays.detach(), white,
                                                      ray_type)
        ctx.args, ctx.white, ctx.ray_type = args, white, ray_type
        ctx.set_materialize_grads(False)
        return o, w_full, w_s, w_d

    @staticmethod
    def backward(ctx, go, gwf, gws, gwd):
        lib = L.load()
        rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals, rays = ctx.args
        n, s = sigma_s.shape
        dev = sigma_s.device
        go = _f32c(go) if go is not None else torch.zeros(n, L.RAY_OUT, device=dev)
        g_rgb_s, g_rgb_d = torch.empty(n, s, 3, device=dev), torch.empty(n, s, 3, device=dev)
        g_ss, g_sd, g_di, g_bl = (torch.empty(n, s, device=dev) for _ in range(4))
        g_rays = torch.empty(n, 6, device=dev)
        rt = {"ndc": L.RAY_NDC, "contract": L.RAY_CONTRACT}.get(ctx.ray_type, 2)
        opt = lambda t: None if t is None else _f32c(t)  # noqa: E731
        gwf, gws, gwd = opt(gwf), opt(gws), opt(gwd)
        L.check(lib.dynrf_composite_bwd(rt, int(bool(ctx.white)), _p(rgb_s), _p(sigma_s), _p(rgb_d), _p(sigma_d), _p(dists),
                                        _p(blending), _p(z_vals), _p(rays), n, s, _p(go), _p(gwf), _p(gws), _p(gwd), _p(g_rgb_s),
                                        _p(g_ss), _p(g_rgb_d), _p(g_sd), _p(g_di), _p(g_bl), _p(g_rays), _stream()),
                "dynrf_composite_bwd")
        return g_rgb_s, g_ss, g_rgb_d, g_sd, g_di, g_bl, None, g_rays, None, None


# ----------------------------------------------------------------------------- the timed train step


def _train_setup(step):
    from . import optim

    static, dyn = step.static, step.dyn
    params = list(static.parameters()) + [p for n_, p in dyn.named_parameters() if not n_.startswith("scene_flow")]
    step.params = params
    # one flat gradient buffer; every p.grad is a view into it, so a data-parallel step all-reduces one tensor and the
    # buffer addresses are stable under CUDA-graph replay.  Views keep each parameter's own (channel-last) strides.
    step.flat_grad = optim.alias_grads(params)
    # one-launch Adam over the flat buffer (torch's foreach / capturable Adam is ~150 small launches on these 66 tensors)
    step.opt = optim.FlatAdam(params, step.flat_grad, lr=1e-4, betas=(0.9, 0.99))


def _dense(p):
    from .optim import dense_storage

    return dense_storage(p)


def train_fwd_bwd(step, rays, ts, target):
    """fwd + bwd of both fields through the public API, MSE on the three composited colours; gradients land in step.flat_grad."""
    import renderer as R

    static, dyn = step.static, step.dyn
    if getattr(step, "opt", None) is None:
        _train_setup(step)
    step.flat_grad.zero_()
    import os

    # every p.grad is a live, zeroed view of step.flat_grad for the whole backward below (DYNRF_DIRECT_GRADS=0: A/B switch)
    DIRECT_PLANE_GRADS[0] = os.environ.get("DYNRF_DIRECT_GRADS", "1") == "1"
    rays_s = rays
    if getattr(step, "rays_grad", False):
        # pose / focal optimisation (cfg 3/4): the rays carry requires_grad; as in the reference's passes 5-9 the gradient reaches
        # them through the STATIC field (sample positions, view direction, dists); the dynamic field sees detached rays
        # (train.py:1094-1159 detaches them in pass 1, and only the *_s outputs of the pose passes are used: train.py:1812)
        rays_s = rays.detach().clone().requires_grad_(True)
    xyz, z, valid = R.sampleXYZ(dyn, rays_s, step.s, step.rt, getattr(step, "jitter", True))
    fs = static(rays_s, ts, None, xyz, z, valid, is_train=True, ray_type=step.rt, N_samples=step.s)
    fd = dyn(rays, ts, None, xyz.detach(), z, valid, is_train=True, ray_type=step.rt, N_samples=step.s)
    o = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, step.rt)
    loss = ((o[0] - target) ** 2).mean() + ((o[4] - target) ** 2).mean() + ((o[8] - target) ** 2).mean()
    try:
        loss.backward()
    finally:
        DIRECT_PLANE_GRADS[0] = False
    if rays_s is not rays:
        step.rays_grad_out = rays_s.grad
    out = torch.cat((o[0], o[1][:, None], o[2][:, None], o[4], o[5][:, None], o[6][:, None], o[8], o[9][:, None],
                     o[10][:, None], o[12][:, None]), -1).detach()
    return out


def train_update(step):
    step.opt.step()


def train_step(step, rays, ts, target):
    """One data-parallel optimisation step: local fwd + bwd, one all-reduce of the flat gradient, Adam."""
    from .parallel import allreduce_flat_mean

    out = train_fwd_bwd(step, rays, ts, target)
    allreduce_flat_mean(step.flat_grad, step.world)
    train_update(step)
    return out
