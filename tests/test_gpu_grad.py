"""GPU gradient parity (-m gpu): backward kernels vs autograd through the CPU oracle (itself pinned
against autograd through the unmodified reference by tests/test_oracle_golden.py::test_gradients)."""
import pytest
import torch

from tests import helpers as H
from oracle import rodynrf_oracle as O

gpu = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _grad_ok(a, b, loose=1e-3, l2=2e-3, kink_rows=2, kink_loose=3e-2, kink_elems=256, n_samples=None):
    """Tight gate with EXPLICIT exceptions for the two discontinuities of the reference's own arithmetic.

    Observed on B200: parameters whose gradient does not pass through a discontinuity agree to max-norm <= 4e-5 of max|ref| and
    relative L2 <= 3e-5.  Two things are discontinuous in the reference itself and can land on the other side on the GPU:
      (1) a ReLU whose pre-activation is ~1e-7 (one in ~10^6 hidden units: a few per test).  It moves ONE row (the unit's
          incoming weights / bias entry) or ONE column (its outgoing weights) of a weight gradient by one sample's contribution
          (observed 4e-3 of max), and -- for layers UPSTREAM of it -- adds that sample's contribution spread over the whole
          tensor: e.g. one flipped unit of the warp MLP changes one ray's d loss / d t_out by ~1/sqrt(S * 64), i.e. the
          time-embedding gradients by ~0.15 / sqrt(#samples) in relative L2 (2.4e-3 at 96 x 40 samples, 6e-4 at 256 x 256);
      (2) the appearance mask `weight > 1e-4` (models/tensorBase.py:774): a sample within 1e-7 of the threshold switches its
          colour on or off, which changes d loss / d weight of THAT sample by O(1) and hence a handful of texels of the plane /
          line gradients (observed max-norm 1.4e-2 on <= 16 entries, L2 unaffected).
    Gate: max-norm <= max(1e-3, 0.15 / sqrt(n_samples)) of max|ref| and relative L2 <= max(2e-3, 0.15 / sqrt(n_samples)).  If the
    max-norm part fails, up to
    `kink_rows` rows and columns (2-D) and up to `kink_elems` single entries (any shape; never more than 1 %) may be set aside provided each is
    within `kink_loose` of max|ref|; everything else must pass the tight gate.  A wrong or missing term fails by orders of
    magnitude.  Returns (ok, max-norm error, relative L2 error)."""
    a, b = a.contiguous(), b.contiguous()       # channel-last gradients arrive with permuted strides
    scale = float(b.abs().max()) + 1e-12
    err = (a - b).abs() / scale
    emax = float(err.max())
    rel2 = float((a - b).norm() / (b.norm() + 1e-12))
    if n_samples:                               # the spread effect of one upstream kink, see (1)
        l2 = max(l2, 0.15 / n_samples ** 0.5)
        loose = max(loose, 0.15 / n_samples ** 0.5)
    if emax <= loose and rel2 <= l2:
        return True, emax, rel2
    if emax > kink_loose or a.dim() == 0:
        return False, emax, rel2
    keep = torch.ones_like(err, dtype=torch.bool)
    if a.dim() == 2 and kink_rows > 0:
        for dim in (1, 0):                                  # worst rows, then worst columns
            worst = (err * keep).amax(dim=dim)
            idx = torch.topk(worst, min(kink_rows, worst.numel())).indices
            idx = idx[worst[idx] > loose]
            if dim == 1:
                keep[idx, :] = False
            else:
                keep[:, idx] = False
    if kink_elems > 0:
        flat = (err * keep).reshape(-1)
        idx = torch.topk(flat, min(kink_elems, max(8, flat.numel() // 100), flat.numel())).indices      # never more than 1 % of the tensor
        keep.view(-1)[idx[flat[idx] > loose]] = False
    d = (a - b) * keep
    ok = float((d.abs() / scale).max()) <= loose and float(d.norm() / ((b * keep).norm() + 1e-12)) <= l2
    return ok, emax, rel2


def _loss_weights(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes]


def _oracle_full(case, rays_grad=False):
    ss = {k: v.clone().requires_grad_(True) for k, v in case["sd_static"].items()}
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd_dynamic"].items()}
    rays = case["rays"].clone().requires_grad_(rays_grad)
    xyz, z, valid = O.sampleXYZ(case["cd"], rays, case["n_samples"], case["ray_type"], False)
    fs = O.field_forward(case["cs"], ss, rays, case["ts"], xyz, z, valid, case["ray_type"])
    fd = O.field_forward(case["cd"], sd, rays, case["ts"], xyz, z, valid, case["ray_type"])
    o = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, case["ray_type"])
    return ss, sd, rays, fs, fd, o


def _native_full(case, rays_grad=False):
    import renderer as R

    static, dyn = H.native_models(case, DEV)
    rays = case["rays"].to(DEV).requires_grad_(rays_grad)
    ts = case["ts"].to(DEV)
    xyz, z, valid = R.sampleXYZ(dyn, rays, case["n_samples"], case["ray_type"], False)
    fs = static(rays, ts, None, xyz, z, valid, ray_type=case["ray_type"])
    fd = dyn(rays, ts, None, xyz, z, valid, ray_type=case["ray_type"])
    o = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, case["ray_type"])
    return static, dyn, rays, fs, fd, o


@gpu
@pytest.mark.parametrize("ray_type,static_mode,grid,n,s", [
    ("ndc", "MLP_Fea", [21, 24, 18], 96, 40),
    ("contract", "MLP_Fea_TimeEmbedding", [16, 19, 20], 70, 33),
    ("ndc", "MLP_Fea", [40, 44, 36], 200, 150),
])
def test_parameter_gradients(ray_type, static_mode, grid, n, s):
    _check_parameter_gradients(H.build_pair(grid, n, s, ray_type=ray_type, static_mode=static_mode, seed=7 * n + s))


def _check_parameter_gradients(case):
    ss, sd, _, fs, fd, o = _oracle_full(case)
    shapes = [t.shape for t in o] + [fd[5].shape, fd[2].shape, fs[4].shape]
    ws = _loss_weights(shapes, 99)
    loss = sum((t * w).sum() for t, w in zip(list(o) + [fd[5], fd[2], fs[4]], ws))
    loss.backward()

    static, dyn, _, fs2, fd2, o2 = _native_full(case)
    loss2 = sum((t * w.to(DEV)).sum() for t, w in zip(list(o2) + [fd2[5], fd2[2], fs2[4]], ws))
    loss2.backward()
    torch.cuda.synchronize()
    assert abs(float(loss2) - float(loss)) <= 1e-4 * max(1.0, abs(float(loss)))
    worst = {}
    for tag, mod, sdict in (("s", static, ss), ("d", dyn, sd)):
        for name, p in mod.named_parameters():
            if name.startswith("scene_flow"):
                continue
            ref = sdict[name].grad
            assert ref is not None, name
            assert p.grad is not None, f"{tag}:{name} got no gradient"
            assert p.grad.shape == ref.shape
            worst[f"{tag}:{name}"] = _grad_ok(p.grad.cpu(), ref, n_samples=case["rays"].shape[0] * case["n_samples"])
    bad = {k: v for k, v in worst.items() if not v[0]}
    print(sorted(((k, v[1], v[2]) for k, v in worst.items()), key=lambda kv: -kv[1])[:8])
    assert not bad, bad


@gpu
def test_parameter_gradients_softplus():
    """softplus density activation: sigma' = sigmoid(f + shift) in both density backward kernels."""
    case = H.build_pair([24, 26, 22], 400, 96, ray_type="ndc", static_mode="MLP_Fea", seed=41, act="softplus", density_shift=-1.5)
    _check_parameter_gradients(case)


@gpu
def test_parameter_gradients_full_grid():
    """cfg 2's 331x368x220 grid, 256 rays x 256 samples (the oracle's autograd finishes in seconds at this batch)."""
    test_parameter_gradients("ndc", "MLP_Fea", [331, 368, 220], 256, 256)


@gpu
@pytest.mark.parametrize("ray_type,static_mode", [("ndc", "MLP_Fea"), ("contract", "MLP_Fea_TimeEmbedding")])
def test_static_ray_gradients(ray_type, static_mode):
    """PASS 5 shape: rays carry grad -> xyz -> static field; d loss / d rays (pose / focal chain)."""
    import renderer as R

    case = H.build_pair([24, 28, 20], 80, 48, ray_type=ray_type, static_mode=static_mode, seed=21)
    ss = {k: v.clone() for k, v in case["sd_static"].items()}
    rays = case["rays"].clone().requires_grad_(True)
    xyz, z, valid = O.sampleXYZ(case["cs"], rays, 48, ray_type, False)
    fs = O.field_forward(case["cs"], ss, rays, case["ts"], xyz, z, valid, ray_type)
    ws = _loss_weights([fs[7].shape, fs[6].shape, fs[4].shape, fs[9].shape], 5)
    loss = (fs[7] * ws[0]).sum() + (fs[6] * ws[1]).sum() + (fs[4] * ws[2]).sum() + (fs[9] * ws[3]).sum()
    loss.backward()

    static, dyn = H.native_models(case, DEV)
    rays2 = case["rays"].to(DEV).requires_grad_(True)
    xyz2, z2, valid2 = R.sampleXYZ(static, rays2, 48, ray_type, False)
    f2 = static(rays2, case["ts"].to(DEV), None, xyz2, z2, valid2, ray_type=ray_type)
    loss2 = sum((t * w.to(DEV)).sum() for t, w in zip((f2[7], f2[6], f2[4], f2[9]), ws))
    loss2.backward()
    torch.cuda.synchronize()
    ok, r, n_out = _grad_ok(rays2.grad.cpu(), rays.grad)
    print("rays grad rel err", r, n_out)
    assert ok


@gpu
@pytest.mark.parametrize("ray_type,white", [("ndc", False), ("ndc", True), ("contract", False), ("contract", True)])
def test_composite_gradients(ray_type, white):
    from dynrf_b200 import backward as B

    g = torch.Generator().manual_seed(3)
    n, s = 37, 150
    rgb_s, rgb_d = torch.rand(n, s, 3, generator=g), torch.rand(n, s, 3, generator=g)
    sig_s, sig_d = torch.relu(torch.randn(n, s, generator=g)) * 2, torch.relu(torch.randn(n, s, generator=g)) * 2
    dists = torch.rand(n, s, generator=g) * 0.3
    blend = torch.rand(n, s, generator=g)
    z = torch.sort(torch.rand(n, s, generator=g), -1).values
    rays, _ = H.synth_rays(n, ray_type, g)
    ins = [rgb_s, sig_s, rgb_d, sig_d, dists, blend, z, rays]
    a = [t.clone().requires_grad_(True) for t in ins]
    o = O.raw2outputs(*a[:7], a[7], is_train=True, ray_type=ray_type, white_bg_draw=white)
    ws = _loss_weights([t.shape for t in o], 17)
    sum((t * w).sum() for t, w in zip(o, ws)).backward()

    b = [t.clone().to(DEV).requires_grad_(True) for t in ins]
    oo, wf, wsn, wd = B.CompositeFn.apply(*b, white, ray_type)
    from dynrf_b200.autograd import _split_ray_out

    rf, df, af, rs, ds, as_, rd, dd, ad, dyn = _split_ray_out(oo)
    o2 = (rf, df, af, wf, rs, ds, as_, wsn, rd, dd, ad, wd, dyn)
    for x, y in zip(o, o2):  # (1 - acc) * 256 in the contract depth maps amplifies 1-ulp differences of acc
        torch.testing.assert_close(y.cpu(), x.detach(), rtol=1e-4, atol=1e-4 if ray_type == "contract" else 2e-5)
    sum((t * w.to(DEV)).sum() for t, w in zip(o2, ws)).backward()
    torch.cuda.synchronize()
    names = ["rgb_s", "sigma_s", "rgb_d", "sigma_d", "dists", "blending", "z", "rays"]
    for nme, x, y in zip(names, a, b):
        if nme == "z":
            continue
        if nme == "rays":
            # only the NDC background term (1 - acc) * (o_z + d_z) reaches rays; with acc ~ 1 it is O(1e-7) noise,
            # so compare absolutely.  Contracted rays get no gradient from the compositing.
            want = torch.zeros_like(x) if x.grad is None else x.grad
            got = torch.zeros_like(want) if y.grad is None else y.grad.cpu()
            torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-5)
            continue
        ok, r, r2 = _grad_ok(y.grad.cpu(), x.grad)
        print(nme, r, r2)
        assert ok, (nme, r, r2)


@gpu
def test_reference_golden_gradients(golden_case):
    """Parameter gradients against values produced by autograd through the UNMODIFIED reference."""
    import renderer as R

    g = golden_case
    case = H.build_pair(list(g.raw["grid"]), 0, int(g.raw["n_samples"]), ray_type=g.meta["ray_type"],
                        static_mode=g.meta["static_mode"], sd_static=g.sd("static"), sd_dynamic=g.sd("dynamic"),
                        rays=g.t("rays"), ts=g.t("ts"), aabb=g.t("aabb"), near_far=tuple(float(v) for v in g.raw["near_far"]))
    static, dyn, rays, fs, fd, o = _native_full(case)
    gen = torch.Generator().manual_seed(20211202 + 13)
    loss = 0.0
    for t in o:
        loss = loss + (t * torch.randn(t.shape, generator=gen).to(DEV)).sum()
    loss = loss + (fd[5] * torch.randn(fd[5].shape, generator=gen).to(DEV)).sum() * 0.1
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g.raw["g_loss"])) <= 1e-3
    bad = {}
    for tag, mod in (("s", static), ("d", dyn)):
        for name, p in mod.named_parameters():
            key = f"g_{tag}/{name}/abssum"
            if key not in g.raw:
                continue
            got = float(p.grad.double().abs().sum())
            want = float(g.raw[key])
            if abs(got - want) > 2e-3 * max(want, 1e-6):
                bad[f"{tag}:{name}"] = (got, want)
            # element-wise: the first 256 entries (NCHW-flattened order) of the gradient the reference's autograd produced
            head = g.t(f"g_{tag}/{name}/head")
            mine = p.grad.detach().cpu().contiguous().reshape(-1)[: head.numel()]
            scale = max(float(head.abs().max()), 1e-3 * want / max(p.numel(), 1), 1e-12)
            e = float((mine - head).abs().max()) / scale
            if e > 5e-3:
                bad[f"{tag}:{name}:head"] = e
    assert not bad, bad


@gpu
def test_interleaved_forwards_keep_their_own_saved_state():
    """Two forwards alive at once (the reference's train loop runs several passes before one backward): the activations a
    forward keeps for its backward must not be shared between calls."""
    import renderer as R

    case_a = H.build_pair([30, 34, 26], 160, 64, seed=21)
    case_b = H.build_pair([30, 34, 26], 160, 64, seed=22, sd_static=case_a["sd_static"], sd_dynamic=case_a["sd_dynamic"])

    def run(interleave):
        static, dyn = H.native_models(case_a, DEV)
        outs = []
        for c in ([case_a, case_b] if interleave else [case_a]):
            rays, ts = c["rays"].to(DEV), c["ts"].to(DEV)
            xyz, z, valid = R.sampleXYZ(dyn, rays, 64, "ndc", False)
            fs = static(rays, ts, None, xyz, z, valid, ray_type="ndc")
            fd = dyn(rays, ts, None, xyz, z, valid, ray_type="ndc")
            outs.append(R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, "ndc"))
        o = outs[0]
        (o[0].sum() + 0.5 * o[8].sum() + o[4].sum()).backward()
        torch.cuda.synchronize()
        return {n: p.grad.clone() for m_ in (static, dyn) for n, p in m_.named_parameters() if p.grad is not None}

    alone, mixed = run(False), run(True)
    assert alone.keys() == mixed.keys()
    for k in alone:
        ok, emax, e2 = _grad_ok(mixed[k].cpu(), alone[k].cpu(), loose=1e-3, l2=1e-4, kink_rows=0, kink_elems=0)
        assert ok, (k, emax, e2)


@gpu
def test_parameter_gradients_cfg3_contract_256cubed():
    """cfg 3 (DAVIS shape): contracted rays, 256^3 grid, S = 192, static head MLP_Fea_TimeEmbedding; 256 rays so that the
    oracle's autograd finishes in seconds."""
    test_parameter_gradients("contract", "MLP_Fea_TimeEmbedding", [256, 256, 256], 256, 192)


@gpu
def test_dynamic_xyz_gradient_request_raises():
    """Nothing silently detaches: asking the dynamic field for d/d(xyz_sampled) must fail loudly."""
    import renderer as R

    case = H.build_pair([20, 22, 18], 32, 24, seed=5)
    static, dyn = H.native_models(case, DEV)
    rays = case["rays"].to(DEV).requires_grad_(True)
    ts = case["ts"].to(DEV)
    xyz, z, valid = R.sampleXYZ(dyn, rays, 24, "ndc", False)
    assert xyz.requires_grad
    fd = dyn(rays, ts, None, xyz, z, valid, ray_type="ndc")
    with pytest.raises(NotImplementedError):
        fd[7].sum().backward()
    # the documented way: detached positions for the dynamic field, the static field carries d/dxyz
    fd = dyn(rays.detach(), ts, None, xyz.detach(), z, valid, ray_type="ndc")
    fs = static(rays, ts, None, xyz, z, valid, ray_type="ndc")
    (fd[7].sum() + fs[7].sum() + fs[6].sum()).backward()
    assert rays.grad is not None and bool(torch.isfinite(rays.grad).all())


@gpu
def test_tc_pack_rejects_unsupported_hyperparameters():
    """the tensor-core packs are compiled for the shipped shapes; anything else raises instead of mis-reading weights."""
    import models.tensoRF as T
    import renderer as R

    dev = torch.device(DEV)
    aabb = torch.tensor(H.NDC_AABB, device=dev)
    kw = dict(H.COMMON, near_far=[0.0, 1.0])
    kw["featureC"] = 64
    m = T.TensorVMSplit(aabb, [16, 16, 16], 12, dev, shadingMode="MLP_Fea", fea_pe=2, **kw)
    g = torch.Generator().manual_seed(0)
    rays, ts = H.synth_rays(16, "ndc", g)
    rays, ts = rays.to(dev), ts.to(dev)
    with torch.no_grad():
        xyz, z, valid = R.sampleXYZ(m, rays, 16, "ndc", False)
        with pytest.raises(NotImplementedError):
            m(rays, ts, None, xyz, z, valid, ray_type="ndc")
    kw = dict(H.COMMON, near_far=[0.0, 1.0])
    m2 = T.TensorVMSplit(aabb, [16, 16, 16], 12, dev, shadingMode="MLP_Fea", fea_pe=6, **kw)   # class-default fea_pe
    with torch.no_grad(), pytest.raises(NotImplementedError):
        m2(rays, ts, None, xyz, z, valid, ray_type="ndc")
    with pytest.raises(ValueError):
        R.raw2outputs(torch.zeros(4, 8, 3, device=dev), torch.zeros(4, 8, device=dev), torch.zeros(4, 8, 3, device=dev),
                      torch.zeros(4, 8, device=dev), torch.zeros(4, 8, device=dev), torch.zeros(4, 8, device=dev),
                      torch.zeros(4, 8, device=dev), torch.zeros(4, 7, device=dev))


@gpu
def test_saved_activations_match_recompute():
    """The backward reads the hidden activations the forward kept (density / blending layer 1, warp MLP layers 3, 4) or recomputes
    them (ops.SAVE_ACTIVATIONS = False): same parameter gradients either way (the kept values come from the tensor-core forward,
    the recomputed ones from the fp32 SIMT / tensor-core backward: a ReLU at ~1e-7 may flip, hence the kink-tolerant gate)."""
    from dynrf_b200 import ops

    case = H.build_pair([26, 30, 22], 160, 70, ray_type="ndc", static_mode="MLP_Fea", seed=77)
    grads = []
    try:
        for keep in (True, False):
            ops.SAVE_ACTIVATIONS = keep
            static, dyn, _, fs, fd, o = _native_full(case)
            ws = _loss_weights([t.shape for t in o] + [fd[5].shape, fd[2].shape], 5)
            loss = sum((t * w.to(DEV)).sum() for t, w in zip(list(o) + [fd[5], fd[2]], ws))
            loss.backward()
            torch.cuda.synchronize()
            grads.append({n: p.grad.detach().cpu().clone() for n, p in dyn.named_parameters() if p.grad is not None})
    finally:
        ops.SAVE_ACTIVATIONS = True
    bad = {}
    for n, g in grads[0].items():
        ok, emax, rel2 = _grad_ok(g, grads[1][n], n_samples=160 * 70)
        if not ok:
            bad[n] = (emax, rel2)
    assert not bad, bad


@gpu
@pytest.mark.parametrize("n,s", [(1, 16), (3, 43), (5, 128), (9, 129), (300, 64)])
def test_two_tile_pipelines_tile_counts(n, s):
    """The tile kernels keep two tiles in flight per CTA (prologue / epilogue of the software pipeline, odd tile counts, CTAs with
    zero or one tile, a partial last tile): forward parity with the oracle and one backward for P = n * s from 16 to 19200."""
    case = H.build_pair([18, 20, 17], n, s, ray_type="ndc", static_mode="MLP_Fea", seed=100 + n)
    rep = H.compare_pass(H.oracle_pass(case), H.native_pass(case, DEV))
    print(rep)
    if n * s <= 2000:
        _check_parameter_gradients(case)
