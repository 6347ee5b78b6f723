"""Full-size checks (-m gpu) at BASELINE.json's configurations, through properties that do not need the CPU oracle at
that size: shard consistency (rendering two halves of the batch == rendering the batch, bit for bit), agreement of the
two independent kernel families (tensor-core vs fp32 SIMT) within the north-star tolerances, compositing invariants,
and a subsample of rays checked against the oracle itself."""
import pytest
import torch

from tests import helpers as H

gpu = pytest.mark.gpu
DEV = "cuda:0"
CFG2 = dict(grid=[331, 368, 220], n=4096, s=256, ray_type="ndc", static_mode="MLP_Fea")
CFG3 = dict(grid=[256, 256, 256], n=4096, s=192, ray_type="contract", static_mode="MLP_Fea_TimeEmbedding")


def _render(case, models, lo=None, hi=None):
    import renderer as R

    static, dyn = models
    rays, ts = case["rays"][lo:hi].to(DEV), case["ts"][lo:hi].to(DEV)
    rt, s = case["ray_type"], case["n_samples"]
    with torch.no_grad():
        xyz, z, valid = R.sampleXYZ(dyn, rays, s, rt, False)
        fs = static(rays, ts, None, xyz, z, valid, is_train=False, ray_type=rt, N_samples=s)
        fd = dyn(rays, ts, None, xyz, z, valid, is_train=False, ray_type=rt, N_samples=s)
        out = R.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], rays, False, rt)
    torch.cuda.synchronize()
    return fs, fd, out


def _build(cfg, seed):
    case = H.build_pair(cfg["grid"], cfg["n"], cfg["s"], ray_type=cfg["ray_type"], static_mode=cfg["static_mode"], seed=seed)
    return case, H.native_models(case, DEV)


@gpu
@pytest.mark.parametrize("cfg", [CFG2, CFG3], ids=["cfg2_ndc_300cubed", "cfg3_contract_256cubed"])
def test_fullsize_shard_consistency_and_invariants(cfg):
    case, models = _build(cfg, seed=11)
    n = cfg["n"]
    fs, fd, out = _render(case, models)
    cut = n // 2 + 37                                   # ragged split
    a = _render(case, models, 0, cut)
    b = _render(case, models, cut, n)
    for i in range(13):
        assert torch.equal(out[i], torch.cat((a[2][i], b[2][i]))), f"output {i} depends on how the batch is sharded"
    for j in (2, 4, 5, 6, 7):                           # blending, weight, xyz_prime, rgb, sigma of the dynamic field
        assert torch.equal(fd[j], torch.cat((a[1][j], b[1][j])))
    # compositing invariants (renderer.py:173-315)
    rgb_full, depth_full, acc_full, w_full, rgb_s, _, acc_s, w_s, rgb_d, _, acc_d, w_d, dyn = out
    for w, acc in ((w_full, acc_full), (w_s, acc_s)):
        assert float(w.min()) >= 0.0
        assert float((w.sum(-1) - acc).abs().max()) <= 1e-5
    # a single field's weights partition at most unit opacity; the blended weights alpha_d b + alpha_s (1 - b) do not
    # (their sum can exceed 1 -- the reference clamps with relu(1 - acc) for that reason, renderer.py:262-267)
    assert float(acc_s.max()) <= 1.0 + 1e-5
    assert float((w_d.sum(-1) - 1.0).abs().max()) <= 1e-4 or float(acc_d.min()) < 1e-6   # w_d is normalised by its own sum
    for c in (rgb_full, rgb_s, rgb_d):
        assert float(c.min()) >= 0.0 and float(c.max()) <= 1.0
    assert float(dyn.min()) >= 0.0 and float(dyn.max()) <= 1.0 + 1e-5
    # per-sample rgb is exactly 0 outside each field's own appearance mask, sigma is exactly 0 on invalid samples
    assert float(fs[6][fs[4] <= 1e-4].abs().max() if (fs[4] <= 1e-4).any() else 0.0) == 0.0
    assert float(fd[6][fd[4] <= 1e-4].abs().max() if (fd[4] <= 1e-4).any() else 0.0) == 0.0
    assert all(torch.isfinite(t).all() for t in out)


@gpu
@pytest.mark.parametrize("cfg", [CFG2, CFG3], ids=["cfg2", "cfg3"])
def test_fullsize_tensor_core_vs_simt(cfg):
    from dynrf_b200 import ops

    case, models = _build(cfg, seed=12)
    try:
        ops.USE_TC = True
        fs1, fd1, o1 = _render(case, models)
        ops.USE_TC = False
        fs0, fd0, o0 = _render(case, models)
    finally:
        ops.USE_TC = True
    # sigma: the quantity the two MLP engines actually compute
    for f1, f0 in ((fs1, fs0), (fd1, fd0)):
        assert float(((f1[7] - f0[7]).abs() / f0[7].abs().clamp_min(1.0)).max()) <= 2e-5
        # alpha = 1 - exp(-sigma * dists) amplifies a sigma error by dists (up to several hundred on the outer samples of a
        # contracted ray), so the 1e-5 alpha budget is checked where the map is well conditioned
        ok = f0[9] <= 4.0
        a1, a0 = 1 - torch.exp(-f1[7] * f1[9]), 1 - torch.exp(-f0[7] * f0[9])
        assert float((a1 - a0)[ok].abs().max()) <= H.TOL_ALPHA
        assert float((a1 - a0).abs().max()) <= H.TOL_ALPHA * max(4.0, float(f0[9].max()))
    assert float((fd1[2] - fd0[2]).abs().max()) <= 2e-5
    assert float((fd1[5] - fd0[5]).abs().max()) <= 2e-5
    # a sample whose weight sits within 1e-6 of the 1e-4 appearance threshold can flip its mask between the two engines and
    # switch a (w <= 1e-4)-weighted colour on or off; rays with such a flip are compared with the flip's own magnitude added
    flips = ((fs1[4] > 1e-4) != (fs0[4] > 1e-4)) | ((fd1[4] > 1e-4) != (fd0[4] > 1e-4))
    assert int(flips.sum()) <= 1e-4 * flips.numel()
    clean = ~flips.any(-1)
    for i in (0, 4, 8):
        rel = (o1[i] - o0[i]).abs() / o0[i].abs().clamp_min(1e-2)
        assert float(rel[clean].max()) <= H.TOL_RGB_REL, (i, float(rel[clean].max()))
        # a flipped sample moves the composited colour by at most its compositing weight (o[i + 3]) times a colour in [0, 1]
        bound = (o0[i + 3] * flips).sum(-1, keepdim=True) + (o1[i + 3] * flips).sum(-1, keepdim=True) + 1e-4
        assert bool(((o1[i] - o0[i]).abs() <= bound).all())


CFG4 = dict(grid=[706, 786, 471], n=4096, s=256, ray_type="ndc", static_mode="MLP_Fea")


@gpu
@pytest.mark.parametrize("cfg,stride", [(CFG2, 8), (CFG3, 16), (CFG4, 32)],
                         ids=["cfg2_ndc_300cubed", "cfg3_contract_256cubed", "cfg4_ndc_640cubed"])
def test_fullsize_subsample_against_oracle(cfg, stride):
    """Every `stride`-th ray of the full batch through the CPU ORACLE at the configuration's full grid:
    cfg 2 -- 512 NDC rays at [331,368,220]; cfg 3 -- 256 contracted rays x 192 samples at 256^3 (static head
    MLP_Fea_TimeEmbedding); cfg 4 -- 128 NDC rays at [706,786,471] ("640^3", 421 MB of parameters)."""
    case, models = _build(cfg, seed=13)
    sub = dict(case, rays=case["rays"][::stride].contiguous(), ts=case["ts"][::stride].contiguous())
    ref = H.oracle_pass(sub)
    got = H.native_pass(sub, DEV, models=models)
    print(H.compare_pass(ref, got, check=cfg["ray_type"] == "ndc"))
    if cfg["ray_type"] != "ndc":
        # contracted rays: dists reach several hundred on the outer samples, so alpha = 1 - exp(-sigma * dists) amplifies a
        # 1e-6 sigma difference beyond the 1e-5 budget there; check alpha where the map is well conditioned and sigma everywhere
        rep = H.compare_pass(ref, got, check=False)
        assert rep["valid_mismatch"] == 0 and rep["xyz"] <= 1e-5 and rep["z"] <= 1e-4
        for key in ("fs", "fd"):
            r, g = ref[key], got[key]
            assert float(((r[7] - g[7]).abs() / r[7].abs().clamp_min(1.0)).max()) <= 2e-5
            ok = r[9] <= 4.0
            a_r, a_g = 1 - torch.exp(-r[7] * r[9]), 1 - torch.exp(-g[7] * g[9])
            assert float((a_r - a_g)[ok].abs().max()) <= H.TOL_ALPHA
            assert float((r[4] - g[4]).abs().max()) <= 5e-5
        assert rep["d_blending"] <= 2e-5 and rep["d_xyz_prime"] <= 2e-5
        flips = rep["s_mask_flips"] + rep["d_mask_flips"]
        for k in ("rgb_full", "rgb_s", "rgb_d"):
            assert rep[f"o_{k}_rel"] <= H.TOL_RGB_REL + 1e-3 * (flips > 0), (k, rep)
        for k in ("acc_full", "acc_s", "w_full", "w_s", "dyn"):
            assert rep[f"o_{k}"] <= 1e-4, (k, rep)


@gpu
def test_cfg4_grid_640cubed_runs_and_agrees():
    """The largest configured grid ([706,786,471], 421 MB of parameters): both kernel families agree, sharding is exact."""
    from dynrf_b200 import ops

    cfg = dict(grid=[706, 786, 471], n=1024, s=256, ray_type="ndc", static_mode="MLP_Fea")
    case, models = _build(cfg, seed=14)
    _, fd1, o1 = _render(case, models)
    a, b = _render(case, models, 0, 500), _render(case, models, 500, 1024)
    assert torch.equal(o1[0], torch.cat((a[2][0], b[2][0])))
    try:
        ops.USE_TC = False
        _, fd0, o0 = _render(case, models)
    finally:
        ops.USE_TC = True
    assert float(((o1[0] - o0[0]).abs() / o0[0].abs().clamp_min(1e-2)).max()) <= H.TOL_RGB_REL
    assert float((fd1[7] - fd0[7]).abs().max()) <= 1e-4 * max(1.0, float(fd0[7].abs().max()))
