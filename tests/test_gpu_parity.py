"""GPU parity (run on the B200 box with -m gpu): the native drop-in path, called through the
C ABI, against (i) the golden vectors produced by the unmodified reference and (ii) the CPU
oracle on seeded synthetic inputs.  Tolerances are the north-star's: alpha 1e-5, composited RGB
1e-4 relative (tests/helpers.py)."""
import ctypes as C

import pytest
import torch

from tests import helpers as H
from oracle import rodynrf_oracle as O

gpu = pytest.mark.gpu
DEV = "cuda:0"


def _golden_as_ref(g):
    fs = (None, None, None, g.t("xyz"), g.t("s_weight"), None, g.t("s_rgb"), g.t("s_sigma"), g.t("z_vals"), g.t("s_dists"))
    fd = (None, None, g.t("d_blending"), g.t("xyz"), g.t("d_weight"), g.t("d_xyz_prime"), g.t("d_rgb"), g.t("d_sigma"),
          g.t("z_vals"), g.t("d_dists"))
    out = tuple(g.t(f"o13_{i}") for i in range(13))
    return dict(xyz=g.t("xyz"), z=g.t("z_vals"), valid=g.t("ray_valid"), fs=fs, fd=fd, out=out)


def _case_from_golden(g):
    return H.build_pair(list(g.raw["grid"]), 0, int(g.raw["n_samples"]), ray_type=g.meta["ray_type"],
                        static_mode=g.meta["static_mode"], sd_static=g.sd("static"), sd_dynamic=g.sd("dynamic"),
                        rays=g.t("rays"), ts=g.t("ts"), aabb=g.t("aabb"), near_far=tuple(float(v) for v in g.raw["near_far"]))


@gpu
def test_native_vs_reference_golden(golden_case):
    """The whole path against outputs of the UNMODIFIED reference (tests/golden/*.npz)."""
    case = _case_from_golden(golden_case)
    got = H.native_pass(case, DEV)
    rep = H.compare_pass(_golden_as_ref(golden_case), got)
    print(rep)


@gpu
@pytest.mark.parametrize("ray_type,static_mode,grid,n,s", [
    ("ndc", "MLP_Fea", [64, 64, 64], 1024, 64),            # BASELINE config 1 shape
    ("ndc", "MLP_Fea", [37, 41, 29], 257, 131),            # ragged: odd grid, N and S not tile multiples
    ("contract", "MLP_Fea_TimeEmbedding", [32, 33, 35], 300, 97),
    ("ndc", "MLP_Fea", [20, 20, 20], 3, 1),                # S = 1
])
def test_native_vs_oracle(ray_type, static_mode, grid, n, s):
    case = H.build_pair(grid, n, s, ray_type=ray_type, static_mode=static_mode, seed=n + s)
    rep = H.compare_pass(H.oracle_pass(case), H.native_pass(case, DEV))
    print(rep)


@gpu
def test_all_invalid_and_all_masked():
    """ray_valid all False (rays outside the aabb) and app_mask all False (zero density)."""
    case = H.build_pair([16, 16, 16], 64, 32, seed=5)
    case["rays"][:, 0] += 10.0  # every sample outside
    ref, got = H.oracle_pass(case), H.native_pass(case, DEV)
    assert not ref["valid"].any() and not got["valid"].any()
    H.compare_pass(ref, got)
    assert float(got["fs"][7].abs().max()) == 0.0 and float(got["fs"][6].abs().max()) == 0.0
    case = H.build_pair([16, 16, 16], 64, 32, seed=6)
    for k in list(case["sd_static"]):
        if k.startswith("density_line"):
            case["sd_static"][k] = torch.zeros_like(case["sd_static"][k])
    ref, got = H.oracle_pass(case), H.native_pass(case, DEV)
    assert float(got["fs"][4].abs().max()) == 0.0 and float(got["fs"][6].abs().max()) == 0.0
    H.compare_pass(ref, got)


@gpu
@pytest.mark.parametrize("ray_type,s", [("ndc", 64), ("ndc", 7), ("contract", 64), ("contract", 33)])
def test_train_mode_depth_vector(ray_type, s):
    """zvec kernels with explicit jitter == the oracle's samplers given the same uniforms (bit-exact)."""
    from dynrf_b200 import lib as L, ops

    lib = L.load()
    g = torch.Generator().manual_seed(s)
    aabb = torch.tensor(H.NDC_AABB if ray_type == "ndc" else H.CONTRACT_AABB)
    nf = (0.0, 1.0) if ray_type == "ndc" else (0.0, 256.0)
    cfg_o = O.FieldCfg(aabb=aabb, near_far=nf)
    rays, _ = H.synth_rays(5, ray_type, g)
    zvec = torch.empty(s, device=DEV)
    if ray_type == "ndc":
        j = torch.rand(1, s, generator=g)
        _, z_ref, _ = O.sample_ray_ndc(cfg_o, rays[:, :3], rays[:, 3:], s, True, jitter=j)
        jd = j[0].to(DEV)
        L.check(lib.dynrf_zvec_ndc(nf[0], nf[1], s, ops._p(jd), ops._p(zvec), None), "zvec")
    else:
        n_in, n_out = s - s // 2, s // 2
        ji, jo = torch.rand(1, n_in + 1, generator=g), torch.rand(1, n_out + 1, generator=g)
        _, z_ref, _ = O.sample_ray_contracted(cfg_o, rays[:, :3], rays[:, 3:], s, True, jitter_inner=ji, jitter_outer=jo)
        jid, jod = ji[0].to(DEV), jo[0].to(DEV)
        L.check(lib.dynrf_zvec_contract(nf[0], nf[1], s, ops._p(jid), ops._p(jod), ops._p(zvec), None), "zvec")
    torch.cuda.synchronize()
    torch.testing.assert_close(zvec.cpu(), z_ref[0], rtol=1e-6, atol=1e-7)


@gpu
def test_fused_generation_matches_materialised():
    """xyz == NULL (samples generated in-kernel) gives the same sigma/weight as the materialised path."""
    from dynrf_b200 import ops

    case = H.build_pair([48, 52, 40], 200, 77, seed=11)
    static, dyn = H.native_models(case, DEV)
    rays = case["rays"].to(DEV)
    cfg = static.native_cfg("ndc")
    zvec = ops.make_zvec(cfg, 77, "ndc", False, rays.device)
    xyz, z, valid = ops.sample_points(cfg, rays, zvec)
    a = ops.static_forward_raw(static, rays, xyz, z, valid, "ndc")
    b = ops.static_forward_raw(static, rays, None, None, None, "ndc", zvec=zvec)
    torch.cuda.synchronize()
    for k in ("sigma", "weight", "dists"):
        assert torch.equal(a[k], b[k]), k
    assert int(a["cnt"]) == int(b["cnt"])


@gpu
def test_octree_render_signature():
    import renderer as R

    case = H.build_pair([16, 16, 16], 50, 16, seed=2)
    static, dyn = H.native_models(case, DEV)
    rays, ts = case["rays"].to(DEV), case["ts"].to(DEV)
    with torch.no_grad():
        xyz, z, valid = R.sampleXYZ(dyn, rays, 16, "ndc", False)
        out = R.OctreeRender_trilinear_fast(rays, ts, None, dyn, xyz, z, valid, chunk=16, N_samples=16, device=DEV)
        full = dyn(rays, ts, None, xyz, z, valid, ray_type="ndc")
    assert len(out) == 11 and out[0] is None and out[6] is None
    torch.testing.assert_close(out[2], full[2])
    torch.testing.assert_close(out[5], full[5] - xyz)
    with torch.no_grad():
        out_s = R.OctreeRender_trilinear_fast(rays, ts, None, static, xyz, z, valid, chunk=32, N_samples=16, device=DEV)
    assert out_s[2] is None and out_s[5] is None and out_s[8].shape == (50, 16)


@gpu
def test_tensor_core_path_matches_simt_path():
    """tcgen05 (split-TF32, TMEM) kernels vs the fp32 SIMT kernels of the same stages: same sigma / blending / xyz_prime."""
    from dynrf_b200 import ops

    case = H.build_pair([37, 41, 29], 300, 77, seed=9)
    models = H.native_models(case, DEV)
    try:
        ops.USE_TC = True
        a = H.native_pass(case, DEV, models)
        ops.USE_TC = False
        b = H.native_pass(case, DEV, models)
    finally:
        ops.USE_TC = True
    for i, name in ((7, "sigma"), (2, "blending"), (5, "xyz_prime"), (4, "weight")):
        d = float((a["fd"][i] - b["fd"][i]).abs().max())
        print(name, d)
        assert d <= 5e-6, (name, d)


@gpu
@pytest.mark.parametrize("ray_type", ["ndc", "contract"])
def test_softplus_density_activation(ray_type):
    """fea2denseAct="softplus" (the constructor default of the reference; the shipped configs use relu)."""
    case = H.build_pair([28, 30, 26], 200, 72, ray_type=ray_type, static_mode="MLP_Fea" if ray_type == "ndc" else "MLP_Fea_TimeEmbedding",
                        seed=31, act="softplus", density_shift=-1.5)
    rep = H.compare_pass(H.oracle_pass(case), H.native_pass(case, DEV))
    print(rep)


@gpu
def test_warp_forward_versions_agree():
    """k_warp_fwd_tc2 (warp-specialised, two tiles in flight, own K order) vs the first-generation k_warp_tc: same xyz_prime."""
    from dynrf_b200 import ops

    case = H.build_pair([33, 29, 31], 260, 90, seed=12)
    models = H.native_models(case, DEV)
    try:
        ops.TC_WARP_VERSION = 2
        a = H.native_pass(case, DEV, models)
        ops.TC_WARP_VERSION = 1
        b = H.native_pass(case, DEV, models)
    finally:
        ops.TC_WARP_VERSION = 2
    d = float((a["fd"][5] - b["fd"][5]).abs().max())
    print("xyz_prime", d)
    assert d <= 2e-6, d
