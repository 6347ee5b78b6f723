"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports exactly the entry
points include/dynrf_b200.h declares (no compute calls -- there is no GPU here)."""
import os
import re

import pytest

from tests import helpers  # noqa: F401  (sys.path)
from dynrf_b200 import lib as L


def _declared():
    hdr = open(os.path.join(helpers.ROOT, "include", "dynrf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dynrf_[a-z0-9_]+)\s*\(", hdr)))


def test_build_and_symbols():
    L.build()
    lib = L.load()
    declared = _declared()
    assert "dynrf_static_sigma_weight_fwd" in declared and len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dynrf_b200.h but not exported"
    assert sorted(set(L.PROTOTYPES) | {"dynrf_strerror"}) == declared
    assert lib.dynrf_abi_version() == 3   # v2: dynrf_vm_t texel strides sv / su; v3: warp-MLP hidden activations kept for the backward
    assert lib.dynrf_pad_k(93) == 96 and lib.dynrf_pad_n(27) == 32
    assert b"argument" in lib.dynrf_strerror(-1)


def test_product_never_imports_oracle():
    pkg = helpers.PKG
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_cpu_call_fails_loudly():
    import torch
    import models.tensoRF as T

    aabb = torch.tensor(helpers.NDC_AABB)
    m = T.TensorVMSplit(aabb, [8, 8, 8], 12, torch.device("cpu"), shadingMode="MLP_Fea", fea_pe=2,
                        **dict(helpers.COMMON, near_far=[0.0, 1.0]))
    rays = torch.zeros(4, 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample_ray_ndc(rays[:, :3], rays[:, 3:], is_train=False, N_samples=8)


def test_state_dict_matches_reference_keys(golden_case):
    """Keys and shapes of the drop-in modules equal the reference's state_dict (from the golden dump)."""
    import torch

    case = helpers.build_pair(list(golden_case.raw["grid"]), 4, 8, ray_type=golden_case.meta["ray_type"],
                              static_mode=golden_case.meta["static_mode"], sd_static=golden_case.sd("static"),
                              sd_dynamic=golden_case.sd("dynamic"), aabb=golden_case.t("aabb"),
                              near_far=tuple(float(v) for v in golden_case.raw["near_far"]))
    import models.tensoRF as T

    kw = dict(helpers.COMMON, near_far=list(case["near_far"]))
    static = T.TensorVMSplit(case["aabb"], case["grid"], 12, torch.device("cpu"), shadingMode=case["static_mode"], fea_pe=2, **kw)
    dyn = T.TensorVMSplit_TimeEmbedding(case["aabb"], case["grid"], 12, torch.device("cpu"), shadingMode="MLP_Fea_late_view",
                                        fea_pe=0, **kw)
    for m, sd in ((static, case["sd_static"]), (dyn, case["sd_dynamic"])):
        mine = m.state_dict()
        assert list(mine.keys()) == list(sd.keys())
        for k in sd:
            assert tuple(mine[k].shape) == tuple(sd[k].shape), k
        m.load_state_dict(sd)
    # channel-last storage, logical NCHW shape
    p = static.density_plane[0]
    assert p.shape[1] == 16 and p.stride(1) == 1


def test_bench_workload_is_the_test_helpers_model():
    """bench.py's native arm builds its models with dynrf_b200/synth.py (no test infrastructure on that path); the CPU arm and the
    tests use tests/helpers.build_pair.  Both must be the same model (same generator, same draw order)."""
    import torch

    from dynrf_b200 import synth

    for rt, mode, grid in (("ndc", "MLP_Fea", [21, 24, 18]), ("contract", "MLP_Fea_TimeEmbedding", [16, 19, 20])):
        case = helpers.build_pair(grid, 8, 8, ray_type=rt, static_mode=mode, seed=1)
        gen = torch.Generator().manual_seed(1)
        a = synth.init_state_dict(False, mode, grid, gen)
        b = synth.init_state_dict(True, mode, grid, gen)
        assert a.keys() == case["sd_static"].keys() and b.keys() == case["sd_dynamic"].keys()
        assert all(torch.equal(a[k], case["sd_static"][k]) for k in a)
        assert all(torch.equal(b[k], case["sd_dynamic"][k]) for k in b)
