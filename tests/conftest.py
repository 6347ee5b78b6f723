import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One tests/golden/*.npz case produced by tests/golden/make_golden.py from the reference."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        self.raw = {k: z[k] for k in z.files}
        self.meta = ast.literal_eval(str(self.raw["meta"])) if "meta" in self.raw else {}

    def t(self, key):
        a = self.raw[key]
        return torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape)   # ascontiguousarray promotes 0-d to 1-d

    def sd(self, which):
        pre = f"sd_{which}/"
        return {k[len(pre):]: torch.from_numpy(np.ascontiguousarray(v)) for k, v in self.raw.items() if k.startswith(pre)}

    def cfgs(self):
        from oracle import rodynrf_oracle as O

        aabb = self.t("aabb")
        nf = tuple(float(v) for v in self.raw["near_far"])
        cs = O.FieldCfg(aabb=aabb, near_far=nf, shadingMode=self.meta["static_mode"], fea_pe=2, dynamic=False)
        cd = O.FieldCfg(aabb=aabb, near_far=nf, shadingMode="MLP_Fea_late_view", fea_pe=0, dynamic=True)
        return cs, cd


@pytest.fixture(scope="session", params=["ndc_small", "contract_small"])
def golden_case(request):
    return Golden(request.param)


@pytest.fixture(scope="session")
def golden_pure():
    return Golden("pure_fns")
