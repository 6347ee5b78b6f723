"""dynrf_b200 -- B200-native (sm_100a) ray-march hot path for RoDynRF's TensoRF fields.

Sub-modules: lib (ctypes binding of libdynrf_b200.so), ops (torch-facing wrappers),
autograd (torch.autograd.Function glue for the backward kernels).
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
