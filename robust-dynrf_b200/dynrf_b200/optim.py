"""FlatAdam: torch.optim.Adam semantics (no weight decay / amsgrad) as ONE kernel launch over a flat gradient buffer.

The train step keeps every gradient in one flat fp32 buffer that each p.grad aliases with the parameter's own dense strides
(channel-last planes included), so storage element i of a parameter pairs with flat_grad[offset + i].  Capturable by
construction: the step counter lives on the device (train.py:934, 2313-2325 use torch.optim.Adam)."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Sequence, Union

import torch

from . import lib as L
from .ops import _p, _stream


def dense_storage(p: torch.Tensor) -> bool:
    """True if p's strides are a permutation of a dense layout over numel() elements."""
    if p.is_contiguous():
        return True
    exp = 1
    for st, sz in sorted(((st, sz) for sz, st in zip(p.shape, p.stride()) if sz > 1)):
        if st != exp:
            return False
        exp *= sz
    return True


_ALIGN = 64   # floats: every view starts on a 256-byte boundary (the scatter kernels use 128-bit vector reductions on p.grad)


def flat_offsets(params: Sequence[torch.Tensor]) -> List[int]:
    """Start of every parameter in the flat buffer (+ total length as the last entry)."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
    return offs + [off]


def alias_grads(params: Sequence[torch.nn.Parameter]) -> torch.Tensor:
    """One flat zero buffer; every p.grad becomes a view of it with p's own strides (each view 256-byte aligned; the padding
    stays zero and is ignored by FlatAdam and harmless in an all-reduce).  -> the flat buffer."""
    dev = params[0].device
    offs = flat_offsets(params)
    flat = torch.zeros(offs[-1], device=dev)
    for p, off in zip(params, offs):
        if not dense_storage(p):
            raise ValueError("FlatAdam needs densely stored parameters")
        p.grad = torch.as_strided(flat, p.shape, p.stride(), off)
    return flat


class FlatAdam:
    def __init__(self, params: Sequence[torch.nn.Parameter], flat_grad: torch.Tensor, lr: Union[float, Sequence[float]] = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8):
        lib = L.load()
        self.params = list(params)
        self.flat_grad = flat_grad
        dev = flat_grad.device
        T = len(self.params)
        offs, blks = flat_offsets(self.params), [0]
        for p, off in zip(self.params, offs):
            if p.dtype != torch.float32 or not dense_storage(p):
                raise ValueError("FlatAdam: fp32, densely stored parameters only")
            if p.grad is None or p.grad.data_ptr() != flat_grad.data_ptr() + 4 * off:
                raise ValueError("FlatAdam: p.grad must be the view optim.alias_grads() created")
            blks.append(blks[-1] + lib.dynrf_adam_blocks(C.c_longlong(p.numel())))
        assert offs[-1] == flat_grad.numel()
        self.T, self.total_blocks = T, blks[-1]
        self.ptrs = torch.tensor([p.data_ptr() for p in self.params], dtype=torch.int64, device=dev)
        # the kernel needs [start, end) of every parameter: ends differ from the next start by the alignment padding
        self.offset = torch.tensor([v for p, off in zip(self.params, offs) for v in (off, off + p.numel())], dtype=torch.int64, device=dev)
        self.blk0 = torch.tensor(blks, dtype=torch.int32, device=dev)
        lrs = [float(lr)] * T if not isinstance(lr, (list, tuple)) else [float(x) for x in lr]
        self.lr = torch.tensor(lrs, dtype=torch.float32, device=dev)
        self.m = torch.zeros_like(flat_grad)
        self.v = torch.zeros_like(flat_grad)
        self.state = torch.zeros(2, dtype=torch.float32, device=dev)     # [update count, kernel scratch]
        self.betas, self.eps = betas, eps

    def step(self):
        from . import ops

        # the kernel moves the weights without touching their autograd version counters: every packed-operand cache
        # (ops.PackedLayer & co.) keys on this epoch, so bump it -- also BEFORE a CUDA-graph capture, so that the pack kernels
        # of the next forward are part of the captured step
        ops.PARAM_EPOCH[0] += 1
        L.check(L.load().dynrf_adam_flat(_p(self.ptrs), _p(self.offset), _p(self.blk0), self.T, self.total_blocks, _p(self.flat_grad),
                                         _p(self.m), _p(self.v), _p(self.state), _p(self.lr), float(self.betas[0]), float(self.betas[1]),
                                         float(self.eps), _stream()), "dynrf_adam_flat")

    def zero_grad(self):
        self.flat_grad.zero_()
