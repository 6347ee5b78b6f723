"""Run under torchrun with 2 ranks (tests/test_gpu_multi.py): the product kernels sharded over NCCL ranks reproduce the
unsharded result -- gathered per-ray maps bit-exact, all-reduced flat gradient equal to the single-process gradient."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from tests import helpers as H
    from dynrf_b200 import train_step as TS
    from dynrf_b200.parallel import shard_bounds

    n, s = 512, 64
    case = H.build_pair([40, 44, 36], n, s, ray_type="ndc", seed=17)
    g = torch.Generator().manual_seed(3)
    target = torch.rand(n, 3, generator=g)
    lo, hi = shard_bounds(n, rank, world)

    def run(rays, ts, tgt, w):
        static, dyn = H.native_models(case, dev)
        step = TS.Step(static, dyn, s, "ndc", "train", w, graph=False, jitter=False)
        out = step(rays.to(dev), ts.to(dev), tgt.to(dev))
        torch.cuda.synchronize()
        return out, step.flat_grad.clone()

    out_sh, grad_sh = run(case["rays"][lo:hi], case["ts"][lo:hi], target[lo:hi], world)
    ok = True
    if rank == 0:
        out_full, grad_full = run(case["rays"], case["ts"], target, 1)
        e_out = float((out_sh - out_full).abs().max())
        rel = float((grad_sh - grad_full).norm() / (grad_full.norm() + 1e-20))
        emax = float((grad_sh - grad_full).abs().max() / (grad_full.abs().max() + 1e-20))
        print(f"NCCL_SHARD_CHECK maps_maxabs={e_out:.3e} grad_rel_l2={rel:.3e} grad_maxnorm={emax:.3e}", flush=True)
        # the maps are per-ray and must not depend on the sharding at all; gradients differ only by the summation order of the
        # atomics and of the all-reduce
        ok = e_out == 0.0 and rel <= 1e-5 and emax <= 1e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
