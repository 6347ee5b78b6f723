"""N>1 plumbing on CPU: world_size-2 gloo processes shard the rays, run the (oracle) path on their shard,
all-gather the per-ray maps and all-reduce the gradients; both must equal the unsharded result."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from tests import helpers as H
from oracle import rodynrf_oracle as O


def _per_ray(out13):
    o = out13
    return torch.cat((o[0], o[1][:, None], o[2][:, None], o[4], o[5][:, None], o[6][:, None], o[8], o[9][:, None],
                      o[10][:, None], o[12][:, None]), -1)


def _worker(rank, world, port, n_total, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    sys.path.insert(0, H.PKG)
    from dynrf_b200 import parallel as P

    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = H.build_pair([12, 14, 10], n_total, 24, seed=4)
    lo, hi = P.shard_bounds(n_total, rank, world)
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd_dynamic"].items()}
    sub = dict(case, rays=case["rays"][lo:hi], ts=case["ts"][lo:hi], sd_dynamic=sd)
    xyz, z, valid = O.sampleXYZ(sub["cd"], sub["rays"], 24, "ndc", False)
    fs = O.field_forward(sub["cs"], sub["sd_static"], sub["rays"], sub["ts"], xyz, z, valid, "ndc")
    fd = O.field_forward(sub["cd"], sd, sub["rays"], sub["ts"], xyz, z, valid, "ndc")
    out = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], sub["rays"], False, "ndc")
    local = _per_ray(out)
    counts = [P.shard_bounds(n_total, r, world)[1] - P.shard_bounds(n_total, r, world)[0] for r in range(world)]
    full = P.gather_ray_maps(local.detach(), world, counts)
    # per-rank mean loss over its shard, then mean of gradients == gradient of the global mean when shards are equal
    loss = (out[0] ** 2).mean()
    loss.backward()
    params = [torch.nn.Parameter(v.detach().clone()) for v in sd.values()]
    for p, v in zip(params, sd.values()):
        p.grad = None if v.grad is None else v.grad.clone()
    # the step driver's variant: one flat buffer that every gradient aliases
    flat = torch.cat([(torch.zeros_like(v) if v.grad is None else v.grad).reshape(-1) for v in sd.values()])
    P.allreduce_flat_mean(flat, world)
    P.allreduce_mean_grads(params, world)
    if rank == 0:
        torch.save({"full": full, "grads": {k: p.grad for k, p in zip(sd.keys(), params)}, "flat": flat}, tmp)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [32, 33])
def test_two_rank_sharding_matches_single(tmp_path, n_total):
    world, port = 2, 29500 + (os.getpid() % 2000)
    tmp = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, port, n_total, tmp), nprocs=world, join=True)
    got = torch.load(tmp)
    case = H.build_pair([12, 14, 10], n_total, 24, seed=4)
    sd = {k: v.clone().requires_grad_(True) for k, v in case["sd_dynamic"].items()}
    ref = H.oracle_pass(case)
    torch.testing.assert_close(got["full"], _per_ray(ref["out"]), rtol=1e-5, atol=1e-6)
    # flat-buffer all-reduce == per-tensor all-reduce
    flat_ref = torch.cat([(torch.zeros_like(case["sd_dynamic"][k]) if g is None else g).reshape(-1) for k, g in got["grads"].items()])
    torch.testing.assert_close(got["flat"], flat_ref, rtol=0, atol=0)
    if n_total % world == 0:  # equal shards: mean of per-shard mean-loss gradients == full-batch gradient
        xyz, z, valid = O.sampleXYZ(case["cd"], case["rays"], 24, "ndc", False)
        fs = O.field_forward(case["cs"], case["sd_static"], case["rays"], case["ts"], xyz, z, valid, "ndc")
        fd = O.field_forward(case["cd"], sd, case["rays"], case["ts"], xyz, z, valid, "ndc")
        out = O.raw2outputs(fs[6], fs[7], fd[6], fd[7], fd[9], fd[2], fd[8], case["rays"], False, "ndc")
        (out[0] ** 2).mean().backward()
        for k, v in sd.items():
            if v.grad is not None:
                torch.testing.assert_close(got["grads"][k], v.grad, rtol=1e-4, atol=1e-7)


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, H.PKG)
    from dynrf_b200 import parallel as P

    for n in (1, 7, 4096, 4097):
        for w in (1, 2, 3, 8):
            sl = [P.shard_bounds(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            sizes = [b - a for a, b in sl]
            assert max(sizes) - min(sizes) <= 1
