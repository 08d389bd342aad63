"""world_size-2 gloo test (CPU) of the N>1 path's host logic: edit sharding and the single
pre-loop broadcast of the cached original-image state.  The per-step path has no
collective, so this is all the distributed logic there is."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import warnings

    from sige_b200.parallel import broadcast_caches, cache_tensors, shard_edits
    from sige_b200.workloads.ddpm import DDPMConfig, SIGEDDPMUNet, init_deterministic, synthetic_inputs

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = DDPMConfig.small()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval()
        # rank 0 sees the true original; the other rank starts from a DIFFERENT image, so its caches are
        # wrong until the broadcast overwrites them
        x0, _, _, t = synthetic_inputs(cfg, 0.05, seed=0 if rank == 0 else 7)
        with torch.no_grad():
            model.set_mode("full")
            model(x0, t)
        before = [v.clone() for _, v in cache_tensors(model)]
        nbytes = broadcast_caches(model, src=0)
        after = cache_tensors(model)
        digest = torch.stack([v.double().sum() for _, v in after])
        gathered = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        changed = any(not torch.equal(a, b[1]) for a, b in zip(before, after))
        q.put((rank, nbytes, len(after), same, changed, shard_edits(8, rank, world)))
    finally:
        dist.destroy_process_group()


def test_broadcast_caches_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, k0, same0, changed0, e0), (r1, n1, k1, same1, changed1, e1) = res
    assert n0 == n1 > 0 and k0 == k1 > 20
    assert same0 and same1, "after the broadcast every rank holds rank 0's caches"
    assert not changed0 and changed1, "rank 0 keeps its caches; rank 1's were overwritten"
    assert e0 == [0, 2, 4, 6] and e1 == [1, 3, 5, 7]
