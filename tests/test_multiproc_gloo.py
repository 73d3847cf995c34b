"""world_size-2 gloo test of the N>1 host logic: rank discovery (ddp_setup), batch sharding for sampling and the
flat-gradient mean all-reduce that FusedAdamW performs (sum all-reduce, 1/world folded into the update)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update({"PMI_SIZE": str(world), "PMI_RANK": str(rank), "MPI_LOCALRANKID": str(rank),
                       "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    from rho_diffusion.b200.sampler import shard_range
    from rho_diffusion.utils import ddp_setup
    w, r, lr = ddp_setup(port=port, backend="gloo")
    assert (w, r, lr) == (world, rank, rank)
    # flat gradient buffer: rank-local values 2.0 and 4.0 -> mean 3.0 on both (the SURVEY probe of the intended semantics)
    flat_g = torch.full((1000,), 2.0 * (rank + 1))
    dist.all_reduce(flat_g, op=dist.ReduceOp.SUM)
    flat_g *= 1.0 / w
    lo, hi = shard_range(101, w, r)
    # bucketed, asynchronously launched all-reduces over slices of one flat buffer (what FusedAdamW does from inside the backward,
    # last bucket first) give exactly the whole-buffer all-reduce
    from rho_diffusion.b200 import partition_buckets
    g0 = torch.Generator().manual_seed(100 + rank)
    sizes = [7, 64, 1, 300, 50, 128, 9, 441]
    offsets = [sum(sizes[:i]) for i in range(len(sizes))]
    total = sum(sizes)
    whole = torch.randn(total, generator=g0)
    bucketed = whole.clone()
    dist.all_reduce(whole, op=dist.ReduceOp.SUM)
    ranges, _ = partition_buckets(offsets, total, 3)
    works = [dist.all_reduce(bucketed[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in reversed(ranges)]
    for wk in works:
        wk.wait()
    assert torch.equal(whole, bucketed)
    out[rank] = (float(flat_g[0]), float(flat_g[-1]), lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:2] == (3.0, 3.0) and out[1][:2] == (3.0, 3.0)
    assert (out[0][2], out[0][3], out[1][2], out[1][3]) == (0, 51, 51, 101)


def test_shard_range_partitions():
    from rho_diffusion.b200.sampler import shard_range
    for total in (1, 7, 8, 100, 101):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_partition_buckets_covers_the_flat_buffer_at_parameter_boundaries():
    from rho_diffusion.b200 import partition_buckets
    import random
    rng = random.Random(5)
    for n_params in (1, 2, 5, 40, 329):
        for n_b in (1, 3, 8, 64):
            sizes = [64 * rng.randint(1, 500) for _ in range(n_params)]
            offsets = [sum(sizes[:i]) for i in range(n_params)]
            total = sum(sizes)
            ranges, bucket_of = partition_buckets(offsets, total, n_b)
            assert ranges[0][0] == 0 and ranges[-1][1] == total and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert len(ranges) <= min(n_b, n_params) and len(bucket_of) == n_params and bucket_of == sorted(bucket_of)
            assert set(bucket_of) == set(range(len(ranges)))
            for off, b in zip(offsets, bucket_of):
                assert ranges[b][0] <= off < ranges[b][1]          # cut at parameter boundaries: a parameter lives in ONE bucket
