"""CPU, world_size 2, gloo: the data-parallel host logic - scene sharding and the single flat
gradient all-reduce (the only collective on the training path)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cns_b200.train import GradBucket, shard_range


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))     # same init on every rank
    bucket = GradBucket(list(net.parameters()))
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    lo, hi = shard_range(8, rank, world)
    bucket.zero()
    net(data[lo:hi]).pow(2).mean().backward()
    bucket.rebind()
    bucket.allreduce_mean()
    ret[rank] = bucket.flat.clone()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n, w in ((10, 3), (8, 2), (1024, 8), (5, 8)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_flat_bucket_allreduce_equals_full_batch_gradient():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29533, ret), nprocs=world, join=True)
    assert torch.equal(ret[0], ret[1])                          # every rank ends with the same bucket
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10.0
    net(data).pow(2).mean().backward()                          # equal shards: mean of means == global mean
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(ret[0], full, atol=1e-6)
