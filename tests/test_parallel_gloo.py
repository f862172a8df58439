"""N>1 host logic on CPU: world_size-2 gloo processes.  Gradient bucket all-reduce == single-process gradient
over the union of the shards; scene sharding covers every scene once."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(8, 5, 6, generator=g), torch.randn(8, 5, 3, generator=g)   # 8 "scenes"


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from panopticsegforlargescalepointcloud_b200 import parallel
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    net = _net()
    if rank == 1:   # broadcast must overwrite a diverged replica
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)

    class M(torch.nn.Module):
        pass
    bucket = parallel.FlatGradBucket(net)
    with torch.no_grad():
        for t in net.parameters():
            dist.broadcast(t, src=0)
    X, Y = _data()
    mine = parallel.shard_scenes(8, rank, world)
    bucket.zero()
    loss = sum(((net(X[i]) - Y[i]) ** 2).mean() for i in mine) / len(mine)
    loss.backward()
    bucket.check_views()
    bucket.all_reduce_mean()
    torch.save({"flat": bucket.flat.clone(), "mine": mine}, out % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process(tmp_path):
    world, port = 2, 29611
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = [torch.load(out % r) for r in range(world)]
    assert sorted(res[0]["mine"] + res[1]["mine"]) == list(range(8))
    assert torch.equal(res[0]["flat"], res[1]["flat"])
    net = _net()
    X, Y = _data()
    loss = sum(((net(X[i]) - Y[i]) ** 2).mean() for i in range(8)) / 8
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(res[0]["flat"], ref, atol=1e-6, rtol=1e-5)


def test_bucket_views_survive_zero_grad():
    sys.path.insert(0, ROOT)
    from panopticsegforlargescalepointcloud_b200 import parallel
    net = _net()
    b = parallel.FlatGradBucket(net)
    opt = torch.optim.Adam(net.parameters())
    opt.zero_grad(set_to_none=True)
    net(torch.randn(4, 6)).sum().backward()
    b.check_views()
    assert all(p.grad.data_ptr() >= b.flat.data_ptr() for p in net.parameters())
    assert float(b.flat.abs().sum()) > 0
    assert parallel.shard_scenes(5, 1, 2) == [1, 3]


def test_data_parallel_step_hooks_zero_and_reattach():
    """DataParallelStep installs the flat-buffer zero hook and the post-backward hook on the model; gradients written
    in place into `param.grad` (what the fused executor's kernels do) land in the flat buffer."""
    sys.path.insert(0, ROOT)
    from panopticsegforlargescalepointcloud_b200 import parallel

    class Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = _net()
            self._grad_hook = None

    m = Model()
    dp = parallel.DataParallelStep(m, broadcast=False)
    assert m._grad_hook is not None and m._zero_grad_hook is not None
    for p in m.parameters():
        p.grad.add_(1.0)                       # in-place accumulation, like the dW / BN kernels
    assert float(dp.bucket.flat.sum()) == sum(p.numel() for p in m.parameters())
    m._zero_grad_hook()
    assert float(dp.bucket.flat.abs().sum()) == 0.0 and all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())
    first = next(m.parameters())
    first.grad = torch.ones_like(first)        # something replaced a view: the hook copies it back and re-attaches
    m._grad_hook()
    assert first.grad.data_ptr() == dp.bucket.views[0].data_ptr() and float(dp.bucket.flat.sum()) == first.numel()
