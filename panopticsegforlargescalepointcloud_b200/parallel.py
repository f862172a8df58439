"""Data parallelism over cylinder samples: one process per GPU, scenes sharded by rank, ONE all-reduce of a
flat fp32 gradient bucket per step (BASELINE.json north_star).

The reference has no distributed code: its Trainer is single-process.  This is the path's only
collective: every scene is independent through hash build, rulebooks, convolutions, clustering and scoring, so
nothing else crosses ranks.  BatchNorm statistics stay per rank (what torch DDP does by default).

Launch: `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N` (reads RANK / LOCAL_RANK /
WORLD_SIZE / MASTER_* from the environment); backend "nccl" on GPUs, "gloo" for the CPU tests.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """-> (rank, world, local_rank).  No-op single-process group when WORLD_SIZE is absent or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        import datetime
        # a rank that leaves the collective pattern must fail fast, not sit in NCCL's default 10-minute watchdog
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("PGS_DIST_TIMEOUT_S", "180"))))
    return rank, world, local


def shard_scenes(n_scenes, rank, world):
    """Round-robin scene ids of this rank (SURVEY 8e 'Partitioning')."""
    return list(range(rank, n_scenes, world))


class FlatGradBucket:
    """All parameter gradients as views into one contiguous fp32 buffer, so that the step's only collective is a
    single all-reduce over `flat` (10.4 M backbone parameters = 41.7 MB)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dt = self.params[0].device, torch.float32
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def check_views(self):
        """Some optimisers / zero_grad(set_to_none=True) drop .grad; re-attach the views."""
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is v:
                continue
            if g is None:
                p.grad = v
            elif g.data_ptr() != v.data_ptr():
                v.copy_(g)
                p.grad = v

    def all_reduce_mean(self, world=None):
        # weight-gradient kernels may still be running on the side stream (me.DW_DIRECT / fastpath.DW_SIDE_STREAM): whoever
        # consumes the bucket waits for them first
        if self.flat.is_cuda:
            from . import me
            me.join_side_stream()
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size() if world is None else world)


class DataParallelStep:
    """Wraps a BaseModel: gradients live in a FlatGradBucket and are averaged across ranks between backward and
    the optimiser step (the model calls `_grad_hook`)."""

    def __init__(self, model, broadcast=True):
        self.model = model
        self.bucket = FlatGradBucket(model)
        if broadcast and dist.is_initialized() and dist.get_world_size() > 1:
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t, src=0)
        model._grad_hook = self._hook
        model._zero_grad_hook = self.bucket.zero   # one memset of the flat buffer instead of one fill per parameter
        self.allreduce_events = None    # set to a list: (start, end) CUDA events around every gradient all-reduce
        self.local_only = False         # True: skip the collective (a single rank running extra, unmatched steps)

    def _hook(self):
        self.bucket.check_views()
        if self.local_only:
            return
        ev = self.allreduce_events
        if ev is not None and self.bucket.flat.is_cuda and dist.is_initialized() and dist.get_world_size() > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.bucket.all_reduce_mean()
            e1.record()
            ev.append((e0, e1))
        else:
            self.bucket.all_reduce_mean()

    def prefetch(self, data, stream=None, wait_event=None):
        """Build the next batch's coordinate maps while the current step runs (panoptic `prefetch_maps`)."""
        return self.model.prefetch_maps(data, stream=stream, wait_event=wait_event)

    def step(self, data, epoch, step=0, batch_size=1):
        self.model.set_input(data, self.model.device)
        self.model.optimize_parameters2(epoch, step, batch_size)
        return self.model.loss.detach()
