"""Data parallelism over NCCL (NVLink 5 / NVSwitch): the replacement for the reference's Horovod
calls -- hvd.init / hvd.broadcast_parameters / hvd.broadcast_optimizer_state /
hvd.DistributedOptimizer (bin/train_se.py:95-134) and DistributedSampler(num_replicas=hvd.size(),
rank=hvd.rank()) (data/dataloader.py:45-53,83-91).

One process per GPU (torchrun), utterances sharded across ranks, no tensor/pipeline parallelism
(the reference has none).  Gradients are averaged with ONE flat-bucket all-reduce per step: the
21 M-parameter model is 84 MB fp32, which NCCL moves in ~0.2 ms over NVSwitch (NVLS when
available), so bucketing for overlap buys nothing next to the ~30 ms step.  Order of operations is
allreduce -> clip -> step (the reference clips before Horovod's synchronize(), a latent race:
SURVEY.md section 2.3 note).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from torchrun's environment.  Returns (rank, world_size, local_rank)."""
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
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def size():
    return dist.get_world_size() if dist.is_initialized() else 1


def broadcast_parameters(module_or_state, root_rank=0):
    """hvd.broadcast_parameters equivalent: every tensor of the state_dict from root."""
    if size() == 1:
        return
    state = module_or_state.state_dict() if hasattr(module_or_state, "state_dict") else module_or_state
    for _, t in sorted(state.items()):
        if torch.is_tensor(t):
            dist.broadcast(t, src=root_rank)


def broadcast_optimizer_state(optimizer, root_rank=0):
    """hvd.broadcast_optimizer_state equivalent (tensor entries of the optimizer state)."""
    if size() == 1:
        return
    for group in optimizer.param_groups:
        for p in group["params"]:
            st = optimizer.state.get(p, {})
            for k in sorted(st):
                if torch.is_tensor(st[k]) and st[k].numel() > 0:
                    dist.broadcast(st[k], src=root_rank)


class GradAverager(object):
    """Flat-bucket gradient all-reduce (mean) for a fixed parameter list."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self._flat = None

    def average(self):
        if size() == 1:
            return
        grads = [p.grad for p in self.params]
        if any(g is None for g in grads):
            for p in self.params:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            grads = [p.grad for p in self.params]
        n = sum(g.numel() for g in grads)
        if self._flat is None or self._flat.numel() != n or self._flat.device != grads[0].device:
            self._flat = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
        views = []
        off = 0
        for g in grads:
            v = self._flat[off:off + g.numel()].view_as(g)
            views.append(v)
            off += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._flat.div_(size())
        torch._foreach_copy_(grads, views)


class DistributedOptimizer(object):
    """hvd.DistributedOptimizer stand-in: ``step()`` averages gradients across ranks first.
    Call ``synchronize()`` explicitly before gradient clipping (what our trainers do)."""

    def __init__(self, optimizer, named_parameters=None):
        self._opt = optimizer
        params = [p for g in optimizer.param_groups for p in g["params"]]
        self._avg = GradAverager(params)
        self._synced = False

    def __getattr__(self, name):
        return getattr(self._opt, name)

    def zero_grad(self, *a, **k):
        self._synced = False
        return self._opt.zero_grad(*a, **k)

    def synchronize(self):
        if not self._synced:
            self._avg.average()
            self._synced = True

    def step(self, *a, **k):
        self.synchronize()
        self._synced = False
        return self._opt.step(*a, **k)


def shard_indices(n, world, rank_):
    """Contiguous-stride sharding of n items, padded by wrap-around like DistributedSampler."""
    per = (n + world - 1) // world
    idx = [(rank_ + i * world) % n for i in range(per)]
    return idx
