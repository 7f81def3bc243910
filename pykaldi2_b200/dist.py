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
    """Flat-bucket gradient all-reduce (mean) for a fixed parameter list.

    The gradients are packed into one flat buffer (one pass), summed by a single all-reduce, scaled, and handed back
    as VIEWS of that buffer: ``p.grad`` points into the bucket afterwards,
    so clipping and the optimizer read the averaged values without a copy back.  ``timers``: set to a list to collect
    (start, end) CUDA events around pack + all-reduce (bench.py reports the mean as ``allreduce_ms``)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self._flat = None
        self.timers = None

    def average(self):
        if size() == 1:
            return
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
        timed = self.timers is not None and self._flat.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._flat.div_(size())
        if timed:
            e1.record()
            self.timers.append((e0, e1))
        for p, v in zip(self.params, views):
            p.grad = v


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


def balanced_shards(lengths, world, per_rank=None, tmax_weight=56.0):
    """Split one global minibatch of utterances across ``world`` data-parallel ranks so that the ranks finish a step
    together.  Step time of a rank is modelled as  tmax_weight * max(length) + sum(length)  (in frames): the BLSTM
    recurrences and the padded GEMMs scale with the LONGEST utterance of the rank's batch (the reference pads to it,
    data/dataloader.py:96-103), the denominator / lattice forward-backward and the output layer with the SUM of
    frames.  tmax_weight = 56 is the least-squares fit of the per-rank compute times of the 1/2/4/8-GPU runs on B200 for
    the 3x512 BLSTM: 22.9 us per padded output frame against 0.41 us per valid output frame (profiles/README_r2.md).

    Longest-first greedy onto the rank with the lowest modelled cost that still has room (``per_rank`` utterances per
    rank, default ceil(n / world)): the rank that receives the longest utterance pays the largest padding term and is
    given fewer frames in exchange.  Deterministic (ties -> lowest rank); every rank computes the same split.
    Returns a list of ``world`` index lists (positions into ``lengths``).  Replaces the random sharding of the
    reference's DistributedSampler (data/dataloader.py:45-53,83-91)."""
    n = len(lengths)
    if per_rank is None:
        per_rank = (n + world - 1) // world
    if per_rank * world < n:
        raise ValueError("balanced_shards: %d utterances do not fit %d ranks x %d" % (n, world, per_rank))
    order = sorted(range(n), key=lambda i: (-float(lengths[i]), i))
    shards = [[] for _ in range(world)]
    tmax = [0.0] * world
    tot = [0.0] * world
    for i in order:
        best, best_cost = -1, None
        for r in range(world):
            if len(shards[r]) >= per_rank:
                continue
            cost = tmax_weight * max(tmax[r], float(lengths[i])) + tot[r] + float(lengths[i])
            if best_cost is None or cost < best_cost:
                best, best_cost = r, cost
        shards[best].append(i)
        tmax[best] = max(tmax[best], float(lengths[i]))
        tot[best] += float(lengths[i])
    return shards


def shard_indices(n, world, rank_):
    """Contiguous-stride sharding of n items, padded by wrap-around like DistributedSampler."""
    per = (n + world - 1) // world
    idx = [(rank_ + i * world) % n for i in range(per)]
    return idx
