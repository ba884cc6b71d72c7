"""Bag-sharded data parallelism of the hot path: one process per GPU, bags (patients /
slides) are independent units, no data-path collective; one exchange step per iteration.

Replaces, for this path, the reference's DistributedSampler partition (main.py:110-118, 326-341), DDP's bucketed
gradient all-reduce (main.py:190,401) and the per-parameter all-reduce loop of the trainers
(train_test.py:223-227 and siblings: ~100 NCCL calls per step) by ONE flat fp32 all-reduce.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and with it the first-touch placement of the pinned host buffers it allocates next) to the CPUs of
    the NUMA node the GPU hangs off.  With 8 ranks each copying a 33.5 MB bag per step, host buffers on the wrong socket
    put every copy on the inter-socket link (round 1: end-to-end scaling 0.90 against 0.98 device-timed).  Best effort:
    returns what it did; never raises."""
    import os
    info = {"numa_node": None, "cpus": None}
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"numa_node": node, "cpus": len(allowed)}
    except Exception as ex:      # containers without /sys access, single-socket hosts, ...
        info["error"] = repr(ex)[:120]
    return info


def shard_bags(num_bags: int, rank: int, world: int, *, epoch: int = 0, seed: int = 0, shuffle: bool = True,
               drop_last: bool = True) -> List[int]:
    """Indices of the bags this rank processes in `epoch`: the same partition as
    torch.utils.data.DistributedSampler(shuffle, seed) + DataLoader(drop_last) - a seeded permutation dealt
    round-robin, every rank the same count."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        order = torch.randperm(num_bags, generator=g).tolist()
    else:
        order = list(range(num_bags))
    if drop_last:
        order = order[: (num_bags // world) * world]
    else:
        pad = (-len(order)) % world
        order = order + order[:pad]
    return order[rank::world]


def balance_bags_by_cost(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment of variable-length bags to ranks (cost ~ N^2 for the deformable
    attention): returns, per rank, the bag indices of one step."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += float(lengths[i]) ** 2
    return out


class LengthBucketedSampler(torch.utils.data.Sampler):
    """DistributedSampler for variable-length bags: every step hands the `world` ranks bags of NEIGHBOURING lengths (bags sorted
    by length, consecutive groups of `world` = one step, the steps shuffled per epoch), so no rank waits at the gradient
    all-reduce for a neighbour that drew a 16k bag against its 4k one (cost ~ N^2).  Same contract as
    DistributedSampler(drop_last=True): every rank the same number of bags, each bag at most once per epoch."""

    def __init__(self, lengths: Sequence[int], world: int, rank: int, seed: int = 0, shuffle: bool = True):
        self.lengths, self.world, self.rank, self.seed, self.shuffle, self.epoch = list(lengths), world, rank, seed, shuffle, 0
        order = sorted(range(len(self.lengths)), key=lambda i: (-self.lengths[i], i))
        usable = (len(order) // world) * world
        self.steps = [order[i:i + world] for i in range(0, usable, world)]

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self):
        return len(self.steps)

    def __iter__(self):
        idx = list(range(len(self.steps)))
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(len(self.steps), generator=g).tolist()
        # within a step the longest bag rotates over the ranks from step to step
        for n, k in enumerate(idx):
            step = self.steps[k]
            yield step[(self.rank + n) % self.world]


class FlatGradAllReducer:
    """Averages the gradients of `params` across ranks with a single collective on one flat fp32 buffer.
    Parameters without a gradient on this rank (the never-used attn2d.* / pooler.* weights, quirk Q7) contribute
    zeros, so every rank reduces the same layout; parameters that have no gradient on ANY rank keep grad = None
    (checked with the same collective: a presence flag per parameter rides at the end of the buffer)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], static_presence: bool = False):
        """static_presence: the set of parameters that receive gradients does not change from step to step (true
        for a fixed model/mode): it is read back from the device once, later steps stay free of host syncs."""
        self.static_presence = static_presence
        self._present = None
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        # every slot starts on a 128-byte boundary: the flat buffers may also hold the parameters themselves
        # (graph.GraphedTrainStep flat_optimizer), and kernels read weights with 16-byte vector loads
        self.offsets, off = [], 0
        for n in self.sizes:
            self.offsets.append(off)
            off += (n + 31) // 32 * 32
        self.total = off
        p0 = self.params[0]
        self.flat = torch.zeros(self.total + len(self.params), dtype=torch.float32, device=p0.device)

    # ---- attached mode: parameters accumulate their gradients directly into views of the flat buffer ----
    def _views(self, flat):
        return [flat[o: o + n] for o, n in zip(self.offsets, self.sizes)]

    def attach(self) -> None:
        """Call after one backward pass: parameters that received a gradient get p.grad = a view of the flat buffer
        (autograd accumulates in place from then on), the others keep grad = None.  Afterwards use zero_grad() instead
        of optimizer.zero_grad(); allreduce() is then a single collective with no copies."""
        views = self._views(self.flat)
        self._present = [p.grad is not None for p in self.params]
        for p, v, has in zip(self.params, views, self._present):
            if has:
                v.copy_(p.grad.reshape(-1))
                p.grad = v.view_as(p)
        self.flat[self.total:] = torch.tensor([1.0 if h else 0.0 for h in self._present], device=self.flat.device)
        self.attached = True

    def attach_views(self) -> None:
        """Re-point p.grad at this reducer's flat buffer (after another reducer / optimizer call replaced them)."""
        views = self._views(self.flat)
        for p, v, has in zip(self.params, views, self._present):
            p.grad = v.view_as(p) if has else None

    def zero_grad(self) -> None:
        self.flat[: self.total].zero_()

    def allreduce(self, group=None) -> None:
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        flat = self.flat
        if getattr(self, "attached", False):
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(flat[: self.total], op=dist.ReduceOp.AVG, group=group)      # one collective, no second pass
            else:                                                                            # gloo has no AVG
                dist.all_reduce(flat[: self.total], op=dist.ReduceOp.SUM, group=group)
                flat[: self.total].div_(world)
            return
        flat.zero_()
        views = self._views(flat)
        flags = flat[self.total:]
        for i, (p, v) in enumerate(zip(self.params, views)):
            if p.grad is not None:
                v.copy_(p.grad.reshape(-1))
                flags[i] = 1.0
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat[: self.total].div_(world)
        if self.static_presence and self._present is not None:
            present = self._present
        else:
            present = self._present = (flags > 0).tolist()
        for p, v, has in zip(self.params, views, present):
            if not has:
                p.grad = None
            elif p.grad is None:
                p.grad = v.view_as(p).clone()
            else:
                p.grad.copy_(v.view_as(p))


def allreduce_grads_flat(params: Iterable[torch.nn.Parameter], group=None) -> None:
    """One-shot form of FlatGradAllReducer (allocates the flat buffer on each call)."""
    FlatGradAllReducer(params).allreduce(group)
