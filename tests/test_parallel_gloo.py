"""N > 1 path on CPU: world_size-2 gloo processes (no GPU) - bag sharding, the flat gradient all-reduce against a
single-process run over the concatenated batch, and GatherLayer forward / backward (backward = local slice,
reference utils/gather.py:16-20)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dml_b200 import parallel
from dml_b200.gather import GatherLayer

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.body = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        self.unused = torch.nn.Linear(4, 4)     # never receives a gradient (like attn2d.* / pooler.*)

    def forward(self, x):
        return self.body(x)


def _tiny_net():
    torch.manual_seed(3)
    return _Tiny()


def _worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        # ---- sharding: ranks partition the bags, same count each ----
        mine = parallel.shard_bags(11, rank, WORLD, epoch=2, seed=42)
        # ---- flat gradient all-reduce ----
        net = _tiny_net()
        torch.manual_seed(100)
        x = torch.randn(8, 6)
        y = torch.randn(8, 3)
        xs, ys = x[rank::WORLD], y[rank::WORLD]
        loss = ((net(xs) - ys) ** 2).mean()
        loss.backward()
        red = parallel.FlatGradAllReducer(net.parameters())
        red.allreduce()
        grads = {k: (None if p.grad is None else p.grad.clone()) for k, p in net.named_parameters()}
        # ---- GatherLayer ----
        t = (torch.arange(6, dtype=torch.float32).reshape(2, 3) + 10 * rank).requires_grad_()
        parts = GatherLayer.apply(t)
        full = torch.cat(parts, dim=0)
        w = torch.arange(full.numel(), dtype=torch.float32).reshape(full.shape) * (rank + 1)
        (full * w).sum().backward()
        # plain numpy payloads: tensors on an mp.Queue are shared through the sender's resource sharer, which is gone
        # if this worker exits before the parent unpickles them (an intermittent FileNotFoundError in the parent)
        q.put((rank, mine, {k: (None if v is None else v.numpy()) for k, v in grads.items()}, full.detach().numpy().copy(),
               t.grad.numpy().copy(), w[2 * rank: 2 * rank + 2].numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_world2_sharding_allreduce_gather():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(WORLD):
        r = q.get(timeout=180)                  # a crashed worker fails the test instead of hanging it
        res[r[0]] = (r[1], {k: (None if v is None else torch.from_numpy(v)) for k, v in r[2].items()},
                     torch.from_numpy(r[3]), torch.from_numpy(r[4]), torch.from_numpy(r[5]))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0

    # sharding: disjoint, equal counts, union = the permutation with the tail dropped
    a, b = res[0][0], res[1][0]
    assert len(a) == len(b) == 5 and not set(a) & set(b)
    g = torch.Generator()
    g.manual_seed(42 + 2)
    assert sorted(a + b) == sorted(torch.randperm(11, generator=g).tolist()[:10])

    # all-reduce: mean of the two half-batch gradients == gradient of the full batch (equal shard sizes)
    net = _tiny_net()
    torch.manual_seed(100)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    ((net(x) - y) ** 2).mean().backward()
    for rank in range(WORLD):
        grads = res[rank][1]
        for k, p in net.named_parameters():
            if p.grad is None:
                assert grads[k] is None, k              # unused everywhere -> stays None
            else:
                assert torch.allclose(grads[k], p.grad, rtol=1e-5, atol=1e-7), k
    for k in res[0][1]:
        g0, g1 = res[0][1][k], res[1][1][k]
        assert (g0 is None and g1 is None) or torch.equal(g0, g1)       # bit-identical on every rank

    # GatherLayer: forward = all ranks' tensors in rank order; backward = this rank's slice of ITS OWN upstream grads
    exp = torch.cat([torch.arange(6, dtype=torch.float32).reshape(2, 3) + 10 * r for r in range(WORLD)], 0)
    for rank in range(WORLD):
        _, _, full, tgrad, wslice = res[rank]
        assert torch.equal(full, exp)
        assert torch.equal(tgrad, wslice)


def _bn_net():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 3))


def _bn_worker(rank, port, q):
    from dml_b200.sync_bn import convert_sync_batchnorm
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        net = convert_sync_batchnorm(_bn_net()).train()
        torch.manual_seed(200)
        x, y = torch.randn(10, 6), torch.randn(10, 3)
        sl = slice(0, 3) if rank == 0 else slice(3, 10)              # UNEQUAL shards: the counts must ride along
        xs = x[sl].clone().requires_grad_()
        out = net(xs)
        # the single-process loss is a mean over all 10 rows: weight each rank's mean by its share
        loss = ((out - y[sl]) ** 2).sum() / (10 * 3)
        loss.backward()
        for p in net.parameters():                                   # SUM (the loss above is already globally normalised)
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
        q.put((rank, out.detach().numpy().copy(), xs.grad.numpy().copy(),
               {k: p.grad.numpy().copy() for k, p in net.named_parameters()},
               {k: v.numpy().copy() for k, v in net.state_dict().items() if "running" in k or "num_batches" in k}))
    finally:
        dist.destroy_process_group()


def test_world2_sync_batchnorm_statistics_match_the_global_batch():
    """SyncBatchNorm1d over 2 gloo ranks with 3 + 7 rows == nn.BatchNorm1d over the 10-row batch: outputs, input gradients,
    parameter gradients, running statistics (reference: SyncBatchNorm conversion at main.py:189,400)."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bn_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(WORLD):
        r = q.get(timeout=180)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    net = _bn_net().train()
    torch.manual_seed(200)
    x, y = torch.randn(10, 6), torch.randn(10, 3)
    xr = x.clone().requires_grad_()
    out = net(xr)
    ((out - y) ** 2).mean().backward()
    got_out = torch.cat([torch.from_numpy(res[0][0]), torch.from_numpy(res[1][0])])
    got_gx = torch.cat([torch.from_numpy(res[0][1]), torch.from_numpy(res[1][1])])
    assert torch.allclose(got_out, out.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(got_gx, xr.grad, rtol=1e-4, atol=1e-6)
    for k, p in net.named_parameters():
        for rank in range(WORLD):
            assert torch.allclose(torch.from_numpy(res[rank][2][k]), p.grad, rtol=1e-4, atol=1e-6), k
    for k, v in net.state_dict().items():
        if "running" in k or "num_batches" in k:
            for rank in range(WORLD):
                assert torch.allclose(torch.from_numpy(res[rank][3][k]).float(), v.float(), rtol=1e-5, atol=1e-6), k


def test_lpt_balancing_of_variable_length_bags():
    lengths = [16384, 4096, 8192, 12000, 6000, 15000, 5000, 9000]
    parts = parallel.balance_bags_by_cost(lengths, 4)
    assert sorted(i for p in parts for i in p) == list(range(8))
    loads = [sum(lengths[i] ** 2 for i in p) for p in parts]
    assert max(loads) <= 1.35 * (sum(loads) / 4)
