"""two bag-sharded ranks of the REAL model (DeformPathomicNet on the sm_100a kernels), gradients after
the flat all-reduce against a single-process run over both bags.  The two ranks share cuda:0 when the box has one GPU (gloo
moves the flat buffer; NCCL refuses two ranks on one device) - the data path has no collective, so this exercises exactly what
the N-GPU job does per step: shard the bags, backward locally, one flat all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dml_b200 import parallel, synth
from dml_b200.model import Args, bag_loss, define_net
from tests import helpers as H

pytestmark = pytest.mark.gpu
WORLD, N, SEED = 2, 700, 71


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _net(dev):
    net = define_net(Args(task_type="survival"))
    net.load_state_dict(synth.fill_like(H.pathomic_shapes(), SEED), strict=True)
    return net.to(dev).eval()


def _grads(net, bag, sl, dev):
    b = {k: v[sl].to(dev) for k, v in bag.items()}
    out = net(x_path=b["x_path"], x_omic_tumor=b["x_omic_tumor"], x_omic_immune=b["x_omic_immune"])
    loss = bag_loss(out[3], b["label_surv"], "survival", b["censor"])
    loss.backward()
    return loss


def _worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        net = _net(dev)
        bag = synth.synthetic_bag(N, SEED, B=WORLD)
        mine = parallel.shard_bags(WORLD, rank, WORLD, shuffle=False)            # one bag per rank
        loss = _grads(net, bag, slice(mine[0], mine[0] + 1), dev)
        red = parallel.FlatGradAllReducer(net.parameters())
        red.allreduce()
        torch.cuda.synchronize()
        q.put((rank, float(loss), {k: (None if p.grad is None else p.grad.detach().cpu().numpy()) for k, p in net.named_parameters()}))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_the_single_process_batch():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(WORLD):
        r = q.get(timeout=600)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    net = _net(dev)
    loss = _grads(net, synth.synthetic_bag(N, SEED, B=WORLD), slice(0, WORLD), dev)          # both bags, mean loss
    assert abs(0.5 * (res[0][0] + res[1][0]) - float(loss)) <= 1e-5 * max(1.0, abs(float(loss)))
    checked = 0
    for k, p in net.named_parameters():
        for rank in range(WORLD):
            g = res[rank][1][k]
            if p.grad is None:
                assert g is None, k
                continue
            g = torch.from_numpy(g)
            if k.endswith("rel_pos_bias.mlp.2.bias"):                     # analytically zero: rounding noise on both sides
                continue
            H.assert_close(g, p.grad.cpu(), 2e-3, f"rank {rank} grad {k}", atol=1e-9)
            checked += 1
        a, b = res[0][1][k], res[1][1][k]
        assert (a is None and b is None) or (a == b).all(), f"{k}: ranks disagree after the all-reduce"
    assert checked > 100
