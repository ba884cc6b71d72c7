"""Batch / domain losses (utils/loss.py) on the streaming similarity kernels of csrc/gram.cu: the reference's own goldens at
world_size 1, the oracle at larger row counts and ragged K, and two ranks (local-rows-only gradient, utils/gather.py:16-20)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dml_b200 import loss as L
from dml_b200 import synth
from dml_b200._lib import call, ptr, stream
from oracle import losses as OL
from oracle.golden_cases import LOSS_CASES, loss_inputs, thin
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3          # north_star: fp32 path


@pytest.mark.parametrize("c", LOSS_CASES, ids=lambda c: c["name"])
def test_losses_match_reference_goldens(c):
    G = H.golden(c["name"])
    x = {k: v.to(DEV).requires_grad_() for k, v in loss_inputs(c).items()}
    pb = L.PathBatchLoss(c["N"], 1)(x["a1_10"], x["a1_20"])
    od = L.OmicDomainScaleLoss(c["N"], 1)(x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"])
    bl = L.BatchLoss(c["N"], 1)(x["omic"], x["vgrid"])
    H.assert_close(pb.cpu(), G["path_batch"], TOL, "PathBatchLoss")
    H.assert_close(od.cpu(), G["omic_domain"], TOL, "OmicDomainScaleLoss")
    H.assert_close(bl.cpu(), G["batch"], TOL, "BatchLoss")
    g = torch.autograd.grad(pb.sum(), (x["a1_10"], x["a1_20"]), retain_graph=True)
    H.assert_close(thin(g[0].cpu()), G["pb.g10"], TOL, "d PathBatchLoss / d att10")
    H.assert_close(thin(g[1].cpu()), G["pb.g20"], TOL, "d PathBatchLoss / d att20")
    g = torch.autograd.grad(od, (x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"]), retain_graph=True)
    for k, v in zip(("a1_10", "a1_20", "a2_10", "a2_20"), g):
        H.assert_close(thin(v.cpu()), G["od.g_" + k], 5 * TOL, "d OmicDomainScaleLoss / d " + k)
    g = torch.autograd.grad(bl.sum(), (x["omic"], x["vgrid"]))
    H.assert_close(thin(g[0].cpu()), G["bl.g_omic"], TOL, "d BatchLoss / d omic")
    H.assert_close(thin(g[1].cpu()), G["bl.g_vgrid"], TOL, "d BatchLoss / d vgrid")


@pytest.mark.parametrize("N,G,K", [(4, 8, 2500 * 144), (16, 8, 5004), (33, 1, 40000), (64, 2, 9996), (1, 8, 64), (7, 3, 4)])
def test_similarity_kernels_against_fp64(N, G, K):
    """sim[g] = A[g] B[g]^T and its adjoint on rows read through the pointer table: both tile widths, K tails, one row."""
    a = synth.normal((N, G, K), 5, "a").to(DEV).requires_grad_()
    b = synth.normal((N, G, K), 5, "b").to(DEV).requires_grad_()
    r = synth.normal((G, N, N), 5, "r").to(DEV)
    for same in (False, True):
        if same and N > 16:
            continue                                      # the adjoint serves <= 16 local rows
        bb = a if same else b
        if N > 16:
            sim = L.gathered_gram(a.detach(), bb.detach(), G, 1)
        else:
            sim = L.gathered_gram(a, bb, G, 1)
        ad, bd = a.detach().double(), b.detach().double()
        ad.requires_grad_(); bd.requires_grad_()
        ref = torch.einsum("igk,jgk->gij", ad, ad if same else bd)
        H.assert_close(sim, ref, 1e-5, f"sim same={same}")
        if N <= 16:
            ga, gb = torch.autograd.grad((sim * r).sum(), (a, b), allow_unused=True)
            ra, rb = torch.autograd.grad((ref * r.double()).sum(), (ad, bd), allow_unused=True)
            H.assert_close(ga, ra, 1e-5, f"d a same={same}")
            if not same:
                H.assert_close(gb, rb, 1e-5, f"d b same={same}")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORLD, BL, L1, L2 = 2, 3, 41, 12


def _maps():
    shp = (WORLD * BL, 8, L1, L2)
    return {k: torch.softmax(synth.normal(shp, 91, k) * 2.0, dim=-1) for k in ("a1_10", "a1_20", "a2_10", "a2_20")}


def _worker(rank, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        x = {k: v[rank * BL:(rank + 1) * BL].to(dev).requires_grad_() for k, v in _maps().items()}
        pb = L.PathBatchLoss(BL, WORLD)(x["a1_10"], x["a1_20"])
        od = L.OmicDomainScaleLoss(BL, WORLD)(x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"])
        (pb.sum() + od).backward()
        torch.cuda.synchronize()
        q.put((rank, pb.detach().cpu().numpy(), float(od), {k: v.grad.detach().cpu().numpy() for k, v in x.items()}))
    finally:
        dist.destroy_process_group()


def test_two_ranks_local_row_gradients_match_the_global_batch():
    """Each rank holds BL rows; the loss value is the global-batch value on every rank and a rank's input gradient is the
    local slice of the global gradient of its own loss copy (GatherLayer semantics, utils/gather.py:16-20)."""
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
    x = {k: v.double().requires_grad_() for k, v in _maps().items()}
    pb = OL.path_batch_loss(x["a1_10"], x["a1_20"])
    od = OL.omic_domain_scale_loss(x["a1_10"], x["a1_20"], x["a2_10"], x["a2_20"])
    (pb.sum() + od).backward()
    for rank in range(WORLD):
        pbr, odr, gr = res[rank]
        H.assert_close(torch.from_numpy(pbr), pb, TOL, f"rank {rank} PathBatchLoss")
        assert abs(odr - float(od)) <= TOL * abs(float(od))
        for k in x:
            H.assert_close(torch.from_numpy(gr[k]), x[k].grad[rank * BL:(rank + 1) * BL], 5 * TOL, f"rank {rank} d {k}")
