"""Trainer glue for the hot path: the step of ``trainDeformPathomicModel`` (``train_test.py:784-1050``) on the
B200 operators - same batch tuple, same loss selection per ``args.task_type`` (``:826-853``), same optimizer / scheduler calls - with
the reference's two gradient exchanges (DDP buckets + the per-parameter all-reduce loop, ``:970-981``) replaced by ONE flat
all-reduce, and optional CUDA-graph replay of the step per bag length (variable-length bags: one captured graph per length bucket).

    torchrun --nproc-per-node N -m dml_b200.train_test --task survival --bags 32 --patches 4096 16384 --epochs 1

Rank-0 evaluation uses accuracy (classification tasks) or Harrell's C-index on the summed survival (``utils/utils.py:CIndex``
semantics: risk = -sum S).  wandb / checkpoints / SHAP of the reference are callers' concerns and are not reproduced.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import parallel
from .graph import GraphedTrainStep
from .model import Args, bag_loss, define_net

LABEL_COL = {"grade": 4, "diag2021": 5, "subtype": 7, "survival": 8}      # train_test.py:820


def _loss_of(logits, label, task_type):
    censor = label[:, 9] if task_type == "survival" else None
    return bag_loss(logits, label[:, LABEL_COL[task_type]], task_type, censor)


def c_index(risk: torch.Tensor, time_: torch.Tensor, event: torch.Tensor) -> float:
    """Harrell's concordance over comparable pairs (the earlier time is an observed event)."""
    r, t, e = risk.double().cpu(), time_.double().cpu(), event.bool().cpu()
    lt = (t[:, None] < t[None, :]) & e[:, None]
    n = int(lt.sum())
    if n == 0:
        return float("nan")
    conc = ((r[:, None] > r[None, :]) & lt).sum() + 0.5 * ((r[:, None] == r[None, :]) & lt).sum()
    return float(conc) / n


class _Stepper:
    """One training step on a batch; eager, or a CUDA graph per bag length when args.use_graph."""

    def __init__(self, model, optimizer, args):
        self.model, self.optimizer, self.args = model, optimizer, args
        self.graphs: Dict[int, GraphedTrainStep] = {}
        params = [p for p in model.parameters() if p.requires_grad]
        self.reducer = parallel.FlatGradAllReducer(params) if params else None

    def _inputs(self, batch):
        x_path, _, x_omic, x_omic_tumor, x_omic_immune, label = batch
        dev = next(self.model.parameters()).device
        return {"x_path": x_path.to(dev, non_blocking=True), "x_omic_tumor": x_omic_tumor.to(dev, non_blocking=True),
                "x_omic_immune": x_omic_immune.to(dev, non_blocking=True), "label": label.to(dev, non_blocking=True).long()}

    def __call__(self, batch) -> torch.Tensor:
        a = self.args
        inp = self._inputs(batch)
        if getattr(a, "use_graph", False):
            key = (inp["x_path"].shape[0], inp["x_path"].shape[1])
            if key not in self.graphs:
                self.graphs[key] = GraphedTrainStep(
                    self.model, lambda out, b: _loss_of(out[3], b["label"], a.task_type), inp, optimizer=self.optimizer,
                    model_keys=["x_path", "x_omic_tumor", "x_omic_immune"])
            g = self.graphs[key]
            g.reducer.attach_views()            # p.grad -> this bucket's flat gradient buffer (each captured graph owns one)
            return g(inp).clone()
        out = self.model(x_path=inp["x_path"], x_omic_tumor=inp["x_omic_tumor"], x_omic_immune=inp["x_omic_immune"])
        loss = _loss_of(out[3], inp["label"], a.task_type)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.reducer is not None:
            self.reducer.allreduce()                # no-op without a process group
        self.optimizer.step()
        return loss.detach()


@torch.no_grad()
def evaluate(model, loader, args) -> Dict[str, float]:
    model.eval()
    dev = next(model.parameters()).device
    preds, labels = [], []
    for x_path, _, x_omic, x_t, x_i, label in loader:
        out = model(x_path=x_path.to(dev), x_omic_tumor=x_t.to(dev), x_omic_immune=x_i.to(dev))
        preds.append(out[3][2].float().cpu())
        labels.append(label)
    model.train()
    p, y = torch.cat(preds), torch.cat(labels)
    if args.task_type == "survival":
        S = torch.cumprod(1 - p, dim=1)
        return {"c_index": c_index(-S.sum(1), y[:, 11], 1 - y[:, 9])}
    return {"accuracy": float((p.argmax(1) == y[:, LABEL_COL[args.task_type]].long()).float().mean())}


def trainDeformPathomicModel(model, dataloader, optimizer, scheduler, logger, args) -> List[Dict[str, float]]:
    """Same call signature as the reference trainer (``train_test.py:784``): dataloader = (train_loader, test_loader)."""
    train_loader, test_loader = dataloader
    rank = int(getattr(args, "rank", 0))
    step = _Stepper(model, optimizer, args)
    model.train()
    history = []
    for epoch in range(args.epochs):
        sampler = getattr(train_loader, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        t0, losses = time.time(), []
        for batch in train_loader:
            losses.append(step(batch))
            if scheduler is not None:
                scheduler.step()
        rec = {"epoch": epoch, "loss": float(torch.stack(losses).mean()) if losses else float("nan"),
               "bags_per_s": len(losses) * train_loader.batch_size * int(getattr(args, "world_size", 1)) / max(time.time() - t0, 1e-9)}
        if rank == 0 and test_loader is not None and (epoch + 1) % int(getattr(args, "eval_every", 1)) == 0:
            rec.update(evaluate(model, test_loader, args))
        if rank == 0 and logger is not None:
            logger.info(" ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in rec.items()))
        history.append(rec)
    return history


def main(argv: Optional[List[str]] = None):
    import logging
    from .data.dataset import SyntheticBagDataset
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--task", default="diag2021", choices=sorted(LABEL_COL))
    ap.add_argument("--bags", type=int, default=16)
    ap.add_argument("--patches", type=int, nargs="+", default=[2500], help="one length, or lo hi for variable-length bags")
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=2e-4)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--bucketed", action="store_true", help="length-bucketed sampler: the ranks of a step get bags of neighbouring lengths")
    a = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    args = Args(task_type=a.task, epochs=a.epochs, rank=rank, world_size=world, batch_size=1, use_graph=a.graph, eval_every=1)
    torch.manual_seed(42)
    model = define_net(args).to(dev)
    n_p = a.patches[0] if len(a.patches) == 1 else (a.patches[0], a.patches[1])
    train = SyntheticBagDataset(a.bags, n_p, seed=42, bag_dtype=torch.bfloat16)
    test = SyntheticBagDataset(max(8, a.bags // 4), n_p, seed=43, bag_dtype=torch.bfloat16)
    sampler = None
    if a.bucketed:
        sampler = parallel.LengthBucketedSampler(train.lengths, world, rank, seed=42)
    elif world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(train, world, rank, shuffle=True, seed=42, drop_last=True)
    tl = torch.utils.data.DataLoader(train, batch_size=1, shuffle=sampler is None, sampler=sampler, num_workers=0, pin_memory=True, drop_last=True)
    vl = torch.utils.data.DataLoader(test, batch_size=1, shuffle=False) if rank == 0 else None
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=a.lr, weight_decay=0.01)
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    hist = trainDeformPathomicModel(model, (tl, vl), opt, None, logging.getLogger("dml_b200.train"), args)
    if world > 1:
        dist.destroy_process_group()
    return hist


if __name__ == "__main__":
    main()
