"""the synthetic dataset (6-tuple batches, label layout of train_test.py:817-820) through the trainer glue - first-step
loss against the oracle, eager and per-length CUDA-graph steps agreeing, parameters moving, rank-0 evaluation."""
import logging

import pytest
import torch

from dml_b200 import synth
from dml_b200.data.dataset import LABEL_COLUMNS, SyntheticBagDataset, TCGA_Dataset
from dml_b200.model import Args, define_net
from dml_b200.train_test import LABEL_COL, c_index, trainDeformPathomicModel
from oracle import towers
from tests import helpers as H

DEV = "cuda"


def test_dataset_contract_cpu():
    ds = SyntheticBagDataset(5, (300, 400), seed=7)
    x_path, x20, x_omic, x_t, x_i, label = ds[3]
    assert x_path.shape == (ds.lengths[3], 1024) and ds.lengths[3] % 2 == 0 and 300 <= ds.lengths[3] <= 400
    assert x_omic.shape == (431,) and x_t.shape == (59,) and x_i.shape == (361,) and label.shape == (LABEL_COLUMNS,)
    assert 0 <= int(label[5]) < 4 and 0 <= int(label[4]) < 3 and 0 <= int(label[8]) < 4 and int(label[9]) in (0, 1)
    again = ds[3]
    assert torch.equal(again[0], x_path) and torch.equal(again[5], label)            # same index, same bag
    ref = TCGA_Dataset(excel_wsi=list(range(6)), args=Args(synthetic_patches=64, seed=3))
    assert len(ref) == 6 and ref[0][0].shape == (64, 1024)
    assert abs(c_index(torch.tensor([3.0, 2.0, 1.0]), torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1, 1, 1])) - 1.0) < 1e-12


def _net(task):
    net = define_net(Args(task_type=task))
    net.load_state_dict(synth.fill_like(H.pathomic_shapes(), 77), strict=True)
    return net.to(DEV)


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["diag2021", "survival"])
def test_trainer_steps_match_oracle_and_graph(task):
    ds = SyntheticBagDataset(4, (200, 260), seed=9)
    ds.lengths = [200, 260, 200, 260]
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False)
    hist = {}
    nets = {}
    for use_graph in (False, True):
        net = _net(task).eval()                                                   # eval(): no dropout, comparable across modes
        args = Args(task_type=task, epochs=2, rank=0, world_size=1, batch_size=1, use_graph=use_graph, eval_every=1)
        opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.0)
        before = {k: v.detach().clone() for k, v in net.named_parameters()}
        net.train = lambda mode=True, _n=net: _n                                   # the trainer calls model.train(): keep eval semantics
        hist[use_graph] = trainDeformPathomicModel(net, (loader, loader), opt, None, logging.getLogger("t"), args)
        nets[use_graph] = net
        moved = sum(float((v - before[k]).abs().max()) > 0 for k, v in net.named_parameters() if v.grad is not None or use_graph)
        assert moved > 50
    for a, b in zip(hist[False], hist[True]):
        assert abs(a["loss"] - b["loss"]) <= 2e-3 * max(1.0, abs(a["loss"])), (a, b)
        assert ("accuracy" in a) or ("c_index" in a)
    assert hist[False][1]["loss"] == hist[False][1]["loss"]                         # finite
    # first step of a fresh net against the oracle restatement of the reference
    net = _net(task).eval()
    x_path, _, _, x_t, x_i, label = next(iter(loader))
    out = net(x_path=x_path.to(DEV), x_omic_tumor=x_t.to(DEV), x_omic_immune=x_i.to(DEV))
    P = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    _, _, _, ref_logits = towers.deform_pathomic_net(x_path, x_t, x_i, P, task_type=task)
    lab = label.long()
    ref_loss = towers.bag_loss(ref_logits, lab[:, LABEL_COL[task]], task, lab[:, 9] if task == "survival" else None)
    from dml_b200.train_test import _loss_of
    loss = _loss_of(out[3], lab.to(DEV), task)
    assert abs(float(loss) - float(ref_loss)) <= 5e-3 * max(1.0, abs(float(ref_loss)))
