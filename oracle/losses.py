"""Oracle restatement of the batch / domain losses (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/utils/loss.py: PathBatchLoss :25-64, OmicDomainScaleLoss :90-143 (+ diag_variance_loss :82-85),
BatchLoss :220-253, on tensors that already hold the rows of ALL ranks (the reference concatenates the GatherLayer output,
:36-38; the multi-rank semantics - only the local rows receive a gradient, utils/gather.py:16-20 - are applied by the tests
through ``local_rows``)."""
from __future__ import annotations

import torch


def _sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    s = a.mm(b.t())
    return s / torch.norm(s, 2, 1).view(-1, 1)


def path_batch_loss(att10: torch.Tensor, att20: torch.Tensor) -> torch.Tensor:
    """att [N, 8, L1, L2] -> [N, N] (unreduced, as the reference returns it)."""
    N = att10.shape[0]
    a10 = att10.reshape(N, 8, -1).transpose(0, 1)
    a20 = att20.reshape(N, 8, -1).transpose(0, 1)
    m10 = torch.stack([_sim(x, x) for x in a10]).mean(0)
    m20 = torch.stack([_sim(x, x) for x in a20]).mean(0)
    return (m10 - m20) ** 2 / N


def omic_domain_scale_loss(a1_10, a1_20, a2_10, a2_20) -> torch.Tensor:
    N = a1_10.shape[0]
    s1 = _sim(a1_10.reshape(N, -1), a1_20.reshape(N, -1))
    s2 = _sim(a2_10.reshape(N, -1), a2_20.reshape(N, -1))
    return 10000 * torch.var(s1.diagonal()) + 10000 * torch.var(s2.diagonal())


def batch_loss(omic: torch.Tensor, vgrid: torch.Tensor) -> torch.Tensor:
    N = omic.shape[0]
    o = omic.reshape(N, -1)
    v = vgrid.reshape(8, N, -1)
    s = _sim(o, o)
    mv = torch.stack([_sim(x, x) for x in v]).mean(0)
    return (s - mv) ** 2 / N
