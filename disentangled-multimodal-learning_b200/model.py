"""Mirror of the callers in the reference ``models/model.py`` that sit directly on the hot path:
MaxNet (:173-218), DeformPathomicNet (:471-568) and the part of ``define_net`` (:51-104) that selects
them.  The reference file itself keeps working unmodified against these operators;
this copy exists so that bench.py / tests can build the network on the GPU box, where /root/reference
is absent.  Same names, forward signatures and state_dict keys."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter

from .DeformCrossTransMIL import DeformCrossTransMIL
from .mil import TransMIL


def init_max_weights(module):
    """utils/utils.py:214-219 - normal(0, 1/sqrt(fan_in)) weights, zero biases for every nn.Linear."""
    for m in module.modules():
        if type(m) == nn.Linear:
            stdv = 1. / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, stdv)
            m.bias.data.zero_()


class MaxNet(nn.Module):
    def __init__(self, input_dim=59, omic_dim=32, return_grad='False', dropout_rate=0.25, label_dim=1, init_max=True):
        super().__init__()
        hidden = [64, 48, 32, 32]
        self.return_grad = return_grad
        dims = [input_dim, hidden[0], hidden[1], hidden[2], omic_dim]
        self.encoder = nn.Sequential(*[
            nn.Sequential(nn.Linear(dims[i], dims[i + 1]), nn.ELU(), nn.AlphaDropout(p=dropout_rate, inplace=False))
            for i in range(4)])
        self.relu = nn.ReLU(inplace=False)
        self.classifier = nn.Sequential(nn.Linear(omic_dim, label_dim))
        if init_max:
            init_max_weights(self)
        self.output_range = Parameter(torch.FloatTensor([6]), requires_grad=False)
        self.output_shift = Parameter(torch.FloatTensor([-3]), requires_grad=False)

    def forward(self, **kwargs):
        x = kwargs['x_omic']
        if x.is_cuda and x.dim() == 2 and len(self.encoder) == 4:
            # the four Linear -> ELU -> AlphaDropout blocks and the ReLU as one kernel per direction (csrc/maxnet.cu): as
            # separate launches they are ~40 + ~60 microsecond-sized kernels in front of / behind each tower
            from . import ops
            lin = [blk[0] for blk in self.encoder]
            p = float(self.encoder[0][2].p) if self.training else 0.0
            features = ops.MaxNetFn.apply(x, p, *[t for l_ in lin for t in (l_.weight, l_.bias)])
        else:
            features = self.relu(self.encoder(x))
        logits = self.classifier(features)
        return features, logits, None


_TOWER_STREAMS = {}


def _tower_stream(dev, which=0):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), which)
    if key not in _TOWER_STREAMS:
        from . import ops
        _TOWER_STREAMS[key] = torch.cuda.Stream(device=key[0], priority=ops.CHAIN_PRIORITY)      # see ops.branch_stream
    return _TOWER_STREAMS[key]


def _omic_ahead(net, x_omic, which):
    """The small omic MLP (some 40 microsecond-sized kernels, and 60 more in the backward) on its own stream, next to the
    tower's fc1 GEMM which does not depend on it.  The result carries the event the tower waits for right before its
    first use (DeformCrossTransMIL.forward); autograd replays the backward on the same streams."""
    cur = torch.cuda.current_stream()
    so = _tower_stream(x_omic.device, which)
    so.wait_stream(cur)
    with torch.cuda.stream(so):
        vec = net(x_omic=x_omic)[0]
        vec._dml_ready = so.record_event()
    return vec


class DeformPathomicNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        init_max = True if args.init_type == "max" else False
        self.args = args
        self.omic_net_tumor = MaxNet(input_dim=args.input_size_omic_tumor, omic_dim=args.omic_dim,
                                     return_grad=args.return_grad, dropout_rate=args.dropout_rate,
                                     label_dim=args.label_dim, init_max=init_max)
        self.omic_net_immune = MaxNet(input_dim=args.input_size_omic_immune, omic_dim=args.omic_dim,
                                      return_grad=args.return_grad, dropout_rate=args.dropout_rate,
                                      label_dim=args.label_dim, init_max=init_max)
        self.pathomic_net_tumor = DeformCrossTransMIL(args)
        self.pathomic_net_immune = DeformCrossTransMIL(args)
        if args.fusion_type != "concat":
            raise NotImplementedError("fusion_type != 'concat' (BilinearFusion) is outside the hot path "
                                      "(SURVEY.md #11); every shipped YAML uses 'concat'")
        self.classifier = nn.Linear(args.mmhid * 2, args.label_dim)
        self.classifier_tumor = nn.Sequential(nn.Linear(args.mmhid, args.label_dim))
        self.classifier_immune = nn.Sequential(nn.Linear(args.mmhid, args.label_dim))
        self.return_grad = args.return_grad
        self.fusion_type = args.fusion_type
        self.output_range = Parameter(torch.FloatTensor([6]), requires_grad=False)
        self.output_shift = Parameter(torch.FloatTensor([-3]), requires_grad=False)

    def forward(self, **kwargs):
        if getattr(self.args, "return_vgrid", False):
            raise NotImplementedError("return_vgrid is broken in the reference for attn_dim == 1 (SURVEY.md Q6)")
        # The two towers are independent until the concat: the immune tower runs on a second stream, so that its many
        # microsecond-sized kernels (projections, norms, the small omic MLP) fill the gaps of the tumor tower and vice
        # versa; autograd replays each backward node on the stream of its forward, so the backward overlaps the same way.
        x_path = kwargs['x_path']
        two_streams = x_path.is_cuda and getattr(self.args, "overlap_towers", True)
        from . import ops
        with ops.concurrent_attention_launches(2 if two_streams else 1):      # the two towers' attention kernels share the SMs
            return self._forward_towers(kwargs, x_path, two_streams)

    def _forward_towers(self, kwargs, x_path, two_streams):
        if two_streams:
            cur = torch.cuda.current_stream()
            side = _tower_stream(x_path.device)
            # issued in the reference's order - tumor, then immune (model.py:517-521) - so that the AlphaDropout draws of the
            # two omic MLPs consume the generator in the same order
            omic_vec_tumor = _omic_ahead(self.omic_net_tumor, kwargs['x_omic_tumor'], 3)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                omic_vec_immune = _omic_ahead(self.omic_net_immune, kwargs['x_omic_immune'], 2)
                vec_immune, _, grads_immune = self.pathomic_net_immune(path=x_path, omic=omic_vec_immune)
        else:
            omic_vec_tumor, _, _ = self.omic_net_tumor(x_omic=kwargs['x_omic_tumor'])
        vec_tumor, _, grads_tumor = self.pathomic_net_tumor(path=x_path, omic=omic_vec_tumor)
        if two_streams:
            cur.wait_stream(side)
            vec_immune.record_stream(cur)
        else:
            omic_vec_immune, _, _ = self.omic_net_immune(x_omic=kwargs['x_omic_immune'])
            vec_immune, _, grads_immune = self.pathomic_net_immune(path=x_path, omic=omic_vec_immune)
        features = torch.cat((vec_tumor, vec_immune), 1)
        if features.is_cuda:
            # classifier(features), classifier_tumor(vec_tumor), classifier_immune(vec_immune) (+ sigmoid for survival,
            # model.py:555-558) as one kernel per direction: nothing else can run between the forward and the backward of a bag
            from . import ops
            hazard, hazard_tumor, hazard_immune = ops.Linear3Fn.apply(
                vec_tumor, vec_immune, self.classifier.weight, self.classifier.bias, self.classifier_tumor[0].weight,
                self.classifier_tumor[0].bias, self.classifier_immune[0].weight, self.classifier_immune[0].bias,
                self.args.task_type == "survival")
        else:
            hazard = self.classifier(features)
            hazard_tumor = self.classifier_tumor(vec_tumor)
            hazard_immune = self.classifier_immune(vec_immune)
            if self.args.task_type == "survival":
                hazard = torch.sigmoid(hazard)
                hazard_tumor = torch.sigmoid(hazard_tumor)
                hazard_immune = torch.sigmoid(hazard_immune)
        logits = [hazard_tumor, hazard_immune, hazard]
        return features, vec_tumor, vec_immune, logits, None, grads_tumor, grads_immune


DIAG2021_CE_WEIGHTS = (1.0, 4.15, 2.93, 2.43)   # train_test.py:790
GRADE_CE_WEIGHTS = (1.47, 1.51, 1.0)            # train_test.py:791


def nll_loss(hazards, S, Y, c, alpha=0.4, eps=1e-7):
    """utils/utils.py:245-261."""
    batch_size = len(Y)
    Y = Y.view(batch_size, 1)
    c = c.view(batch_size, 1).float()
    if S is None:
        S = torch.cumprod(1 - hazards, dim=1)
    S_padded = torch.cat([torch.ones_like(c), S], 1)
    uncensored = -(1 - c) * (torch.log(torch.gather(S_padded, 1, Y).clamp(min=eps))
                             + torch.log(torch.gather(hazards, 1, Y).clamp(min=eps)))
    censored = -c * torch.log(torch.gather(S_padded, 1, Y + 1).clamp(min=eps))
    neg_l = censored + uncensored
    return ((1 - alpha) * neg_l + alpha * uncensored).mean()


_CE_WEIGHT_CACHE = {}


def _ce_weight(values, device):
    """Class weights as a device tensor, created once per device (a host->device copy of a fresh tensor is not
    allowed while a CUDA graph is being captured)."""
    key = (values, str(device))
    if key not in _CE_WEIGHT_CACHE:
        _CE_WEIGHT_CACHE[key] = torch.tensor(values, device=device)
    return _CE_WEIGHT_CACHE[key]


def _cumprod_bins(x):
    """torch.cumprod(x, dim=1) for the handful of survival bins as running products: the same values, but its backward is
    plain multiplications - torch's cumprod backward tests the input for zeros on the host, which a CUDA-graph capture of
    the step cannot contain."""
    cols, out = x.unbind(1), []
    run = cols[0]
    out.append(run)
    for c in cols[1:]:
        run = run * c
        out.append(run)
    return torch.stack(out, dim=1)


def bag_loss(logits, label, task_type, censor=None):
    """The loss trainDeformPathomicModel back-propagates (train_test.py:826-853): fused head only."""
    hz = logits[2]
    if task_type == "diag2021":
        return F.cross_entropy(hz, label, weight=_ce_weight(DIAG2021_CE_WEIGHTS, hz.device))
    if task_type == "grade":
        return F.cross_entropy(hz, label, weight=_ce_weight(GRADE_CE_WEIGHTS, hz.device))
    if task_type == "survival":
        return nll_loss(hz, _cumprod_bins(1 - hz), label, censor, alpha=0)
    raise ValueError(task_type)


class Args:
    """Namespace with the YAML keys the hot-path models read (config/config_mine_*.yaml)."""

    def __init__(self, **kw):
        d = dict(path_dim=128, omic_dim=128, mmhid=128, attn_dim=1, return_vgrid=False, label_dim=4,
                 input_size_omic_tumor=59, input_size_omic_immune=361, return_grad="False", dropout_rate=0.1,
                 init_type="max", fusion_type="concat", task_type="diag2021", mode="deformpathomic")
        d.update(kw)
        self.__dict__.update(d)


def define_net(args):
    """The two branches of the reference define_net (model.py:51-104) that reach the hot path."""
    if args.mode == "deformpathomic":
        return DeformPathomicNet(args=args)
    if args.mode in ("path", "transmil"):       # model.py:58-59: TransMIL is the commented-out alternative
        return TransMIL(args)
    raise NotImplementedError(f"mode {args.mode!r} is outside the hot path (SURVEY.md section 2)")
