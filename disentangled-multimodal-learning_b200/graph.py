"""CUDA-graph capture of one training step of the hot path (forward + loss + backward into a flat gradient buffer).

A 16k-patch bag runs ~40 library / custom kernels per tower in each direction, most of them microseconds long: launched
eagerly the step is bound by the host (Python + launch latency), not by the GPU.  Capturing the step once and replaying it
removes that: every launch of the step, including the ctypes calls into libdml_b200.so (plain stream-ordered launches and
memsets, no host synchronisation, caller-owned buffers), lands in one cudaGraphLaunch.

Gradients end up in one flat fp32 buffer (`parallel.FlatGradAllReducer`), so the N-GPU exchange step is a single all-reduce
on that buffer between the graph replay and the optimizer step.  Inside the captured step the parameters' .grad is None when
the backward starts - autograd then keeps each incoming gradient tensor instead of launching an `add` into an existing one -
and one multi-tensor copy moves them into the flat buffer at the end (`_zero_and_fwd_bwd`; measured 3.643 -> 3.599 ms per
DeformPathomicNet step against accumulating into views of the flat buffer, which DML_B200_STEAL_GRADS=0 restores).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import os

import torch

from .parallel import FlatGradAllReducer


def _chain_priority():
    from .ops import CHAIN_PRIORITY
    return CHAIN_PRIORITY


# DML_B200_STEAL_GRADS=0: accumulate into views of the flat buffer instead (one `add` kernel per parameter and step)
STEAL_GRADS = os.environ.get("DML_B200_STEAL_GRADS", "1") != "0"


class GraphedTrainStep:
    """step(inputs) -> loss (a static device tensor): copies `inputs` (dict of tensors, device or pinned host) into
    static device buffers, replays the captured forward + loss + backward, all-reduces the flat gradient buffer when a
    process group is initialised, then runs `optimizer.step()`.

    net: module called as net(**inputs_without_label_keys); loss_fn(outputs, inputs) -> scalar.

    flat_optimizer: instead of `optimizer`, a callable `params -> optimizer`.  The parameters are then moved into one
    flat buffer laid out like the flat gradient buffer (each parameter becomes a view of it, values unchanged) and the
    optimizer is built over the few contiguous runs of parameters that receive gradients - an element-wise optimizer
    with uniform hyper-parameters (AdamW here) computes exactly the same update in one small launch instead of one
    multi-tensor launch chain over ~120 tensors (85 us per step for DeformPathomicNet).  Parameters that never receive a
    gradient are left out, as torch's optimizers skip parameters whose grad is None."""

    def __init__(self, net: torch.nn.Module, loss_fn: Callable, example: Dict[str, torch.Tensor],
                 optimizer: Optional[torch.optim.Optimizer] = None, model_keys=None, warmup: int = 3,
                 flat_optimizer: Optional[Callable] = None):
        self.net, self.loss_fn, self.optimizer = net, loss_fn, optimizer
        dev = next(net.parameters()).device
        self.static = {k: v.to(dev).clone() for k, v in example.items()}
        self.model_keys = list(model_keys) if model_keys is not None else list(example.keys())
        params = [p for p in net.parameters() if p.requires_grad]
        self.reducer = FlatGradAllReducer(params, static_presence=True)

        cur = torch.cuda.current_stream()
        # kernel nodes keep the priority of the capturing stream (ops.CHAIN_PRIORITY, default 0)
        side = torch.cuda.Stream(device=dev, priority=_chain_priority())
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # eager warm-up: one-time library initialisation (kernel attributes, cuBLAS handles/workspaces, driver entry
            # points) and discovery of which parameters receive gradients
            for _ in range(max(1, warmup)):
                for p in params:
                    p.grad = None
                self._fwd_bwd()
            self.reducer.attach()              # p.grad := views of the flat buffer (None for never-used parameters)
            if flat_optimizer is not None:
                self.optimizer = flat_optimizer(self._flatten_params())
            self._zero_and_fwd_bwd()           # once more on the attached layout
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._zero_and_fwd_bwd()

    def _flatten_params(self):
        """Parameters -> views of one flat buffer (same layout as reducer.flat); returns one Parameter per contiguous
        run of parameters that receive gradients, its .grad the matching slice of the flat gradient buffer."""
        r = self.reducer
        self.flat_params = torch.empty(r.total, dtype=torch.float32, device=r.flat.device)
        segs, run = [], None
        self.flat_params.zero_()
        for p, offs, n, has in zip(r.params, r.offsets, r.sizes, r._present):
            assert p.dtype == torch.float32, "flat optimizer: fp32 parameters only"
            slot = self.flat_params[offs: offs + n]
            slot.copy_(p.data.reshape(-1))
            p.data = slot.view_as(p)
            end = offs + (n + 31) // 32 * 32          # the padding rides along (zero parameter, zero gradient: stays zero)
            if has:
                run = [offs, end] if run is None else [run[0], end]
            elif run is not None:
                segs.append(run)
                run = None
        if run is not None:
            segs.append(run)
        out = []
        for a, b in segs:
            q = torch.nn.Parameter(self.flat_params[a:b])
            q.grad = r.flat[a:b]
            out.append(q)
        return out

    def _fwd_bwd(self):
        out = self.net(**{k: self.static[k] for k in self.model_keys})
        loss = self.loss_fn(out, self.static)
        loss.backward()
        return loss.detach()

    def _zero_and_fwd_bwd(self):
        if not STEAL_GRADS or getattr(self.reducer, "_present", None) is None:
            self.reducer.zero_grad()
            return self._fwd_bwd()
        # Autograd adds every incoming gradient into p.grad when that is defined: ~100 microsecond-sized `add` kernels strung
        # along the backward's chains.  With p.grad = None it keeps the incoming tensor itself (no kernel); the gradients then
        # reach the flat buffer in one multi-tensor copy.  The padding between the slots stays zero (zeroed at construction), and
        # every parameter that ever received a gradient receives one per step (static presence), so no per-step memset either.
        r = self.reducer
        for p in r.params:
            p.grad = None
        loss = self._fwd_bwd()
        views, grads = [], []
        for p, v, has in zip(r.params, r._views(r.flat), r._present):
            if has:
                if p.grad is None:            # used by another captured step (e.g. another bag length) but not by this one
                    v.zero_()
                else:
                    views.append(v.view_as(p))
                    grads.append(p.grad)
        if views:
            torch._foreach_copy_(views, grads)
        return loss

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.reducer.allreduce()
        if self.optimizer is not None:
            self.optimizer.step()
        return self.loss
