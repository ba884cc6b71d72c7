"""CUDA-graph capture of one training step of the hot path (forward + loss + backward into a flat gradient buffer).

A 16k-patch bag runs ~40 library / custom kernels per tower in each direction, most of them microseconds long: launched
eagerly the step is bound by the host (Python + launch latency), not by the GPU.  Capturing the step once and replaying it
removes that: every launch of the step, including the ctypes calls into libdml_b200.so (plain stream-ordered launches and
memsets, no host synchronisation, caller-owned buffers), lands in one cudaGraphLaunch.

Gradients are accumulated straight into views of one flat fp32 buffer (`parallel.FlatGradAllReducer.attach`), so the
N-GPU exchange step is a single all-reduce on that buffer between the graph replay and the optimizer step.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from .parallel import FlatGradAllReducer


class GraphedTrainStep:
    """step(inputs) -> loss (a static device tensor): copies `inputs` (dict of tensors, device or pinned host) into
    static device buffers, replays the captured forward + loss + backward, all-reduces the flat gradient buffer when a
    process group is initialised, then runs `optimizer.step()`.

    net: module called as net(**inputs_without_label_keys); loss_fn(outputs, inputs) -> scalar."""

    def __init__(self, net: torch.nn.Module, loss_fn: Callable, example: Dict[str, torch.Tensor],
                 optimizer: Optional[torch.optim.Optimizer] = None, model_keys=None, warmup: int = 3):
        self.net, self.loss_fn, self.optimizer = net, loss_fn, optimizer
        dev = next(net.parameters()).device
        self.static = {k: v.to(dev).clone() for k, v in example.items()}
        self.model_keys = list(model_keys) if model_keys is not None else list(example.keys())
        params = [p for p in net.parameters() if p.requires_grad]
        self.reducer = FlatGradAllReducer(params, static_presence=True)

        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            # eager warm-up: one-time library initialisation (kernel attributes, cuBLAS handles/workspaces, driver entry
            # points) and discovery of which parameters receive gradients
            for _ in range(max(1, warmup)):
                for p in params:
                    p.grad = None
                self._fwd_bwd()
            self.reducer.attach()              # p.grad := views of the flat buffer (None for never-used parameters)
            self._zero_and_fwd_bwd()           # once more on the attached layout
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._zero_and_fwd_bwd()

    def _fwd_bwd(self):
        out = self.net(**{k: self.static[k] for k in self.model_keys})
        loss = self.loss_fn(out, self.static)
        loss.backward()
        return loss.detach()

    def _zero_and_fwd_bwd(self):
        self.reducer.zero_grad()
        return self._fwd_bwd()

    def __call__(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        self.reducer.allreduce()
        if self.optimizer is not None:
            self.optimizer.step()
        return self.loss
