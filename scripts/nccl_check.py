"""2-rank NCCL check of the N>1 plumbing on real GPUs (the CPU suite covers the same logic with gloo):
GatherLayer forward/backward and FlatGradAllReducer against a single-process gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dml_b200.gather import GatherLayer
from dml_b200.parallel import FlatGradAllReducer
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
t = (torch.arange(6, dtype=torch.float32, device=dev).reshape(2, 3) + 10 * rank).requires_grad_()
full = torch.cat(GatherLayer.apply(t), 0)
w = torch.arange(full.numel(), dtype=torch.float32, device=dev).reshape(full.shape) * (rank + 1)
(full * w).sum().backward()
exp = torch.cat([torch.arange(6, dtype=torch.float32, device=dev).reshape(2, 3) + 10 * r for r in range(world)], 0)
assert torch.equal(full.detach(), exp) and torch.equal(t.grad, w[2 * rank: 2 * rank + 2])
torch.manual_seed(3)
net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3)).to(dev)
torch.manual_seed(100)
x, y = torch.randn(8, 6, device=dev), torch.randn(8, 3, device=dev)
((net(x[rank::world]) - y[rank::world]) ** 2).mean().backward()
FlatGradAllReducer(net.parameters()).allreduce()
g = [p.grad.clone() for p in net.parameters()]
net.zero_grad()
((net(x) - y) ** 2).mean().backward()
for a, p in zip(g, net.parameters()):
    assert torch.allclose(a, p.grad, rtol=1e-5, atol=1e-7)
dist.barrier()
if rank == 0:
    print("nccl_check ok: GatherLayer fwd/bwd, flat gradient all-reduce ==", world, "ranks")
dist.destroy_process_group()
