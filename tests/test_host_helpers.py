"""Host-side helpers of dml_b200.ops that are plain torch (no kernels): they run on the CPU too."""
import torch

from dml_b200 import ops


def test_colsum_is_the_column_sum():
    torch.manual_seed(0)
    x = torch.randn(1037, 24)
    assert torch.allclose(ops.colsum(x), x.sum(0), rtol=1e-5, atol=1e-5)
    assert torch.allclose(ops.colsum(x[:5]), x[:5].sum(0), rtol=1e-5, atol=1e-6)      # another row count: another cached ones vector


def test_wgrad_mm_matches_the_single_gemm():
    torch.manual_seed(1)
    for rows in (100, 2048, 5000):          # below the chunking threshold, exact multiple, remainder rows
        a, b = torch.randn(rows, 12), torch.randn(rows, 7)
        assert torch.allclose(ops.wgrad_mm(a, b), a.t() @ b, rtol=1e-4, atol=1e-3)


def test_add_row_bias_gradients():
    torch.manual_seed(2)
    x = torch.randn(2, 50, 6, dtype=torch.float64, requires_grad=True)
    v = torch.randn(2, 1, 6, dtype=torch.float64, requires_grad=True)
    w = torch.randn(2, 50, 6, dtype=torch.float64)
    # colsum accumulates through a float32 ones vector: compare against plain broadcasting in float32
    xf, vf = x.detach().float().requires_grad_(), v.detach().float().requires_grad_()
    (ops.AddRowBiasFn.apply(xf, vf) * w.float()).sum().backward()
    ((x + v) * w).sum().backward()
    assert torch.allclose(xf.grad.double(), x.grad, rtol=1e-6, atol=1e-6)
    assert torch.allclose(vf.grad.double(), v.grad, rtol=1e-5, atol=1e-5)


def test_grad_scale_is_a_power_of_two_in_range():
    for amax in (3e-7, 0.02, 1.0, 900.0):
        t = torch.tensor([[-amax, 0.5 * amax], [0.1 * amax, 0.0]])
        s, inv = ops.grad_scale(t).tolist()
        assert 4.0 < s * amax <= 8.0 and abs(s * inv - 1.0) < 1e-6
        assert abs(torch.log2(torch.tensor(s)).item() - round(torch.log2(torch.tensor(s)).item())) < 1e-6
    assert ops.grad_scale(torch.zeros(3, 3)).tolist() == [1.0, 1.0]
