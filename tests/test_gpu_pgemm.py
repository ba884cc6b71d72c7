"""dml_pgemm (csrc/pgemm.cu): the bf16-pair tcgen05 GEMM against fp64 matmuls - every operand form (NT / NN / TN / TT),
batch addressing, front-padding offsets, each epilogue stage, the fused row softmax and its backward, split-K."""
import pytest
import torch

from dml_b200 import synth
from dml_b200.pairs import Pair, pgemm
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 3e-5        # 16-bit operand pairs, fp32 accumulation


def rnd(shape, name, seed=5):
    return synth.normal(shape, seed, name).to(DEV)


def ref_mm(a, b, a_trans, b_trans):
    a, b = a.double(), b.double()
    a = a.transpose(-1, -2) if a_trans else a
    b = b if b_trans else b.transpose(-1, -2)
    return a @ b


@pytest.mark.parametrize("a_trans", [False, True])
@pytest.mark.parametrize("b_trans", [False, True])
@pytest.mark.parametrize("M,N,K", [(304, 200, 152), (128, 64, 64), (520, 136, 1000)])
def test_pgemm_operand_forms(a_trans, b_trans, M, N, K):
    a = rnd((K, M) if a_trans else (M, K), "a")
    b = rnd((K, N) if b_trans else (N, K), "b")
    out, pr = pgemm(Pair.from_f32(a), Pair.from_f32(b), M=M, N=N, K=K, a_trans=a_trans, b_trans=b_trans, want_pair=True)
    ref = ref_mm(a, b, a_trans, b_trans)
    H.assert_close(out, ref, TOL, "C")
    H.assert_close(pr.float(), ref, TOL, "pair(C)")


def test_pair_roundtrip_and_exact_plane():
    x = rnd((77, 264), "x") * 1e-3
    p = Pair.from_f32(x, mult=3.0)
    H.assert_close(p.float(), 3.0 * x.double(), 1e-5, "pair")
    xb = rnd((200, 320), "xb").to(torch.bfloat16)
    w = rnd((96, 320), "w")
    out, _ = pgemm(Pair.exact(xb), Pair.from_f32(w), M=200, N=96, K=320)
    H.assert_close(out, xb.double() @ w.double().t(), TOL, "exact-plane A")


def test_pgemm_batches_head_slices_and_front_padding():
    """q . k_l^T per (bag, head) straight from a fused [B, n, 3 H d] buffer whose first `pad` logical rows are zero padding."""
    B, Hh, d, n, pad, m = 2, 8, 16, 200, 56, 64
    n_pad = n + pad
    qkv = rnd((B, n, 3 * Hh * d), "qkv")
    kl = rnd((B, Hh, m, d), "kl")
    P = Pair.from_f32(qkv)
    q = Pair(P.planes.view(2, B, n, 3, Hh, d)[:, :, :, 0].permute(0, 1, 3, 2, 4))          # [2, B, H, n, d] view
    out, _ = pgemm(q, Pair.from_f32(kl), M=n_pad, N=m, K=d, batch=(B, Hh), a_row_offset=-pad, alpha=0.25)
    qf = torch.nn.functional.pad(qkv[..., : Hh * d].reshape(B, n, Hh, d).transpose(1, 2), (0, 0, pad, 0)).double()
    ref = 0.25 * qf @ kl.double().transpose(-1, -2)
    H.assert_close(out, ref, TOL, "sim1")
    assert float(out[:, :, :pad].abs().max()) == 0.0
    # shared B along the outer batch, k offset on a transposed operand: attn3 @ v with v front-padded
    a3 = rnd((B, Hh, m, n_pad), "a3")
    v = Pair(P.planes.view(2, B, n, 3, Hh, d)[:, :, :, 2].permute(0, 1, 3, 2, 4))
    out2, _ = pgemm(Pair.from_f32(a3), v, M=m, N=d, K=n_pad, batch=(B, Hh), b_trans=True, b_k_offset=-pad)
    vf = torch.nn.functional.pad(qkv[..., 2 * Hh * d:].reshape(B, n, Hh, d).transpose(1, 2), (0, 0, pad, 0)).double()
    H.assert_close(out2, a3.double() @ vf, TOL, "attn3 @ v")


def test_pgemm_epilogue_stages():
    M, N, K = 260, 136, 200
    a, b = rnd((M, K), "a"), rnd((N, K), "b")
    A, Bp = Pair.from_f32(a), Pair.from_f32(b)
    base = a.double() @ b.double().t()
    bias, resid = rnd((N,), "bias"), rnd((M, N), "resid")
    out, _ = pgemm(A, Bp, M=M, N=N, K=K, alpha=0.5, bias=bias, relu=True)
    H.assert_close(out, torch.relu(0.5 * base + bias.double()), TOL, "bias + relu")
    out, _ = pgemm(A, Bp, M=M, N=N, K=K, alpha=0.5, ncol_split=40, alpha2=3.0, resid=resid)
    ref = 0.5 * base
    ref[:, :40] *= 3.0
    H.assert_close(out, ref + resid.double(), TOL, "column-range factor + residual")
    acc = rnd((M, N), "acc")
    ref = acc.double() + base
    pgemm(A, Bp, M=M, N=N, K=K, out=acc, accumulate=True)
    H.assert_close(acc, ref, TOL, "accumulate")
    sq = rnd((N, K), "sq")
    out, _ = pgemm(Pair.from_f32(sq), Bp, M=N, N=N, K=K, alpha=0.1, diag=7.0)
    H.assert_close(out, 7.0 * torch.eye(N, device=DEV, dtype=torch.float64) - 0.1 * sq.double() @ b.double().t(), TOL, "diag")
    half = torch.empty(M, N, device=DEV, dtype=torch.float16)
    amax = torch.zeros(1, device=DEV, dtype=torch.int32)
    hs = torch.tensor([0.125], device=DEV)
    out, _ = pgemm(A, Bp, M=M, N=N, K=K, half_out=half, half_scale_dev=hs, absmax=amax)
    H.assert_close(half.float(), 0.125 * base, 1e-3, "fp16 output")
    assert abs(float(amax.view(torch.float32)) - float(base.abs().max())) <= 1e-4 * float(base.abs().max())
    sc = torch.tensor([2.5], device=DEV)
    out, _ = pgemm(A, Bp, M=M, N=N, K=K, alpha_dev=sc)
    H.assert_close(out, 2.5 * base, TOL, "device-side factor")


@pytest.mark.parametrize("N", [256, 104, 64])
def test_pgemm_row_softmax_and_its_backward(N):
    M, K = 300, 64
    a, b = rnd((3, M, K), "a"), rnd((3, N, K), "b")
    sim = 0.3 * a.double() @ b.double().transpose(-1, -2)
    _, attn = pgemm(Pair.from_f32(a), Pair.from_f32(b), M=M, N=N, K=K, batch=(3,), alpha=0.3, softmax=1, want_f32=False,
                    want_pair=True)
    ref = sim.softmax(-1)
    H.assert_close(attn.float(), ref, TOL, "softmax(sim)")
    # backward: dA = g . w^T (a product), dS = A * (dA - rowsum(dA * A))
    g, w = rnd((3, M, 72), "g"), rnd((3, N, 72), "w")
    dA = g.double() @ w.double().transpose(-1, -2)
    dS, _ = pgemm(Pair.from_f32(g), Pair.from_f32(w), M=M, N=N, K=72, batch=(3,), softmax=2, aux=attn)
    A_ = attn.float().double()
    H.assert_close(dS, A_ * (dA - (dA * A_).sum(-1, keepdim=True)), 1e-4, "softmax backward")


def test_pgemm_split_k():
    M, N, K = 512, 192, 16640
    a, b = rnd((K, M), "a"), rnd((K, N), "b")
    out = torch.zeros(M, N, device=DEV)
    pgemm(Pair.from_f32(a), Pair.from_f32(b), M=M, N=N, K=K, a_trans=True, b_trans=True, out=out, splits=9, alpha=0.5)
    H.assert_close(out, 0.5 * a.double().t() @ b.double(), TOL, "split-K weight gradient")


def test_colsum():
    from dml_b200._lib import call, ptr, stream
    x = rnd((16385, 136), "x")
    out = torch.empty(136, device=DEV)
    call("dml_colsum", ptr(x), x.shape[0], 136, 136, ptr(out), stream())
    H.assert_close(out, x.double().sum(0), 1e-5, "colsum")
