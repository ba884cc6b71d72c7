"""bf16 operand pairs and the tcgen05 pair GEMM (csrc/pgemm.cu, `dml_pgemm`).

A ``Pair`` holds an fp32-class matrix x [..., rows, cols] as two bf16 planes (x ~= hi + lo, 16 significant bits, the fp32
exponent range: no scales).  ``pgemm`` multiplies two of them in any of the NT / NN / TN forms straight from their row-major
storage and fuses the epilogue stages the path needs (include/dml_b200.h, `dml_pgemm_args`).
"""
from __future__ import annotations

import ctypes as C
import threading
from contextlib import contextmanager
from typing import Optional

import torch

from . import _lib
from ._lib import PgemmArgs, PgOperand, call, ptr, stream

BF16 = torch.bfloat16
F32 = torch.float32
F16 = torch.float16


_tls = threading.local()


class _Chain:
    """Problems recorded inside `chain()`: launched together by dml_pgemm_chain (one cooperative kernel per group of
    dml_pgemm_chain_max() dependent problems, a grid barrier between problems), or one by one where the chained form does not
    apply (tile width other than 128, more tiles than SMs)."""

    def __init__(self):
        self.args, self.keep = [], []

    def flush(self):
        if not self.args:
            return
        lib = _lib.load(check_device=True)
        cap = lib.dml_pgemm_chain_max()
        st = stream()
        i = 0
        while i < len(self.args):
            grp = self.args[i:i + cap]
            arr = (PgemmArgs * len(grp))(*grp)
            if _lib._timing_hook is not None:
                _lib._timing_hook("dml_pgemm_chain", 0)
            rc = lib.dml_pgemm_chain(C.addressof(arr), len(grp), st) if len(grp) > 1 else -2
            if _lib._timing_hook is not None:
                _lib._timing_hook("dml_pgemm_chain", 1)
            if rc == 0:
                _lib.launch_count += 1
            elif rc == -2:                       # DML_E_UNSUPPORTED: the same problems as separate launches, in order
                for a in grp:
                    call("dml_pgemm", C.addressof(a), st)
            else:
                raise _lib.DmlError(f"dml_pgemm_chain failed: {rc}")
            i += cap
        self.args, self.keep = [], []


import os as _os

CHAIN_DEFAULT = _os.environ.get("DML_B200_PGEMM_CHAIN", "0") != "0"


@contextmanager
def chain(enabled: Optional[bool] = None):
    """Record the pgemm() calls of the block and launch them as chained cooperative kernels at its end.  Only pgemm() calls
    may touch the recorded operands inside the block (their results do not exist until it closes)."""
    if enabled is None:
        enabled = CHAIN_DEFAULT
    if not enabled or getattr(_tls, "chain", None) is not None:
        yield None
        return
    c = _Chain()
    _tls.chain = c
    try:
        yield c
    finally:
        _tls.chain = None
    c.flush()


class Pair:
    """planes: bf16 [2, *shape] (plane 0 = hi, plane 1 = lo) or [1, *shape] for data that is exact in bf16."""

    __slots__ = ("planes",)

    def __init__(self, planes: torch.Tensor):
        assert planes.dtype == BF16 and planes.shape[0] in (1, 2)
        self.planes = planes

    @property
    def shape(self):
        return self.planes.shape[1:]

    @property
    def nplanes(self):
        return self.planes.shape[0]

    @staticmethod
    def empty(shape, device, planes: int = 2) -> "Pair":
        return Pair(torch.empty((planes,) + tuple(shape), device=device, dtype=BF16))

    @staticmethod
    def from_f32(x: torch.Tensor, mult: float = 1.0) -> "Pair":
        """One pass over a contiguous fp32 tensor (rows x last dim)."""
        x = x.contiguous().float()
        cols = x.shape[-1]
        assert cols % 8 == 0, "pair tensors keep the last dimension a multiple of 8"
        out = Pair.empty(x.shape, x.device)
        call("dml_pair_from_f32", ptr(x), x.numel() // cols, cols, cols, float(mult), ptr(out.planes), cols, x.numel(), stream())
        return out

    @staticmethod
    def exact(x_bf16: torch.Tensor) -> "Pair":
        """A bf16 tensor as a one-plane operand (no copy)."""
        assert x_bf16.dtype == BF16
        return Pair(x_bf16.unsqueeze(0))

    def float(self) -> torch.Tensor:
        return self.planes.float().sum(0)

    def b1(self) -> "Pair":
        """[P, R, C] -> [P, 1, R, C]: a 2-D operand shared by every entry of a one-dimensional batch."""
        return Pair(self.planes.unsqueeze(1))

    def head_slices(self, which: int, n_groups: int, H: int, d: int) -> "Pair":
        """planes [P, B, n, n_groups * H * d] -> view [P, B, H, n, d] of column group `which` (q / k / v of a fused buffer)."""
        P_, B, n, _ = self.planes.shape
        return Pair(self.planes.view(P_, B, n, n_groups, H, d)[:, :, :, which].permute(0, 1, 3, 2, 4))

    def view(self, *shape) -> "Pair":
        return Pair(self.planes.view(self.nplanes, *shape))

    def cols(self, c0: int, c1: int) -> "Pair":
        """Column slice of the last dimension (a view)."""
        return Pair(self.planes[..., c0:c1])


def _operand(t: Pair, trans: bool, K: int, rows: int, batch_dims: int, row_offset=0, k_offset=0) -> PgOperand:
    """t.planes: [P, (bo,) (bi,) R, Cc] view (last dim contiguous).  trans=False: memory rows = operand rows, columns = K
    (layout 0).  trans=True: memory rows = K, columns = operand rows (layout 1)."""
    pl = t.planes
    assert pl.stride(-1) == 1, "pair operands are row-major views"
    o = PgOperand()
    o.base = pl.data_ptr()
    o.plane_stride = pl.stride(0) if pl.shape[0] == 2 else 0
    nd = pl.dim() - 1
    assert nd == 2 + batch_dims, (pl.shape, batch_dims)
    o.bs_outer = o.bs_inner = 0
    if batch_dims == 2:
        o.bs_outer = pl.stride(1) if pl.shape[1] > 1 else 0
        o.bs_inner = pl.stride(2) if pl.shape[2] > 1 else 0
    elif batch_dims == 1:
        o.bs_inner = pl.stride(1) if pl.shape[1] > 1 else 0
    o.ld = pl.stride(-2)
    R, Cc = pl.shape[-2], pl.shape[-1]
    if not trans:
        o.layout, o.rows, o.k_mem = 0, R, Cc
    else:
        o.layout, o.rows, o.k_mem = 1, Cc, R
    o.row_offset, o.k_offset = int(row_offset), int(k_offset)
    return o


def pgemm(A: Pair, B: Pair, *, M: int, N: int, K: int, a_trans=False, b_trans=False, batch=(), a_row_offset=0, a_k_offset=0,
          b_row_offset=0, b_k_offset=0, alpha=1.0, alpha_dev=None, ncol_split=0, alpha2=1.0, bias=None, relu=False,
          diag=None, resid=None, resid_scale=1.0, accumulate=False, out: Optional[torch.Tensor] = None, want_f32=True, want_pair=False,
          pair_out: Optional[Pair] = None, half_out: Optional[torch.Tensor] = None, half_scale_dev=None, absmax=None,
          softmax=0, aux: Optional[Pair] = None, splits=1):
    """value[b][m, n] = alpha * sum_k A_b(m, k) B_b(n, k) with the epilogue of `dml_pgemm_args`.

    A: planes [P, *batch, RA, CA]; a_trans=False reads A_b(m, k) = mem[m + a_row_offset, k + a_k_offset], a_trans=True
    A_b(m, k) = mem[k + a_k_offset, m + a_row_offset]; the same for B with n in place of m.  `batch` = () | (bi,) | (bo, bi);
    an operand whose batch extent is 1 is shared.  bias: float [*batch?, N]; resid / out: float [*batch, M, N] views with
    unit column stride.  Returns (out_f32 or None, pair_out or None)."""
    dev = A.planes.device
    nb = len(batch)
    assert nb <= 2
    bo, bi = (batch + (1, 1))[:2] if nb == 2 else ((1, batch[0]) if nb == 1 else (1, 1))
    a = PgemmArgs()
    a.A = _operand(A, a_trans, K, M, nb, a_row_offset, a_k_offset)
    a.B = _operand(B, b_trans, K, N, nb, b_row_offset, b_k_offset)
    a.M, a.N, a.K, a.nb_inner, a.nb_outer, a.splits = M, N, K, bi, bo, max(1, int(splits))
    a.alpha, a.alpha2, a.ncol_split = float(alpha), float(alpha2), int(ncol_split)
    a.alpha_dev = ptr(alpha_dev) if alpha_dev is not None else None
    keep = [A, B]

    def bstrides(t, lead):
        """(inner, outer) element strides of the `lead` leading batch dims of t."""
        if lead == 0:
            return 0, 0
        if lead == 1:
            return (t.stride(0) if t.shape[0] > 1 else 0), 0
        return (t.stride(1) if t.shape[1] > 1 else 0), (t.stride(0) if t.shape[0] > 1 else 0)

    if bias is not None:
        assert bias.dtype == F32 and bias.stride(-1) == 1 and bias.shape[-1] == N
        lead = bias.dim() - 1
        a.bias = ptr(bias)
        a.bias_bs_inner, a.bias_bs_outer = bstrides(bias, lead) if lead == nb else (0, 0)
        assert lead in (0, nb)
        keep.append(bias)
    a.relu = int(bool(relu))
    if diag is not None:
        a.use_diag, a.diag = 1, float(diag)
    if resid is not None:
        assert resid.dtype == F32 and resid.stride(-1) == 1 and resid.dim() == nb + 2
        a.resid, a.ldr, a.resid_scale = ptr(resid), resid.stride(-2), float(resid_scale)
        a.r_bs_inner, a.r_bs_outer = bstrides(resid, nb)
        keep.append(resid)
    a.accumulate = int(bool(accumulate))
    shape = tuple(batch) + (M, N)
    reduce_batch = out is not None and out.dim() == 2 and nb > 0      # every batch entry reduces into ONE [M, N] output
    if reduce_batch:
        splits = max(2, int(splits))                                   # the reducing (atomic) epilogue
        a.splits = splits
    if splits > 1:
        assert out is not None, "split-K reduces into a caller-zeroed output"
    if out is None and (want_f32 or accumulate):
        out = torch.empty(shape, device=dev, dtype=F32)
    if out is not None:
        assert out.dtype == F32 and out.stride(-1) == 1 and (out.dim() == nb + 2 or reduce_batch)
        a.c, a.ldc = ptr(out), out.stride(-2)
        a.c_bs_inner, a.c_bs_outer = (0, 0) if reduce_batch else bstrides(out, nb)
    if pair_out is None and want_pair:
        assert N % 8 == 0
        pair_out = Pair.empty(shape, dev)
    if pair_out is not None:
        pp = pair_out.planes
        assert pp.shape[0] == 2 and pp.stride(-1) == 1 and pp.dim() == nb + 3
        a.pair, a.ldp, a.p_plane = pp.data_ptr(), pp.stride(-2), pp.stride(0)
        a.p_bs_inner, a.p_bs_outer = bstrides(pp[0], nb)
    if half_out is not None:
        assert half_out.dtype == F16 and half_out.stride(-1) == 1 and half_out.dim() == nb + 2
        a.half_out, a.ldh = ptr(half_out), half_out.stride(-2)
        a.h_bs_inner, a.h_bs_outer = bstrides(half_out, nb)
        a.half_scale_dev = ptr(half_scale_dev) if half_scale_dev is not None else None
    if absmax is not None:
        a.absmax = ptr(absmax)
    a.softmax = int(softmax)
    if aux is not None:
        xp = aux.planes
        assert xp.shape[0] == 2 and xp.stride(-1) == 1 and xp.dim() == nb + 3
        a.aux, a.ldx, a.x_plane = xp.data_ptr(), xp.stride(-2), xp.stride(0)
        a.x_bs_inner, a.x_bs_outer = bstrides(xp[0], nb)
        keep.append(aux)
    rec = getattr(_tls, "chain", None)
    if rec is not None:
        rec.args.append(a)
        rec.keep.append((A, B, bias, resid, out, pair_out, half_out, aux, alpha_dev, half_scale_dev, absmax))
    else:
        call("dml_pgemm", C.addressof(a), stream())
    return out, pair_out
