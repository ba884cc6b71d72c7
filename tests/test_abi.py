"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/dml_b200.h declares; the ctypes table covers exactly those symbols; state_dict keys of the
mirror modules equal the reference's.  No compute call is made."""
import ctypes
import os
import re

import pytest
import torch

from dml_b200 import _lib
from dml_b200.model import Args, define_net
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dml_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dml_b200.h but not exported"
    assert lib.dml_version().startswith(b"dml_b200")
    assert lib.dml_cpb_table_bytes() > 30000 and lib.dml_cpb_seg_max() >= 1089


def test_ctypes_table_matches_header():
    assert set(_lib.SIGNATURES) == header_symbols()


def test_test_only_library_is_separate_from_the_product_library():
    """The legacy mma.sync cross-check kernels and the debug knobs live in libdml_b200_test.so (include/dml_b200_test.h):
    the product library neither declares nor exports them (no process-global mutable state behind the product ABI)."""
    src = open(os.path.join(ROOT, "include", "dml_b200_test.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    tsyms = set(re.findall(r"\b(dml_[a-z0-9_]+)\s*\(", src))
    assert tsyms == set(_lib.TEST_SIGNATURES)
    assert not (tsyms & header_symbols())
    prod, test = _lib.load(), _lib.load_test()
    for s in tsyms:
        assert hasattr(test, s), s
        assert not hasattr(prod, s), f"{s} must not be exported by the product library"


def test_header_arity_matches_ctypes():
    src = open(os.path.join(ROOT, "include", "dml_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (res, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else len(params.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, ctypes table {len(args)}"


def test_integer_geometry_entry_point():
    lib = _lib.load()
    assert lib.dml_offsets_kv_len(16385, 6, 4) == 4096
    assert lib.dml_offsets_kv_len(2049, 6, 4) == 512
    assert lib.dml_offsets_kv_len(6, 6, 4) == 1


def test_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dml_b200.DeformableAttention1D import DeformCrossAttention1D
    m = DeformCrossAttention1D(dim=128, downsample_factor=4, offset_scale=2, offset_kernel_size=6)
    with pytest.raises(Exception):
        m(torch.randn(1, 128, 65), torch.randn(1, 128, 65))


def test_state_dict_contract():
    sd = define_net(Args()).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == H.pathomic_shapes()
    assert len(sd) == 118 and sum(v.numel() for v in sd.values()) == 1161288
    sd = define_net(Args(mode="path", label_dim=4)).state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == H.transmil_shapes(label_dim=4)
    assert len(sd) == 27 and sum(v.numel() for v in sd.values()) == 2738836


def test_dkv_worklist_partitions_every_item():
    """Host logic of the attention backward (no device needed): the work list handed to the dK/dV kernel covers every
    item's query-tile range exactly once, has no slivers, and is ordered longest first."""
    import ctypes as C
    from dml_b200 import _lib
    lib = _lib.load()
    cap = 400
    buf = (C.c_int * (3 * cap))()
    # north-star shape: 32 key blocks x 4 head pairs on 148 SMs
    B, H, n, n_kv, nsm = 1, 8, 16385, 4096, 148
    npieces = lib.dml_debug_dkv_worklist(B, H, n, n_kv, nsm, buf, cap)
    items, ntiles = 32 * 4, (n + 31) // 32
    assert items < npieces <= 2 * nsm
    pieces = [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(npieces)]
    cover = {}
    for it, t0, t1 in pieces:
        assert 0 <= it < items and 0 <= t0 < t1 <= ntiles
        assert t1 - t0 >= 24 or (t0 == 0 and t1 == ntiles)
        cover.setdefault(it, []).append((t0, t1))
    assert sorted(cover) == list(range(items))
    for it, rs in cover.items():
        rs.sort()
        assert rs[0][0] == 0 and rs[-1][1] == ntiles and all(a[1] == b[0] for a, b in zip(rs, rs[1:])), (it, rs)
    # pieces of a similar size near the front, the small remainders at the back (cost-ordered)
    lens = [t1 - t0 for _, t0, t1 in pieces]
    assert max(lens[: nsm // 2]) <= ntiles and min(lens[: nsm // 2]) > max(lens[-8:])
    # many items / short sequences: one CTA per item
    assert lib.dml_debug_dkv_worklist(1, 8, 100000, 25000, nsm, buf, cap) == 0
    assert lib.dml_debug_dkv_worklist(1, 8, 2000, 512, nsm, buf, cap) == 0
    # batch 2, few key blocks: still an exact partition
    npieces = lib.dml_debug_dkv_worklist(2, 8, 6400, 300, nsm, buf, cap)
    assert npieces > 0
    tot = sum(buf[3 * i + 2] - buf[3 * i + 1] for i in range(npieces))
    assert tot == (3 * 4 * 2) * ((6400 + 31) // 32)
