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
