"""ctypes binding of libdml_b200.so (the C ABI declared in include/dml_b200.h).

There is no fallback: if the library is missing or the device is not a B200 (sm_100a) every
call raises.  The library is built in-tree by ``build.py`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdml_b200.so")

_vp, _i, _f, _ll, _fp = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_void_p

# name -> (restype, argtypes).  Every symbol of include/dml_b200.h is listed; tests/test_abi.py checks
# that the header and this table agree and that the shared object exports all of them.
SIGNATURES = {
    "dml_runtime_check": (_i, []),
    "dml_version": (C.c_char_p, []),
    "dml_cpb_table_bytes": (C.c_size_t, []),
    "dml_cpb_seg_max": (_i, []),
    "dml_cpb_table_build": (_i, [_fp] * 6 + [_i, _i, _f, _vp, _vp]),
    "dml_cpb_eval": (_i, [_vp, _fp, _i, _fp, _vp, _vp]),
    "dml_cpb_param_grad": (_i, [_fp] * 6 + [_i, _i, _vp, _fp, _fp, _vp]),
    "dml_offsets_kv_len": (_i, [_i, _i, _i]),
    "dml_offsets_fwd": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _f, _fp, _fp, _vp]),
    "dml_offsets_bwd": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _f, _i, _i, _i, _i, _i, _i, _f, _fp, _fp, _vp, _vp]),
    "dml_offsets_bwd_pair": (_i, [_vp, _fp, _fp, _fp, _fp, _fp, _f, _i, _i, _i, _i, _i, _i, _f, _fp, _fp, _vp, _vp, _ll, _vp]),
    "dml_kv_gather_fwd": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _vp]),
    "dml_kv_gather_bwd": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _f, _f, _fp, _fp, _vp]),
    "dml_deform_attn_fwd_tc": (_i, [_vp, _vp, _vp, _fp, _vp] + [_i] * 11 + [_f, _vp, _fp, _vp]),
    "dml_deform_attn_fwd_tc_split": (_i, [_vp, _vp, _vp, _fp, _vp] + [_i] * 11 + [_f, _vp, _fp, _i, _vp]),
    "dml_deform_attn_bwd_tc": (_i, [_vp, _vp, _vp, _fp, _vp, _vp, _vp, _fp] + [_i] * 11 + [_f] + [_fp] * 7 + [_vp, _vp]),
    "dml_deform_attn_bwd_ws_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "dml_deform_attn_dq_from_ds": (_i, [_vp, _vp, _fp] + [_i] * 6 + [_fp, _vp]),
    "dml_layernorm_fwd": (_i, [_fp, _fp, _fp, _ll, _i, _f, _fp, _fp, _fp, _vp]),
    "dml_layernorm_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _ll, _i, _fp, _fp, _fp, _vp]),
    "dml_pgemm": (_i, [C.c_void_p, _vp]),
    "dml_pgemm_chain_max": (_i, []),
    "dml_pgemm_chain": (_i, [C.c_void_p, _i, _vp]),
    "dml_pair_from_f32": (_i, [_fp, _ll, _i, _i, _f, _vp, _i, _ll, _vp]),
    "dml_colsum": (_i, [_fp, _ll, _i, _i, _fp, _vp]),
    "dml_scale_to_half": (_i, [_fp, _fp, _ll, _vp, _vp]),
    "dml_loss_scale_from_amax": (_i, [_vp, _fp, _vp]),
    "dml_relu_mask_pair": (_i, [_fp, _fp, _fp, _ll, _i, _i, _i, _vp, _i, _ll, _vp]),
    "dml_layernorm_fwd_pair": (_i, [_fp, _fp, _fp, _ll, _i, _f, _fp, _vp, _ll, _fp, _fp, _vp]),
    "dml_ny_landmark_pool": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _i, _f, _f, _vp, _ll, _vp]),
    "dml_ny_softmax_rows_fwd": (_i, [_fp, _ll, _i, _vp, _ll, _vp]),
    "dml_ny_softmax_rows_bwd": (_i, [_vp, _ll, _fp, _ll, _i, _vp, _ll, _vp]),
    "dml_ny_res_conv_fwd": (_i, [_fp, _vp, _ll, _i, _i, _fp, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    "dml_ny_res_conv_bwd": (_i, [_fp, _vp, _ll, _i, _i, _fp, _i, _i, _i, _i, _i, _fp, _i, _i, _fp, _vp]),
    "dml_ny_dqkv_finalize": (_i, [_fp, _fp, _i, _i, _i, _i, _i, _f, _vp, _ll, _vp]),
    "dml_ppeg_stencil": (_i, [_fp, _fp, _fp, _i, _i, _i, _i, _fp, _vp]),
    "dml_ppeg_wgrad": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp, _vp]),
    "dml_maxnet_fwd": (_i, [_fp, _vp, _vp, _vp, _i, _fp, _f, _fp, _fp, _fp, _vp]),
    "dml_maxnet_bwd": (_i, [_fp, _vp, _vp, _vp, _i, _fp, _f, _fp, _fp, _fp, _fp, _fp, _vp]),
    "dml_tower_head_fwd": (_i, [_fp, _ll, _i, _i, _fp, _fp, _f, _fp, _fp, _i, _fp, _fp, _i, _fp, _fp, _fp, _fp, _vp]),
    "dml_tower_head_bwd": (_i, [_fp, _ll, _i, _i, _fp, _fp, _i, _fp, _i, _fp, _fp, _fp, _fp, _fp, _fp, _ll, _vp]),
    "dml_linear3_fwd": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _fp, _fp, _fp, _vp]),
    "dml_linear3_bwd": (_i, [_fp, _fp, _i, _i, _i, _fp, _fp, _fp, _i, _i, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp, _vp]),
    "dml_debug_dkv_worklist": (_i, [_i, _i, _i, _i, _i, C.POINTER(C.c_int), _i]),
    "dml_ny_pinv_init_sums_floats": (C.c_size_t, [_i, _i]),
    "dml_ny_pinv_init_part_floats": (C.c_size_t, [_i, _i]),
    "dml_ny_pinv_init_fwd": (_i, [_fp, _i, _i, _fp, _vp, _ll, _vp]),
    "dml_ny_pinv_init_bwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _fp, _fp, _vp]),
    "dml_gram_splits": (_i, [_i, _ll, _i]),
    "dml_gram_fwd": (_i, [_vp, _ll, _vp, _ll, _i, _i, _ll, _fp, _vp]),
    "dml_rows_mix": (_i, [_fp, _vp, _ll, _i, _i, _i, _ll, _fp, _ll, _ll, _vp]),
    "dml_coattn_chunks": (_i, [_i, _i, _i]),
    "dml_coattn_fq_fwd_ws_floats": (C.c_size_t, [_i, _i, _i, _i]),
    "dml_coattn_fq_bwd_ws_floats": (C.c_size_t, [_i, _i, _i, _i]),
    "dml_coattn_fk_bwd_ws_floats": (C.c_size_t, [_i, _i, _i, _i]),
    "dml_coattn_fq_fwd": (_i, [_fp, _ll, _ll, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _vp]),
    "dml_coattn_fq_bwd": (_i, [_fp, _ll, _ll, _fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _vp]),
    "dml_coattn_fk_fwd": (_i, [_fp, _ll, _ll, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _vp]),
    "dml_coattn_fk_bwd": (_i, [_fp, _ll, _ll, _fp, _ll, _ll, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _vp]),
    "dml_da2_kv_side": (_i, [_i, _i, _i]),
    "dml_da2_gproj_fwd": (_i, [_fp, _fp, _ll, _fp, _vp]),
    "dml_da2_gproj_parts": (_i, [_ll]),
    "dml_da2_gproj_bwd": (_i, [_fp, _fp, _fp, _ll, _i, _fp, _fp, _fp, _vp]),
    "dml_da2_reduce_parts": (_i, [_fp, _i, _ll, _i, _fp, _vp]),
    "dml_da2_offsets_fwd": (_i, [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _fp, _fp, _vp]),
    "dml_da2_offsets_parts": (_i, [_i, _i, _i, _i]),
    "dml_da2_offsets_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _i, _i, _i, _i, _f, _fp, _fp, _fp, _fp, _vp]),
    "dml_da2_gather_fwd": (_i, [_fp, _fp, _i, _i, _i, _fp, _vp]),
    "dml_da2_gather_bwd": (_i, [_fp, _fp, _fp, _i, _i, _i, _fp, _fp, _vp]),
    "dml_da2_bias_fwd": (_i, [_fp] * 7 + [_i, _i, _i, _fp, _vp]),
    "dml_da2_bias_bwd_parts": (_i, [_i, _i]),
    "dml_da2_bias_bwd": (_i, [_fp] * 7 + [_i, _i, _i, _fp, _fp, _fp, _vp]),
    "dml_da2_attn_ws_bytes": (C.c_size_t, [_i, _i, _i, _i]),
    "dml_da2_attn_fwd": (_i, [_fp, _fp, _fp, _fp, _vp, _f, _i, _i, _i, _f, _vp, _fp, _vp]),
    "dml_da2_cols_chunks": (_i, [_i, _i, _i]),
    "dml_da2_attn_bwd": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _vp, _f, _i, _i, _i, _f, _vp, _fp, _fp, _fp, _fp, _vp]),
    "dml_dpc_split": (_i, [_fp, _fp, _ll, _i, _vp, _fp, _fp, _vp]),
    "dml_dpc_density": (_i, [_vp, _fp, _fp, _fp, _i, _i, _i, _fp, _fp, _vp]),
    "dml_dpc_parent": (_i, [_vp, _fp, _fp, _fp, _fp, _i, _i, _i, _fp, _vp]),
    "dml_dpc_assign": (_i, [_fp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "dml_merge_ws_floats": (_ll, [_i, _i, _i]),
    "dml_merge_fwd": (_i, [_fp, _fp, _vp, _i, _i, _i, _i, _fp, _fp, _fp, _vp]),
    "dml_merge_bwd": (_i, [_fp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, _i, _fp, _fp, _vp]),
}



class PgOperand(C.Structure):
    """dml_pg_operand (include/dml_b200.h)."""
    _fields_ = [("base", C.c_void_p), ("plane_stride", C.c_longlong), ("bs_inner", C.c_longlong), ("bs_outer", C.c_longlong),
                ("layout", C.c_int), ("ld", C.c_int), ("rows", C.c_int), ("row_offset", C.c_int), ("k_offset", C.c_int),
                ("k_mem", C.c_int)]


class PgemmArgs(C.Structure):
    """dml_pgemm_args (include/dml_b200.h)."""
    _fields_ = [("A", PgOperand), ("B", PgOperand),
                ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("nb_inner", C.c_int), ("nb_outer", C.c_int), ("splits", C.c_int),
                ("alpha", C.c_float), ("alpha2", C.c_float), ("ncol_split", C.c_int),
                ("alpha_dev", C.c_void_p), ("bias", C.c_void_p), ("bias_bs_inner", C.c_longlong), ("bias_bs_outer", C.c_longlong),
                ("relu", C.c_int), ("use_diag", C.c_int), ("diag", C.c_float),
                ("resid", C.c_void_p), ("ldr", C.c_int), ("r_bs_inner", C.c_longlong), ("r_bs_outer", C.c_longlong),
                ("resid_scale", C.c_float), ("accumulate", C.c_int),
                ("c", C.c_void_p), ("ldc", C.c_int), ("c_bs_inner", C.c_longlong), ("c_bs_outer", C.c_longlong),
                ("pair", C.c_void_p), ("ldp", C.c_int), ("p_bs_inner", C.c_longlong), ("p_bs_outer", C.c_longlong),
                ("p_plane", C.c_longlong),
                ("half_out", C.c_void_p), ("ldh", C.c_int), ("h_bs_inner", C.c_longlong), ("h_bs_outer", C.c_longlong),
                ("half_scale_dev", C.c_void_p), ("absmax", C.c_void_p), ("softmax", C.c_int),
                ("aux", C.c_void_p), ("ldx", C.c_int), ("x_bs_inner", C.c_longlong), ("x_bs_outer", C.c_longlong),
                ("x_plane", C.c_longlong)]


# libdml_b200_test.so (include/dml_b200_test.h): TEST-ONLY, loaded by tests / scripts through load_test(), never by the product
# path.  It also exports its own dml_deform_attn_bwd_tc / dml_deform_attn_dq_from_ds (the knob build of the same source).
TEST_LIB_PATH = os.path.join(_HERE, "lib", "libdml_b200_test.so")
TEST_SIGNATURES = {
    "dml_deform_attn_fwd": (_i, [_vp, _vp, _vp, _fp, _vp] + [_i] * 10 + [_f, _vp, _fp, _vp]),
    "dml_deform_attn_bwd": (_i, [_vp, _vp, _vp, _fp, _vp, _vp, _vp, _fp] + [_i] * 10 + [_f] + [_fp] * 7 + [_vp]),
    "dml_debug_set_trace": (_i, [_vp]),
    "dml_debug_set_seg_limit": (_i, [_i]),
    "dml_test_mma_sync_peak": (_i, [_i, _i, _i, _fp, _vp]),
}

_lock = threading.Lock()
_lib = None
_checked_devices = set()


class DmlError(RuntimeError):
    pass


def load(check_device: bool = False):
    """dlopen the shared object (once) and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise DmlError(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  dml_b200 has no CPU / eager fallback.")
                lib = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = lib
    if check_device:
        dev = torch.cuda.current_device()
        if dev not in _checked_devices:
            rc = _lib.dml_runtime_check()
            if rc != 0:
                raise DmlError(f"dml_b200 needs an sm_100a (B200) device; dml_runtime_check() = {rc}")
            _checked_devices.add(dev)
    return _lib


_test_lib = None


def load_test():
    """dlopen the test-only library (tests and profiling scripts only)."""
    global _test_lib
    if _test_lib is None:
        with _lock:
            if _test_lib is None:
                if not os.path.exists(TEST_LIB_PATH):
                    raise DmlError(f"{TEST_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
                lib = C.CDLL(TEST_LIB_PATH)
                both = dict(TEST_SIGNATURES)
                both["dml_deform_attn_bwd_tc"] = SIGNATURES["dml_deform_attn_bwd_tc"]
                for name, (res, args) in both.items():
                    fn = getattr(lib, name)
                    fn.restype = res
                    fn.argtypes = args
                _test_lib = lib
    return _test_lib


def call_test(name: str, *args):
    """Invoke an int-returning entry point of the TEST-ONLY library; raise on failure."""
    load(check_device=True)
    rc = getattr(load_test(), name)(*args)
    if rc != 0:
        what = _ERR.get(rc) or (f"CUDA error {rc}" if rc > 0 else f"error {rc}")
        raise DmlError(f"{name} (test library) failed: {what}")


_ERR = {-1: "invalid argument", -2: "unsupported shape/config", -3: "workspace too small"}

# kernels launched per entry point (memsets not counted) - bench.py reports the total as gpu_launches
KERNELS_PER_CALL = {
    "dml_cpb_table_build": 1, "dml_cpb_eval": 1, "dml_cpb_param_grad": 1, "dml_offsets_fwd": 1, "dml_offsets_bwd": 2, "dml_offsets_bwd_pair": 2,
    "dml_kv_gather_fwd": 1, "dml_kv_gather_bwd": 1, "dml_deform_attn_fwd_tc": 1, "dml_deform_attn_fwd_tc_split": 1, "dml_deform_attn_bwd_tc": 3, "dml_deform_attn_dq_from_ds": 1,
    "dml_layernorm_fwd": 1, "dml_layernorm_bwd": 1,
    "dml_pgemm": 1, "dml_pgemm_chain": 1, "dml_pair_from_f32": 1, "dml_colsum": 1, "dml_layernorm_fwd_pair": 1, "dml_ny_landmark_pool": 1,
    "dml_ny_softmax_rows_fwd": 1, "dml_ny_softmax_rows_bwd": 1, "dml_ny_res_conv_fwd": 1, "dml_ny_res_conv_bwd": 1,
    "dml_ny_dqkv_finalize": 1, "dml_ppeg_stencil": 1, "dml_ppeg_wgrad": 1, "dml_relu_mask_pair": 1, "dml_scale_to_half": 1, "dml_maxnet_fwd": 1, "dml_maxnet_bwd": 1, "dml_tower_head_fwd": 1, "dml_tower_head_bwd": 1,
    "dml_linear3_fwd": 1, "dml_linear3_bwd": 1,
    "dml_ny_pinv_init_fwd": 2, "dml_ny_pinv_init_bwd": 2, "dml_gram_fwd": 1, "dml_rows_mix": 1,
    "dml_coattn_fq_fwd": 2, "dml_coattn_fq_bwd": 1, "dml_coattn_fk_fwd": 1, "dml_coattn_fk_bwd": 1,
    "dml_da2_gproj_bwd": 3, "dml_da2_offsets_bwd": 3, "dml_da2_bias_bwd": 2, "dml_da2_attn_bwd": 5, "dml_da2_attn_fwd": 3, "dml_merge_fwd": 2,
}
launch_count = 0        # kernels of libdml_b200.so launched by this process
_timing_hook = None     # bench.py installs a (name, phase) callback to bracket calls with CUDA events


def call(name: str, *args, kernels=None):
    """Invoke an int-returning entry point on the current CUDA device/stream; raise on failure.  kernels: how many kernels
    this particular call launches when that differs from KERNELS_PER_CALL (staged calls)."""
    global launch_count
    lib = load(check_device=True)
    if _timing_hook is not None:
        _timing_hook(name, 0)
    rc = getattr(lib, name)(*args)
    if _timing_hook is not None:
        _timing_hook(name, 1)
    launch_count += KERNELS_PER_CALL.get(name, 1) if kernels is None else kernels
    if rc != 0:
        what = _ERR.get(rc) or (f"CUDA error {rc}" if rc > 0 else f"error {rc}")
        raise DmlError(f"{name} failed: {what}")


def ptr(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise DmlError("dml_b200 ops need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream
