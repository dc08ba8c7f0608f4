"""ctypes binding of libsegnn_b200.so (the C ABI in include/segnn_b200.h).

There is deliberately NO fallback: if the shared library is missing the import fails loudly, and every
entry point raises on a non-zero return code with the library's own message.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegnn_b200.so")

MODE_FP32 = 0
MODE_BF16_TC = 1
MODE_FP16_TC = 2
MODE_FP16_PACKED = 3
OPERAND_BF16, OPERAND_FP16 = 0, 1

_c = ctypes
_ptr = _c.c_void_p
_int = _c.c_int

# name -> (restype, argtypes); must list every symbol include/segnn_b200.h declares
PROTOTYPES = {
    "segnn_version": (_int, []),
    "segnn_last_error": (_c.c_char_p, []),
    "segnn_edge_index": (_int, [_int, _int, _ptr, _ptr]),
    "segnn_edge_attr": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_prep_fwd": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_edge_attr_lmax": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_prep_fwd_lmax": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_embed_fwd": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr]),
    "segnn_node_gemm": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _int, _ptr]),
    "segnn_node_gemm_tc": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _int, _int, _ptr]),
    "segnn_node_gemm_tc_out16": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _int, _ptr, _int, _ptr]),
    "segnn_tp_combine_y16": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_tp_combine_y16_x16": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_embed_fwd_x16": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_node_gemm_tc_x16": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _int, _int, _int,
                                      _ptr]),
    "segnn_edge_layer_fwd_out16": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                          _ptr]),
    "segnn_node_gemm_tc_pair16": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _int, _int, _ptr]),
    "segnn_pack_node_weight_tc": (_int, [_ptr, _int, _int, _int, _ptr, _ptr]),
    "segnn_tp_combine": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_edge_layer_fwd": (_int, [_int, _ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr,
                                    _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_pack_w2_tc": (_c.c_int64, [_ptr, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr]),
    "segnn_head_fwd": (_int, [_ptr, _ptr, _ptr, _int, _int, _ptr, _ptr]),
    "segnn_integrate": (_int, [_ptr, _ptr, _ptr, _int, _ptr, _ptr, _ptr, _int, _ptr]),
    "segnn_counter_add": (_int, [_ptr, _int, _ptr]),
    "segnn_colsum_workspace": (_c.c_int64, [_c.c_int64, _int]),
    "segnn_colsum": (_int, [_ptr, _ptr, _c.c_int64, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_colsum2": (_int, [_ptr, _ptr, _c.c_int64, _int, _int, _ptr, _ptr, _ptr, _c.c_int64, _int, _int, _ptr, _ptr,
                             _ptr]),
    "segnn_lincomb": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _c.c_int64, _int, _ptr, _ptr]),
    "segnn_bn_coeffs_fwd": (_int, [_ptr, _ptr, _int, _int, _c.c_double, _c.c_double, _ptr, _ptr, _ptr, _ptr, _c.c_double,
                                   _c.c_double, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_bn_coeffs_bwd": (_int, [_ptr, _ptr, _int, _c.c_double, _c.c_double, _ptr, _ptr, _int, _ptr, _ptr, _ptr, _ptr]),
    "segnn_add3": (_int, [_ptr, _ptr, _ptr, _c.c_int64, _ptr, _ptr]),
    "segnn_tp_combine_bwd": (_int, [_ptr, _ptr, _int, _int, _int, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_node_gemm_wgrad_workspace": (_c.c_int64, [_int, _int, _int]),
    "segnn_node_gemm_wgrad": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _int, _int, _int, _ptr, _ptr, _ptr, _ptr]),
    "segnn_edge_layer_bwd": (_int, [_int, _ptr, _ptr, _int, _int, _int] + [_ptr] * 25),
    "segnn_edge_layer_bwd_workspace": (_c.c_int64, [_int, _int, _int]),
    "segnn_node_gemm_tf32x3_workspace": (_c.c_int64, [_int, _int]),
    "segnn_node_gemm_tf32x3": (_int, [_ptr, _ptr, _int, _int, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _int, _ptr, _ptr]),
    "segnn_gemm_tn_tf32x3_workspace": (_c.c_int64, [_c.c_int64, _int, _int]),
    "segnn_gemm_tn_tf32x3": (_int, [_ptr, _c.c_int64, _ptr, _c.c_int64, _c.c_int64, _int, _int, _ptr, _c.c_int64, _int,
                                    _ptr, _ptr]),
    "segnn_gemm_tn_grouped_tf32x3": (_int, [_ptr, _c.c_int64, _ptr, _c.c_int64, _c.c_int64, _int, _int, _int, _ptr,
                                            _c.c_int64, _int, _ptr, _ptr]),
    "segnn_l2_planarize": (_int, [_ptr, _c.c_int64, _int, _c.c_int64, _ptr, _ptr, _ptr, _ptr]),
    "segnn_l2_msg_rows": (_int, [_ptr, _ptr, _int, _int, _int, _c.c_int64, _ptr, _c.c_int64, _ptr, _c.c_int64, _ptr,
                                 _c.c_int64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _c.c_int64, _c.c_int64, _c.c_int64, _ptr,
                                 _ptr, _ptr, _ptr]),
    "segnn_l2_gate_aggregate": (_int, [_int, _int, _int, _c.c_int64, _ptr, _c.c_int64, _ptr, _c.c_int64, _ptr, _c.c_int64,
                                       _ptr, _ptr, _ptr, _ptr, _ptr]),
    "segnn_edge_layer_gemm_workspace": (_c.c_int64, [_int, _int, _int, _int, _c.c_int64]),
    "segnn_edge_layer_gemm_fwd": (_int, [_ptr, _ptr, _int, _int, _int] + [_ptr] * 13 + [_c.c_int64, _ptr]),
    "segnn_edge_layer_gemm_bwd": (_int, [_ptr, _ptr, _int, _int, _int] + [_ptr] * 25 + [_c.c_int64, _ptr, _c.c_int64,
                                                                                         _ptr]),
    "segnn_edge_layer_gemm_bwd_phases": (_int, [_ptr, _ptr, _int, _int, _int] + [_ptr] * 25 +
                                         [_c.c_int64, _ptr, _c.c_int64, _int, _ptr]),
    "segnn_embed_bwd": (_int, [_ptr, _ptr, _ptr, _int, _int, _ptr, _ptr]),
    "segnn_head_bwd": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_generic_tp": (_int, [_ptr, _int, _ptr, _int, _c.c_int64, _ptr, _ptr, _int, _ptr, _ptr, _int, _ptr, _ptr]),
    "segnn_generic_tp_expand": (_int, [_ptr, _int, _ptr, _int, _c.c_int64, _ptr, _int, _ptr, _int, _int, _ptr, _ptr]),
    "segnn_generic_tp_scatter": (_int, [_ptr, _c.c_int64, _int, _int, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_generic_tp_expand_ld": (_int, [_ptr, _int, _ptr, _int, _c.c_int64, _ptr, _int, _ptr, _int, _int, _c.c_int64, _ptr,
                                          _ptr]),
    "segnn_generic_tp_scatter_ld": (_int, [_ptr, _c.c_int64, _c.c_int64, _int, _int, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_gemm_tf32x3_workspace": (_c.c_int64, [_int, _int]),
    "segnn_gemm_tf32x3": (_int, [_ptr, _c.c_int64, _ptr, _c.c_int64, _c.c_int64, _int, _int, _ptr, _c.c_int64, _ptr, _ptr]),
    "segnn_generic_hoisted_msg1": (_int, [_ptr, _int, _ptr, _int, _ptr, _int, _int, _int, _ptr, _int, _ptr, _int, _ptr, _int,
                                          _int, _ptr, _ptr, _ptr, _int, _ptr, _ptr]),
    "segnn_generic_gate": (_int, [_ptr, _c.c_int64, _int, _int, _int, _ptr, _ptr, _ptr]),
    "segnn_generic_message_input": (_int, [_ptr, _ptr, _int, _int, _int, _int, _ptr, _ptr]),
    "segnn_generic_aggregate": (_int, [_ptr, _int, _int, _int, _ptr, _ptr]),
    "segnn_edge_attr_list": (_int, [_ptr, _ptr, _ptr, _c.c_int64, _int, _ptr, _ptr, _ptr]),
    "segnn_generic_message_input_list": (_int, [_ptr, _ptr, _ptr, _c.c_int64, _int, _int, _ptr, _ptr]),
    "segnn_segment_reduce": (_int, [_ptr, _ptr, _ptr, _c.c_int64, _int, _int, _ptr, _ptr]),
    "segnn_prep_fwd_list": (_int, [_ptr, _ptr, _ptr, _c.c_int64, _int, _ptr, _ptr, _ptr]),
    "segnn_sim_gravity": (_int, [_ptr, _ptr, _ptr, _int, _int, _c.c_double, _c.c_double, _c.c_double, _int, _int, _ptr,
                                 _ptr, _ptr, _ptr]),
    "segnn_macros_counters": (_int, [_ptr, _ptr, _int, _int, _int, _int, _c.c_float, _c.c_float, _c.c_float, _ptr, _ptr,
                                     _ptr]),
    "segnn_macros_energy_momentum": (_int, [_ptr, _ptr, _int, _int, _int, _c.c_float, _c.c_float, _ptr, _ptr]),
    "segnn_sim_charged": (_int, [_ptr, _ptr, _ptr, _int, _int, _c.c_double, _c.c_double, _c.c_double, _int, _int, _ptr,
                                 _ptr, _ptr]),
    "segnn_macros_group_collisions_workspace": (_c.c_int64, [_int, _int, _int]),
    "segnn_macros_group_collisions": (_int, [_ptr, _int, _int, _int, _int, _c.c_float, _ptr, _ptr, _ptr]),
    "segnn_knn_edge_index": (_int, [_ptr, _int, _int, _int, _int, _ptr, _ptr]),
    "segnn_pack_weights_size": (_c.c_int64, [_int, _int]),
    "segnn_pack_weights": (_int, [_int, _int, _ptr, _ptr, _ptr, _ptr]),
    "segnn_fold_batchnorm": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _c.c_float, _c.c_float, _ptr, _ptr, _ptr]),
    "segnn_instance_norm": (_int, [_ptr, _ptr, _int, _int, _ptr, _int, _ptr, _ptr, _c.c_float, _ptr, _ptr]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the sm_100a kernels first (python -c 'import __graft_entry__ as g; "
            f"g.build()' or csrc/build.sh). There is no CPU or PyTorch fallback for the SEGNN hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


class _NvtxLib:
    """SEGNN_NVTX=1: every C-ABI entry point runs inside an NVTX range named after it (one range per kernel family:
    segnn_edge_layer_fwd, segnn_node_gemm_tc, ...), so that nsys / ncu --nvtx timelines attribute device time to the
    reference function each export replaces.  Off by default: a range costs ~1 us of host time per call."""

    def __init__(self, inner):
        import torch
        self._inner, self._nvtx = inner, torch.cuda.nvtx

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not name.startswith("segnn_") or name in ("segnn_last_error", "segnn_version"):
            return fn
        nvtx = self._nvtx

        def ranged(*args):
            nvtx.range_push(name)
            try:
                return fn(*args)
            finally:
                nvtx.range_pop()
        ranged.__name__ = name
        setattr(self, name, ranged)
        return ranged


lib = _load()
if os.environ.get("SEGNN_NVTX", "0") not in ("", "0"):
    lib = _NvtxLib(lib)


class SegnnKernelError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.segnn_last_error()
        raise SegnnKernelError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
