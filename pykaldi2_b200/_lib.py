"""ctypes binding of libpk2.so (C ABI: include/pk2.h).

The product path has no CPU fallback: if the CUDA library is missing or a call
fails, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpk2.so")
_lib = None

vp = C.c_void_p


class SupBatch(C.Structure):
    _fields_ = [("n_seq", C.c_int)] + [(n, vp) for n in (
        "seq_state_off", "lvl_base", "level_off", "num_frames",
        "out_off", "out_dst", "out_pdf", "out_w",
        "in_off", "in_src", "in_pdf", "in_w",
        "final_cost", "state_time")]


class LatBatch(C.Structure):
    _fields_ = [("n_seq", C.c_int)] + [(n, vp) for n in (
        "seq_state_off", "lvl_base", "level_off", "num_frames",
        "out_off", "out_dst", "out_tid", "out_gc",
        "in_off", "in_src", "in_tid", "in_gc",
        "eps_off", "eps_src", "eps_dst", "eps_gc",
        "final_cost", "state_time", "tid2pdf", "num_ali", "frame_base", "keep")]


class LstmFwdArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("H", C.c_int)] + [(n, vp) for n in (
        "gx", "whh", "y", "gates", "cstate", "sync")]


class LstmBwdArgs(C.Structure):
    _fields_ = [("B", C.c_int), ("T", C.c_int), ("H", C.c_int)] + [(n, vp) for n in (
        "dy", "whh_t", "gates", "cstate", "dgates", "sync", "whh_t_perm")]


_SIGS = {
    "pk2_version": (C.c_int, []),
    "pk2_last_error": (C.c_char_p, []),
    "pk2_launch_count": (C.c_int64, []),
    "pk2_fbank_plan_create": (C.c_int, [vp, C.POINTER(vp)]),
    "pk2_fbank_plan_destroy": (C.c_int, [vp]),
    "pk2_fbank": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "pk2_colmean": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp]),
    "pk2_gather_norm": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "pk2_ce_softmax": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_float, vp, vp, vp]),
    "pk2_den_graph_create": (C.c_int, [C.c_int, C.c_int, vp, vp, vp, vp, vp, C.POINTER(vp)]),
    "pk2_den_graph_destroy": (C.c_int, [vp]),
    "pk2_denfb_workspace_bytes": (C.c_size_t, [vp, C.c_int, C.c_int]),
    "pk2_denfb": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_float, C.c_float,
                            vp, vp, vp, C.c_int, vp]),
    "pk2_numfb": (C.c_int, [C.POINTER(SupBatch), vp, C.c_int, C.c_int64, C.c_float, vp, vp, vp, vp, vp]),
    "pk2_numfb_post": (C.c_int, [C.POINTER(SupBatch), vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]),
    "pk2_numfb_scatter": (C.c_int, [C.POINTER(SupBatch), C.c_int, vp, C.c_int, C.c_int64, C.c_float, vp, vp]),
    "pk2_chain_guard": (C.c_int, [vp, vp, C.c_int, C.c_int64, vp, vp]),
    "pk2_latfb_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int]),
    "pk2_latfb_mmi": (C.c_int, [C.POINTER(LatBatch), vp, C.c_int, C.c_int, C.c_int64, C.c_float,
                                C.c_float, vp, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp]),
    "pk2_latfb_mpe": (C.c_int, [C.POINTER(LatBatch), vp, vp, vp, C.c_int, C.c_int, C.c_int64, C.c_float, C.c_float,
                                vp, C.c_int64, C.c_int64, C.c_float, vp, vp, vp, vp]),
    "pk2_gemm_bf16_nt": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, vp]),
    "pk2_gemm_bf16_ex": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, vp, vp]),
    "pk2_gemm_set_max_ctas": (C.c_int, [C.c_int]),
    "pk2_cast_bf16": (C.c_int, [vp, vp, C.c_int64, vp]),
    "pk2_transpose_bf16": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "pk2_lstm_hprev_t": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "pk2_colsum_bf16": (C.c_int, [vp, vp, C.c_int64, C.c_int, vp]),
    "pk2_lstm_hprev": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "pk2_lstm_pack_layer": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp]),
    "pk2_gather_rows_bf16": (C.c_int, [vp, C.c_int, vp, vp, C.c_int64, C.c_int, vp]),
    "pk2_zero_pad_rows": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int64, vp]),
    "pk2_lstm_input_proj": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "pk2_lstm_layer_fwd": (C.c_int, [C.POINTER(LstmFwdArgs), vp]),
    "pk2_lstm_set_profile_buffer": (C.c_int, [vp]),
    "pk2_den_set_profile_buffer": (C.c_int, [vp]),
    "pk2_den_plan": (C.c_longlong, [vp, C.c_int, C.c_int, C.c_int, vp]),
    "pk2_den_set_sm_budget": (C.c_int, [vp, C.c_int, C.c_int]),
    "pk2_lstm_layer_bwd": (C.c_int, [C.POINTER(LstmBwdArgs), vp]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load libpk2.so (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "pykaldi2_b200: %s not found -- build it with `python -m pykaldi2_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if os.environ.get("PK2_DEV_PARTIAL") and not hasattr(L, name):
                continue      # development only: library built without some translation units
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed: %s" % (what, lib().pk2_last_error().decode()))


def ptr(t):
    """Device (or host) pointer of a tensor as c_void_p; None -> NULL."""
    if t is None:
        return vp(0)
    if isinstance(t, vp):
        return t
    assert t.is_contiguous(), "pk2: tensor must be contiguous"
    return vp(t.data_ptr())


def ptr_at(t, elems):
    """Pointer to element ``elems`` of a tensor's storage view (operand sub-matrices with a leading dimension)."""
    return vp(t.data_ptr() + int(elems) * t.element_size())


def stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("pykaldi2_b200: %s must be a CUDA tensor (no CPU fallback)" % name)
