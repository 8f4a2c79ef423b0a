"""ctypes binding of libqtx_b200.so (the C ABI declared in include/qtx_b200.h).

There is no CPU fallback: if the shared library is missing every compute entry point of the
package raises.  PyTorch is used only as the owner of device memory and of the CUDA stream;
tensors cross the boundary as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqtx_b200.so")

QTX_F32, QTX_F64 = 0, 1
QTX_LOCAL_FLIP, QTX_SPIN_EXCHANGE = 0, 1

_vp, _i32, _i64, _u64, _f64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/qtx_b200.h
SIGNATURES = {
    "qtx_last_error": (C.c_char_p, []),
    "qtx_abi_version": (_i32, []),
    "qtx_launch_count": (_i64, []),
    "qtx_launch_count_reset": (None, []),
    "qtx_rbm_forward": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "qtx_rbm_workspace_size": (_sz, [_i32, _i32, _i32]),
    "qtx_rbm_sweep": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _f64,
                             _vp, _vp, _vp, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qtx_rbm_oloc": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _sz, _vp]),
    "qtx_rbm_ref_forward": (_i32, [_i32, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _vp]),
    "qtx_rbm_jacobian": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "qtx_rbm_colmean_workspace_size": (_sz, [_i32, _i32, _i32, _i64]),
    "qtx_rbm_jacobian_colmean": (_i32, [_i32, _vp, _vp, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qtx_resconv_tc_available": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "qtx_resconv_tc_backward_available": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "qtx_resconv_nparams": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "qtx_resconv_workspace_size": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "qtx_resconv_forward": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _sz,
                                   _vp]),
    "qtx_resconv_jacobian": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _vp, _i64,
                                    _vp, _vp, _vp, _sz, _vp]),
    "qtx_metropolis_propose": (_i32, [_i32, _vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _u64, _u64, _u64, _vp, _vp,
                                      _vp]),
    "qtx_metropolis_accept": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _f64, _vp, _u64, _u64, _u64, _vp,
                                     _vp, _vp]),
    "qtx_compact_moved": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "qtx_resconv_forward_n": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp,
                                     _vp, _sz, _vp]),
    "qtx_metropolis_accept_compact": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _f64, _vp, _u64,
                                             _u64, _u64, _vp, _vp, _vp]),
    "qtx_symm_images": (_i32, [_vp, _i64, _i32, _vp, _i32, _i32, _vp, _vp]),
    "qtx_symm_combine": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "qtx_weighted_rowsum": (_i32, [_i32, _vp, _i64, _vp, _i64, _i32, _i64, _vp, _i64, _vp]),
    "qtx_conn_count": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "qtx_exclusive_scan_i32": (_i32, [_vp, _i64, _vp, _vp, _vp]),
    "qtx_conn_fill": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "qtx_apply_diag": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "qtx_oloc_reduce": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "qtx_colmean": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qtx_center_scale": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qtx_ebar": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "qtx_gram_workspace_size": (_sz, [_i32, _i64, _i64, _i32]),
    "qtx_gram": (_i32, [_i32, _vp, _i64, _i64, _i64, _i32, _vp, _i32, _vp, _sz, _vp]),
    "qtx_peer_alloc": (_i32, [_sz, _vp]),
    "qtx_peer_free": (_i32, [_vp]),
    "qtx_peer_export": (_i32, [_vp, _vp]),
    "qtx_peer_open": (_i32, [_vp, _vp]),
    "qtx_peer_close": (_i32, [_vp]),
    "qtx_gram_push": (_i32, [_i32, _vp, _i64, _i64, _i64, _i32, _vp, _i32, _i32, _vp, _vp, _sz, _vp]),
    "qtx_peer_signal": (_i32, [_vp, _i32, _i32, _u64, _vp]),
    "qtx_gram_reduce": (_i32, [_vp, _i32, _i64, _vp, _vp, _u64, _f64, _vp]),
    "qtx_pinv_eig_workspace_size": (_sz, [_i64]),
    "qtx_pinv_eig_solve": (_i32, [_vp, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qtx_pinv_eig_solve_snr": (_i32, [_vp, _i64, _vp, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qtx_eigh": (_i32, [_vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "qtx_rows_dot_snr": (_i32, [_vp, _i64, _i64, _i64, _vp, _f64, _vp, _vp]),
    "qtx_pinv_apply": (_i32, [_vp, _i64, _vp, _vp, _f64, _f64, _vp, _vp]),
    "qtx_pinv_rational_workspace_size": (_sz, [_i64]),
    "qtx_sym_absmax_eig": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _sz, _vp]),
    "qtx_pinv_rational_partial": (_i32, [_vp, _i64, _vp, _f64, _f64, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "qtx_dd_sum_scale": (_i32, [_vp, _i32, _i64, _f64, _vp, _vp]),
    "qtx_resconv_sweep_workspace_size": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32, _i32, _i32]),
    "qtx_resconv_sweep": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _vp, _i32,
                                 _i32, _f64, _u64, _u64, _u64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "qtx_comm_unique_id": (_i32, [_vp]),
    "qtx_comm_init": (_i32, [_vp, _i32, _i32, _vp]),
    "qtx_comm_adopt": (_i32, [_vp, _vp]),
    "qtx_comm_destroy": (_i32, [_vp]),
    "qtx_comm_size": (_i32, [_vp]),
    "qtx_comm_rank": (_i32, [_vp]),
    "qtx_comm_all_reduce": (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "qtx_comm_all_gather": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "qtx_comm_broadcast": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "qtx_comm_all_to_all": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "qtx_minsr_solve_dist_workspace_size": (_sz, [_vp, _i32, _i64, _i64, _i32]),
    "qtx_minsr_solve_dist_timing": (_i32, [_i32]),
    "qtx_minsr_solve_dist_phases": (_i32, [_vp]),
    "qtx_minsr_solve_dist": (_i32, [_vp, _i32, _vp, _i64, _i64, _i64, _vp, _f64, _f64, _i32, _i32, _i32, _vp, _vp, _vp,
                                    _sz, _vp]),
    "qtx_pinv_ldlt_workspace_size": (_sz, [_i64, _i32]),
    "qtx_sym_absmax_eig_ws": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _sz, _i32, _vp]),
    "qtx_pinv_ldlt_partial": (_i32, [_vp, _i64, _vp, _f64, _f64, _vp, _i32, _i32, _vp, _i32, _vp, _vp, _sz, _vp]),
    "qtx_shift_chol_workspace_size": (_sz, [_i64]),
    "qtx_shift_chol_solve": (_i32, [_vp, _i64, _vp, _f64, _f64, _vp, _vp, _vp, _sz, _vp]),
    "qtx_col_sumsq": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp]),
    "qtx_matvec_t": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp, _i32, _vp]),
    "qtx_matvec": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "qtx_axpby": (_i32, [_i64, _f64, _vp, _f64, _vp, _vp]),
    "qtx_rank1_update": (_i32, [_i64, _f64, _vp, _vp, _vp]),
    "qtx_div_add": (_i32, [_i64, _vp, _vp, _f64, _vp, _vp, _vp]),
    "qtx_second_moment": (_i32, [_i64, _f64, _vp, _vp, _vp, _vp]),
    "qtx_fourth_root": (_i32, [_i64, _vp, _f64, _f64, _vp, _vp]),
    "qtx_scale_columns": (_i32, [_i32, _vp, _i64, _i64, _i64, _vp, _vp]),
    "qtx_apply_update": (_i32, [_i32, _vp, _vp, _f64, _i64, _vp, _vp]),
    "qtx_resconv_forward_cplx": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _vp,
                                        _sz, _vp]),
    "qtx_resconv_jacobian_cplx": (_i32, [_i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i32, _vp,
                                         _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "qtx_apply_sign_phase": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "qtx_metropolis_accept_cplx": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _f64, _vp, _u64, _u64, _u64,
                                          _vp, _vp, _vp]),
    "qtx_oloc_reduce_cplx": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp]),
    "qtx_symm_combine_cplx": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "qtx_weighted_rowsum_cplx": (_i32, [_i32, _vp, _i64, _i64, _vp, _i64, _i32, _i64, _vp, _i64, _i64, _vp]),
    "qtx_ebar_cplx": (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "qtx_real_to_cplx": (_i32, [_vp, _i64, _vp, _vp]),
    "qtx_rbm_conv_expand": (_i32, [_i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "qtx_rbm_conv_jacobian": (_i32, [_i32, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
}

_lib = None


class QtxError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C quantax_b200/csrc`).  quantax_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def call(name, *args):
    """Call an int-returning entry point and raise QtxError with the library message on failure."""
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        raise QtxError(f"{name} failed ({rc}): {L.qtx_last_error().decode()}")


def ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise QtxError("quantax_b200 kernels need CUDA tensors; there is no CPU path")
    if not t.is_contiguous():
        raise QtxError("non-contiguous tensor passed to the C ABI")
    return t.data_ptr()


def ptr2d(t):
    """Row-major matrix whose rows may be padded (leading dimension = t.stride(0))."""
    if t is None:
        return None
    if not t.is_cuda:
        raise QtxError("quantax_b200 kernels need CUDA tensors; there is no CPU path")
    if t.ndim != 2 or t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        raise QtxError("expected a row-major matrix with unit column stride")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(dt):
    if dt == torch.float32:
        return QTX_F32
    if dt == torch.float64:
        return QTX_F64
    raise QtxError(f"unsupported dtype {dt}")
