"""ctypes binding of liberd_b200.so (include/erd_b200.h).  No torch types cross this line.

The product path has no CPU fallback: if the shared library is missing it is built with
nvcc (``erd_b200/build.py``), and if that fails the import error propagates.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

MAX_LEVELS = 5
ABI_VERSION = 6


class ErdShape(C.Structure):
    _fields_ = [('num_imgs', C.c_int32), ('num_levels', C.c_int32), ('num_classes', C.c_int32),
                ('ori_classes', C.c_int32), ('reg_max', C.c_int32),
                ('level_h', C.c_int32 * MAX_LEVELS), ('level_w', C.c_int32 * MAX_LEVELS),
                ('stride', C.c_int32 * MAX_LEVELS), ('total_gt', C.c_int32),
                ('anchor_scale', C.c_float), ('loss_weight_cls', C.c_float),
                ('loss_weight_bbox', C.c_float), ('loss_weight_dfl', C.c_float),
                ('loss_weight_ld', C.c_float), ('kd_temperature', C.c_float),
                ('max_gt_per_img', C.c_int32)]


class ErdSizes(C.Structure):
    _fields_ = [('anchors_per_img', C.c_int64), ('sel_cap', C.c_int64), ('num_losses', C.c_int64),
                ('workspace_bytes', C.c_size_t)]


class ErdPredictConfig(C.Structure):
    _fields_ = [('nms_pre', C.c_int32), ('max_per_img', C.c_int32), ('score_thr', C.c_float),
                ('iou_threshold', C.c_float), ('min_bbox_size', C.c_float)]


class ErdStepBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ('cls_inds', 'cls_count', 'box_inds', 'box_count', 'thr',
                                           'sel_flags', 'gt_inds', 'num_pos', 'keep', 'keep_count', 'avg')]


class ErdTeacherHead(C.Structure):
    _fields_ = [('w_cls', C.c_void_p), ('w_reg', C.c_void_p), ('b_cls', C.c_void_p), ('b_reg', C.c_void_p),
                ('scale', C.c_float * MAX_LEVELS)]


PtrArray = C.c_void_p * MAX_LEVELS
_P, _I, _F = C.c_void_p, C.c_int32, C.c_float
_SH = C.POINTER(ErdShape)

# name -> argtypes; every function returns int (ErdStatus) unless noted
SIGNATURES = {
    'erd_abi_version': [],
    'erd_last_error': [],
    'erd_sizes': [_SH, C.POINTER(ErdSizes)],
    'erd_workspace_init': [_SH, _P, _P],
    'erd_predict_workspace_bytes': [_SH, C.POINTER(ErdPredictConfig), C.POINTER(C.c_size_t)],
    'erd_predict': [_SH, C.POINTER(ErdPredictConfig), PtrArray, PtrArray, _P, _P, _P, _P, _P, _P, _P],
    'erd_workspace_field': [_SH, _P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_size_t)],
    'erd_create': [C.POINTER(_P)],
    'erd_destroy': [_P],
    'erd_ers_select': [_SH, PtrArray, PtrArray, _P, _P, _P, _P, _P, _P, _P, _P],
    'erd_teacher_head_packed_floats': [_I],
    'erd_teacher_head_pack': [_P, _I, _P, _P],
    'erd_teacher_head_fused': [_SH, C.POINTER(ErdTeacherHead), PtrArray, PtrArray, C.POINTER(PtrArray), C.POINTER(PtrArray),
                               _P, _P, _P, _P],
    'erd_ers_select_cached': [_SH, _P, _P, _P, _P, _P, _P, _P, _P],
    'erd_atss_assign': [_SH, _P, _P, _P, _P, _P, _P, _P, _P],
    'erd_avg_factors': [_SH, PtrArray, PtrArray, _P, _P, _P, _P, _P, _P, _P, _P],
    'erd_teacher_nms': [_SH, _P, _P, _P, _F, _P, _P, _P, _P, _P],
    'erd_loss_fwd_bwd': [_P, _SH, PtrArray, PtrArray, PtrArray, PtrArray, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                         _P, _P, _P, _F, _P, _I, _P, PtrArray, PtrArray, _P, _P],
    'erd_step_prepare': [_P, _SH, PtrArray, PtrArray, PtrArray, PtrArray, _P, _P, _P, _P, _F,
                         C.POINTER(ErdStepBuffers), _P, _P, C.c_uint32],
    'erd_avg_exchange_bytes': [],
    'erd_avg_exchange': [_P, C.POINTER(_P), _I, _I, _P],
    'erd_context_set_exchange': [_P, C.POINTER(_P), _I, _I],
    'erd_profile_enable': [C.c_uint],
    'erd_launch_count': [],
    'erd_profile_num_kernels': [],
    'erd_profile_kernel_name': [C.c_int],
    'erd_profile_collect': [C.POINTER(C.c_float), C.POINTER(C.c_int)],
    'erd_profile_mark': [_P],
    'erd_profile_timeline': [C.POINTER(C.c_float), C.POINTER(C.c_float)],
}

_lib = None


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the sources are newer) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get('ERD_B200_NO_BUILD') != '1':
        _build.build()
    lib = C.CDLL(_build.LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == header/library drift
        fn.argtypes = argtypes
        fn.restype = (C.c_char_p if name in ('erd_last_error', 'erd_profile_kernel_name')
                      else C.c_ulonglong if name == 'erd_launch_count'
                      else C.c_size_t if name in ('erd_avg_exchange_bytes', 'erd_teacher_head_packed_floats') else C.c_int)
    if lib.erd_abi_version() != ABI_VERSION:
        raise RuntimeError(f'liberd_b200 ABI {lib.erd_abi_version()} != binding {ABI_VERSION}')
    _lib = lib
    return lib


class ErdError(RuntimeError):
    pass


def check(rc: int, what: str):
    if rc != 0:
        msg = load().erd_last_error()
        raise ErdError(f'{what} failed with status {rc}: {msg.decode() if msg else ""}')
