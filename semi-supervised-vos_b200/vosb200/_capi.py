"""ctypes binding of include/vos_prop.h.  No fallback: if libvosprop.so is missing this raises."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAX_REFS = 32
MAX_CLASSES = 24
MAX_DENSE_CLASSES = 14
MAX_TOPK = 64
FEAT_DIM = 256
F32, F16, BF16 = 0, 1, 2
NCHW, NHWC = 0, 1
KERNEL_TC, KERNEL_SIMT, KERNEL_TC_DENSE = 0, 1, 2
PREC_SPLIT3, PREC_F16, PREC_BF16 = 0, 1, 2
ABI_VERSION = 4
OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4

# VOS_LIB_NAME selects a variant build of the same sources (vosb200/build.py); the product is libvosprop.so
LIB_PATH = Path(__file__).resolve().parent.parent / 'csrc' / os.environ.get('VOS_LIB_NAME', 'libvosprop.so')


class Config(C.Structure):
    _fields_ = [('device', C.c_int32), ('max_pixels', C.c_int32), ('ring_slots', C.c_int32),
                ('max_fullres_pixels', C.c_int32)]


class Step(C.Structure):
    _fields_ = [('frame_idx', C.c_int32), ('n_refs', C.c_int32),
                ('ref_frames', C.c_int32 * MAX_REFS), ('ref_sigma', C.c_float * MAX_REFS),
                ('temperature', C.c_float), ('probability_propagation', C.c_int32),
                ('write_labels', C.c_int32), ('topk', C.c_int32), ('kernel', C.c_int32),
                ('out_prediction', C.c_void_p), ('out_mask_lowres', C.c_void_p),
                ('out_mask_fullres', C.c_void_p), ('out_topk_idx', C.c_void_p),
                ('wait_event', C.c_void_p), ('record_event', C.c_void_p)]


class VosPropError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f'libvosprop error {code}: {msg}')
        self.code = code


EXPORTS = {
    'vosprop_last_error': (C.c_char_p, []),
    'vosprop_abi_version': (C.c_int, []),
    'vosprop_create': (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    'vosprop_destroy': (None, [C.c_void_p]),
    'vosprop_reset': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'vosprop_append_features': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    'vosprop_block_skip': (C.c_int, [C.c_void_p, C.c_int32]),
    'vosprop_block_skip_state': (C.c_int, [C.c_void_p]),
    'vosprop_normalize_u8': (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_int32,
                                       C.c_void_p]),
    'vosprop_append_frames': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    'vosprop_set_labels_index': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'vosprop_set_labels_dense': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    'vosprop_propagate': (C.c_int, [C.c_void_p, C.POINTER(Step), C.c_void_p]),
    'vosprop_sample_frames': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    'vosprop_plan_step': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.POINTER(Step)]),
    'vosprop_ring_slots': (C.c_int, [C.c_void_p]),
    'vosprop_num_sms': (C.c_int, [C.c_void_p]),
    'vosprop_debug_decompose': (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    'vosprop_debug_flags': (C.c_int, [C.c_void_p, C.c_int32]),
    'vosprop_debug_clocks': (C.c_int, [C.c_void_p, C.c_void_p]),
    'vosprop_launch_count': (C.c_int64, [C.c_void_p]),
    'vosprop_timing_enable': (C.c_int, [C.c_void_p, C.c_int32]),
    'vosprop_timing_select': (C.c_int, [C.c_void_p, C.c_int32]),
    'vosprop_timing_read': (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.is_file():
            raise ImportError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; '
                              f'g.build()"` (there is no CPU fallback)')
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.vosprop_abi_version() != ABI_VERSION:
            raise ImportError('libvosprop ABI version mismatch; rebuild')
        _lib = handle
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        raise VosPropError(rc, lib().vosprop_last_error().decode())
    return rc
