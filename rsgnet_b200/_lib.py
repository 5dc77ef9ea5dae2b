"""ctypes binding of librsg_b200.so (C ABI declared in include/rsg_b200.h).

The library is the only compute path of this package: if it is missing or a call fails, the
caller gets an exception -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librsg_b200.so')

MAX_TAPS = 16
MAX_RES = 4


class Ref(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('ext_slot', C.c_int32), ('offset', C.c_int64),
                ('crop_stride', C.c_int64)]


class Res(C.Structure):
    _fields_ = [('src', Ref), ('cs', C.c_int32), ('co', C.c_int32), ('H', C.c_int32),
                ('W', C.c_int32), ('shift', C.c_int32), ('batch_stride0', C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [('inp', Ref), ('in_cs', C.c_int32), ('in_co', C.c_int32), ('Hin', C.c_int32),
                ('Win', C.c_int32), ('Cin', C.c_int32),
                ('w', Ref), ('w_tc5', Ref), ('bias', Ref), ('Cout', C.c_int32), ('CoutPad', C.c_int32),
                ('ntaps', C.c_int32), ('tap_dy', C.c_int8 * MAX_TAPS), ('tap_dx', C.c_int8 * MAX_TAPS),
                ('stride', C.c_int32), ('Hout', C.c_int32), ('Wout', C.c_int32),
                ('out', Ref), ('out_cs', C.c_int32), ('out_co', C.c_int32), ('oH', C.c_int32),
                ('oW', C.c_int32), ('omul', C.c_int32), ('ooy', C.c_int32), ('oox', C.c_int32),
                ('out_f32', Ref), ('nres', C.c_int32), ('res', Res * MAX_RES),
                ('relu', C.c_int32), ('engine', C.c_int32), ('pixel_shuffle_c', C.c_int32)]


def null_ref():
    return Ref(None, -1, 0, 0)


def abs_ref(ptr):
    return Ref(int(ptr) if ptr else None, -1, 0, 0)


def ext_ref(slot, offset=0, crop_stride=0):
    return Ref(None, int(slot), int(offset), int(crop_stride))


class PermEntry(C.Structure):
    """rsg_perm_entry (include/rsg_b200.h)."""
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('D0', C.c_int32), ('D1', C.c_int32), ('D2', C.c_int32),
                ('V1', C.c_int32), ('V2', C.c_int32), ('s0', C.c_longlong), ('s1', C.c_longlong), ('s2', C.c_longlong),
                ('accumulate', C.c_int32)]


class RsgError(RuntimeError):
    pass


_lib = None


def use_library(path):
    """Bind another build of the library (tools/ experiments use librsg_b200_dbg.so = `make DEBUG_SWITCHES=1`, which keeps
    the RSG_* environment switches of the kernels).  Must be called before the first lib() call."""
    global LIB_PATH
    if _lib is not None:
        raise RsgError('use_library: the library is already loaded')
    LIB_PATH = os.path.abspath(path)

_SIGS = {
    'rsg_abi_version': (C.c_int, []),
    'rsg_last_error': (C.c_char_p, []),
    'rsg_device_info': (C.c_int, [C.POINTER(C.c_int)]),
    'rsg_flip_avg_decode': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p] * 2 +
                            [C.c_int] * 2 + [C.c_void_p] * 4),
    'rsg_flip_back': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4),
    'rsg_oks_nms': (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
    'rsg_soft_oks_nms': (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
    'rsg_oks_iou': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double]),
    'rsg_rescore': (C.c_int, [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_double, C.c_void_p]),
    'rsg_evaluate_workspace_bytes': (C.c_int, [C.c_int, C.POINTER(C.c_size_t)]),
    'rsg_evaluate': (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                               C.c_void_p, C.c_size_t] + [C.c_void_p] * 6),
    'rsg_warp_affine': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 4 + [C.c_void_p] * 3),
    'rsg_plan_create': (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    'rsg_plan_destroy': (None, [C.c_void_p]),
    'rsg_plan_num_ops': (C.c_int, [C.c_void_p]),
    'rsg_plan_add_stem': (C.c_int, [C.c_void_p, Ref, C.c_int, C.c_int, Ref, Ref, Ref]),
    'rsg_plan_add_conv': (C.c_int, [C.c_void_p, C.POINTER(ConvDesc)]),
    'rsg_plan_add_fuse': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(Res), Ref] + [C.c_int] * 6),
    'rsg_plan_add_maxpool': (C.c_int, [C.c_void_p, Ref] + [C.c_int] * 5 + [Ref]),
    'rsg_plan_add_attention': (C.c_int, [C.c_void_p, Ref, C.c_int, C.c_int, Ref, C.c_int, C.c_int,
                                         Ref, C.c_int, C.c_int, C.c_int, C.c_int]),
    'rsg_plan_add_attention_f32': (C.c_int, [C.c_void_p, Ref, C.c_int, C.c_int, Ref, C.c_int, C.c_int,
                                             Ref, C.c_int, C.c_int, Ref, C.c_int, C.c_int]),
    'rsg_plan_add_trp_tail': (C.c_int, [C.c_void_p, Ref, Ref, Ref, Ref, Ref, C.c_int, C.c_float, Ref, C.c_int, C.c_int,
                                        C.c_int, C.c_int]),
    'rsg_plan_add_relation_scores': (C.c_int, [C.c_void_p, Ref] + [C.c_int] * 4 + [Ref]),
    'rsg_plan_add_groupnorm': (C.c_int, [C.c_void_p, Ref, C.c_int, C.c_int, Ref, Ref, C.c_int,
                                         C.c_float, Ref, C.c_int, C.c_int, C.c_int, C.c_int]),
    'rsg_basic_block_supported': (C.c_int, [C.c_int] * 3),
    'rsg_plan_add_basic_block': (C.c_int, [C.c_void_p, Ref, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, Ref, Ref, Ref, Ref,
                                           Ref, C.c_int, C.c_int]),
    'rsg_bottleneck_supported': (C.c_int, [C.c_int] * 5),
    'rsg_plan_add_bottleneck': (C.c_int, [C.c_void_p, Ref] + [C.c_int] * 5 + [Ref] * 7 + [C.c_int, C.c_int, Ref, C.c_int, C.c_int]),
    'rsg_plan_add_bilinear2x': (C.c_int, [C.c_void_p, Ref, Ref] + [C.c_int] * 4),
    'rsg_plan_begin_aux': (C.c_int, [C.c_void_p]),
    'rsg_plan_run': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                               C.c_int, C.c_int, C.c_int]),
    'rsg_plan_last_launches': (C.c_int, [C.c_void_p]),
    'rsg_plan_profile': (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    'rsg_conv_tc5_config': (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 3),
    'rsg_conv_ws_config': (C.c_int, [C.c_int] * 5 + [C.POINTER(C.c_int)]),
    'rsg_conv_ws_config2': (C.c_int, [C.c_int] * 6 + [C.POINTER(C.c_int)]),
    'rsg_conv_ws2_config': (C.c_int, [C.c_int] * 5),
    'rsg_conv_run': (C.c_int, [C.c_void_p, C.POINTER(ConvDesc), C.c_int]),
    # ---- training step (include/rsg_b200.h 'Training step') ----
    'rsg_train_gemm': (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 7 + [C.c_longlong] * 3 + [C.c_int] * 4 +
                       [C.POINTER(C.c_int), C.c_int]),
    'rsg_train_wgrad': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 6 + [C.POINTER(C.c_int), C.c_int]),
    'rsg_train_bn_fwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_float, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4),
    'rsg_train_bn_bwd': (C.c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_int] + [C.c_void_p] * 3 + [C.c_int] +
                         [C.c_void_p] * 4),
    'rsg_train_bn_fwd_res': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                       C.c_float, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5),
    'rsg_train_bn_bwd_res': (C.c_int, [C.c_void_p] * 4 + [C.c_longlong, C.c_int] + [C.c_void_p] * 3 + [C.c_int] +
                             [C.c_void_p] * 5),
    'rsg_train_colsum': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'rsg_train_gn_fwd': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_float] +
                         [C.c_void_p] * 3),
    'rsg_train_gn_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 6),
    'rsg_train_add': (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_void_p]),
    'rsg_train_ew': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_longlong, C.c_void_p]),
    'rsg_train_copy2d': (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_longlong,
                                   C.c_int, C.c_int]),
    'rsg_train_permute3': (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_longlong] * 3 + [C.c_int] * 3),
    'rsg_train_permute3_batch': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    'rsg_train_resample': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + [C.c_void_p]),
    'rsg_train_maxpool': (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    'rsg_train_mse_joints': (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_float, C.c_void_p, C.c_void_p]),
    'rsg_train_bce': (C.c_int, [C.c_void_p] * 3 + [C.c_longlong, C.c_double, C.c_float, C.c_void_p, C.c_void_p]),
    'rsg_train_relation_mse': (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p]),
    'rsg_train_trp_dscore': (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p]),
    'rsg_train_person_mask': (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]),
    'rsg_train_zero': (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    'rsg_train_d2f': (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]),
    'rsg_train_adam': (C.c_int, [C.c_void_p] * 5 + [C.c_longlong] + [C.c_float] * 4 + [C.c_int, C.c_float]),
    'rsg_train_adam_graph': (C.c_int, [C.c_void_p] * 5 + [C.c_longlong] + [C.c_float] * 4 + [C.c_void_p, C.c_float]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Load librsg_b200.so (once).  Raises if it has not been built: there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RsgError(f'{LIB_PATH} is missing: build it with `python __graft_entry__.py` '
                           '(or make -C rsgnet_b200/csrc); rsgnet_b200 has no CPU fallback')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.rsg_abi_version() != 1:
            raise RsgError('librsg_b200.so ABI version mismatch')
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RsgError(lib().rsg_last_error().decode('utf-8', 'replace') or f'librsg_b200 error {rc}')


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RsgError('rsgnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback')


def stream_ptr(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
