"""ctypes binding of libdisyolo_b200.so (C ABI declared in include/disyolo.h).

Fails loudly: if the shared library has not been built the first use raises RuntimeError -- there
is no Python / CPU fallback for any compute entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DISYOLO_LIB: A/B aid (scripts/): load an alternative build of the same C ABI
LIB_PATH = os.environ.get('DISYOLO_LIB') or os.path.join(_HERE, 'libdisyolo_b200.so')

PRECISION_BF16 = 0
PRECISION_FP32 = 1
MASKS_NONE, MASKS_FULL, MASKS_CROPPED = 0, 1, 2      # enum dy_mask_mode


class DyConfig(C.Structure):
    """struct dy_config (include/disyolo.h)."""
    _fields_ = [
        ('num_classes', C.c_int32),
        ('anchors', C.c_float * 18),
        ('image_size', C.c_int32),
        ('k_map', C.c_int32),
        ('alpha', C.c_float),
        ('bn_eps', C.c_float),
        ('iou_threshold', C.c_float),
        ('max_detection', C.c_int32),
        ('max_batch', C.c_int32),
        ('precision', C.c_int32),
        ('device', C.c_int32),
        ('lock', C.c_uint8 * 82),
        ('reserved', C.c_uint8 * 2),
    ]


_P = C.c_void_p
_I = C.c_int32
_F = C.c_float

# name -> (restype, argtypes); every symbol include/disyolo.h declares
SIGNATURES = {
    'dy_version': (C.c_char_p, []),
    'dy_last_error': (C.c_char_p, []),
    'dy_device_count': (C.c_int, []),
    'dy_create': (C.c_int, [C.POINTER(DyConfig), C.POINTER(_P)]),
    'dy_destroy': (C.c_int, [_P]),
    'dy_load_weights': (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    'dy_finalize_weights': (C.c_int, [_P]),
    'dy_forward': (C.c_int, [_P, _P, _I, _P, _F, _P, _P, _P, _P, _P]),
    'dy_forward_host': (C.c_int, [_P, _P, _I, _P, _F, _P, _P, _P, _P]),
    'dy_forward_host_begin': (C.c_int, [_P, _P, _I, _P, _F, _I, C.POINTER(_I)]),
    'dy_forward_host_end': (C.c_int, [_P, _I, _P, _P, _P, _P]),
    'dy_forward_host_begin_u8': (C.c_int, [_P, _P, _I, _P, _F, _I, C.POINTER(_I)]),
    'dy_forward_host_end_cropped': (C.c_int, [_P, _I, _P, _P, _P, _P, _P, C.c_int64]),
    'dy_forward_network': (C.c_int, [_P, _P, _I, _P]),
    'dy_forward_profile': (C.c_int, [_P, _P, _I, _P, _P]),
    'dy_layer_shape': (C.c_int, [_P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    'dy_get_activation': (C.c_int, [_P, _I, _I, _P, _P]),
    'dy_get_yolo': (C.c_int, [_P, _I, _I, _P, _P]),
    'dy_get_mask_pos': (C.c_int, [_P, _I, _P, _P]),
    'dy_decode': (C.c_int, [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    'dy_detect': (C.c_int, [_P, _P, _P, _P, _I, _P, _F, _P, _P, _P, _P]),
    'dy_nms': (C.c_int, [_P, _P, _P, _P, _I, _I, _F, _P, _P, _P, _P]),
    'dy_assemble_masks': (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P]),
    'dy_letterbox': (C.c_int, [_P, _I, _I, _I, _P, _P, _P]),
    'dy_postprocess': (C.c_int, [_P, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'dy_letterbox_batch': (C.c_int, [_P, C.c_int64, _I, _I, _I, _I, _P, _P, _P]),
    'dy_postprocess_batch': (C.c_int, [_P, _P, _I, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'dy_mask_overlaps': (C.c_int, [_P, _I, _P, _I, C.c_int64, _P, _P]),
    'dy_assign_labels': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    'dy_polygon_masks': (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    'dy_mask_boxes': (C.c_int, [_P, _I, _I, _I, _P, _P]),
    'dy_augment_image': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'dy_augment_masks': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'dy_salt_pepper': (C.c_int, [_P, _I, _P, _I, _P, _I, _P]),
    'dy_change_light': (C.c_int, [_P, C.c_int64, C.c_double, _P]),
    'dy_motion_blur3': (C.c_int, [_P, _I, _P, _P, _P]),
    'dy_u8_to_unit_float': (C.c_int, [_P, _P, C.c_int64, _P]),
    'dy_postproc_profile': (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _P, _F, _P, _I, _P, _P]),
    'dy_conv_layer': (C.c_int, [_I, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _I, _F, _P, _P, _P]),
    'dy_conv_backward': (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P]),
    'dy_train_init': (C.c_int, [_P]),
    'dy_set_loss_params': (C.c_int, [_P, _F, _F, _F, _F, _F, _F]),
    'dy_train_param_count': (C.c_int64, [_P]),
    'dy_train_layer_span': (C.c_int, [_P, _I, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'dy_train_forward': (C.c_int, [_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P]),
    'dy_train_backward': (C.c_int, [_P, _I, _I, _I, _P, _P]),
    'dy_train_apply': (C.c_int, [_P, _P, _F, _F, _P]),
    'dy_train_get_tensor': (C.c_int, [_P, _I, _I, _I, _P, _P]),
    'dy_get_weights': (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    'dy_crc32c': (C.c_uint32, [_P, C.c_uint64, C.c_uint32]),
    'dy_set_option': (C.c_int, [C.c_char_p, _I]),
    'dy_launch_count': (C.c_int64, [_I]),
}

_lib = None


class DisYoloError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                'libdisyolo_b200.so is not built (%s). Run `python -c "import __graft_entry__ as g; '
                'g.build()"` or `make -C dis-yolo_b200/csrc`. There is no CPU fallback.' % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what=''):
    if rc != 0:
        msg = lib().dy_last_error().decode('utf-8', 'replace')
        raise DisYoloError('%s failed (status %d): %s' % (what or 'libdisyolo_b200 call', rc, msg))


def require_gpu():
    if lib().dy_device_count() <= 0:
        raise DisYoloError('no CUDA device visible: disyolo_b200 has no CPU fallback')
