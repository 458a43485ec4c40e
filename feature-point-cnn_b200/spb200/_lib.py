"""ctypes binding of libspb200.so (the C ABI declared in include/spb200.h).

There is no fallback: if the library has not been built (``python feature-point-cnn_b200/build.py``)
importing this module raises, and creating an engine without a Blackwell GPU fails in the library.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'libspb200.so')

# every symbol include/spb200.h declares: name -> (restype, argtypes)
_c = ctypes
_P = _c.c_void_p
SYMBOLS = {
    'spb200_create': (_c.c_int, [_c.c_int, _c.POINTER(_P)]),
    'spb200_debug_halo': (_c.c_int, [_P]),
    'spb200_destroy': (None, [_P]),
    'spb200_last_error': (_c.c_char_p, [_P]),
    'spb200_load_checkpoint': (_c.c_int, [_P, _c.c_char_p]),
    'spb200_load_tensor': (_c.c_int, [_P, _c.c_char_p, _P, _c.POINTER(_c.c_int64), _c.c_int]),
    'spb200_finalize_weights': (_c.c_int, [_P, _c.c_int]),
    'spb200_finalize_weights_split': (_c.c_int, [_P, _c.c_int, _c.c_int]),
    'spb200_set_params': (_c.c_int, [_P, _c.c_float, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    'spb200_forward': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'spb200_detect': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P, _P, _P]),
    'spb200_detect_host': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'spb200_detect_u8': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P, _P, _P]),
    'spb200_detect_host_u8': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'spb200_detect_host_submit': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P, _c.POINTER(_c.c_int)]),
    'spb200_detect_host_wait': (_c.c_int, [_P, _c.c_int]),
    'spb200_set_descriptor_format': (_c.c_int, [_P, _c.c_int]),
    'spb200_homography_adaptation': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    'spb200_match': (_c.c_int, [_P, _P, _P, _P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _P, _P, _P]),
    'spb200_heatmap_from_logits': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    'spb200_restore_prob_map': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _P]),
    'spb200_preprocess_u8': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_int, _c.c_int, _P]),
    'spb200_preprocess_f32': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _P, _c.c_int, _c.c_int, _P]),
    'spb200_nms': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'spb200_sample_descriptors': (_c.c_int, [_P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'spb200_descriptor_dim': (_c.c_int, [_P]),
    'spb200_max_keypoints': (_c.c_int, [_c.c_int, _c.c_int, _c.c_int]),
    'spb200_kernel_launches': (_c.c_long, [_P]),
    'spb200_reset_kernel_launches': (None, [_P]),
    'spb200_export_activation': (_c.c_int, [_P, _c.c_int, _P, _c.c_int, _P]),
    'spb200_activation_dims': (_c.c_int, [_P, _c.c_int, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int)]),
    'spb200_profile_begin': (_c.c_int, [_P]),
    'spb200_profile_end': (_c.c_int, [_P, _c.c_int, _P, _P, _P, _P, _c.POINTER(_c.c_int)]),
    'spb200_checkpoint_num_tensors': (_c.c_int, [_c.c_char_p]),
    'spb200_checkpoint_tensor': (_c.c_int, [_c.c_char_p, _c.c_char_p, _P, _c.c_long, _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int)]),
    'spb200_test_conv_tc': (_c.c_int, [_c.c_int, _P, _P, _P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                       _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P]),
    'spb200_test_conv_kernel': (_c.c_int, [_c.c_int, _c.c_int, _P, _P, _P, _P, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                           _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P]),
}

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError('libspb200.so is missing (%s): build it with `python feature-point-cnn_b200/build.py`; '
                              'there is no CPU or PyTorch fallback' % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
