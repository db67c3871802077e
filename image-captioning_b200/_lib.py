"""ctypes binding of libdcap.so (the C ABI declared in include/dcap.h).

There is NO CPU fallback: if the shared library is missing or a CUDA device is not usable the
calls raise -- the oracle under ``oracle/`` is test infrastructure and is never imported here.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdcap.so")

_lib = None
_lock = threading.Lock()


class DcapError(RuntimeError):
    """Error reported by libdcap.so (negative return code + dc_last_error message)."""

    def __init__(self, code, msg):
        super().__init__("libdcap error %d: %s" % (code, msg))
        self.code = code


c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_void = ctypes.c_void_p

# name -> (restype, argtypes); kept in one table so tests can check it against include/dcap.h
SIGNATURES = {
    "dc_last_error": (ctypes.c_char_p, []),
    "dc_device_info": (ctypes.c_int, [c_int_p, c_int_p]),
    "dc_fpn_levels_f32": (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         c_void, c_void]),
    "dc_pyramid_roi_align_f32": (ctypes.c_int, [c_void, c_void * 4, ctypes.c_int * 4,
                                                ctypes.c_int * 4] + [ctypes.c_int] * 7 +
                                 [c_void, c_void, c_void]),
    "dc_pyramid_roi_align_bf16out": (ctypes.c_int, [c_void, c_void * 4, ctypes.c_int * 4,
                                                    ctypes.c_int * 4] + [ctypes.c_int] * 7 +
                                     [c_void, c_void, c_void]),
    "dc_pyramid_roi_align_host_f32": (ctypes.c_int, [c_void, c_void * 4, ctypes.c_int * 4,
                                                     ctypes.c_int * 4] + [ctypes.c_int] * 7 +
                                      [c_void, c_void]),
    "dc_pyramid_roi_align_backward_f32": (ctypes.c_int, [c_void, c_void, c_void * 4, ctypes.c_int * 4, ctypes.c_int * 4] +
                                          [ctypes.c_int] * 7 + [c_void]),
    "dc_decoder_create": (ctypes.c_int, [c_void, c_void]),
    "dc_decoder_destroy": (ctypes.c_int, [c_void]),
    "dc_decoder_weight_count": (ctypes.c_int, [c_void]),
    "dc_decoder_weight_name": (ctypes.c_char_p, [c_void, ctypes.c_int]),
    "dc_decoder_weight_numel": (ctypes.c_int64, [c_void, ctypes.c_int]),
    "dc_decoder_set_weight": (ctypes.c_int, [c_void, ctypes.c_char_p, c_void, ctypes.c_int64]),
    "dc_decoder_get_weight": (ctypes.c_int, [c_void, ctypes.c_char_p, c_void, ctypes.c_int64]),
    "dc_decoder_finalize": (ctypes.c_int, [c_void, c_void]),
    "dc_head_forward": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void]),
    "dc_decoder_greedy": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void, c_void]),
    "dc_decoder_greedy_scored": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void, c_void]),
    "dc_refine_generations": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                             c_void, c_void, c_void]),
    "dc_proposal_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "dc_proposal_layer": (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, ctypes.c_float,
                                         ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_float, c_void, c_void, c_void,
                                         c_void, ctypes.c_size_t, c_void]),
    "dc_normalize_boxes": (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_float, ctypes.c_float, c_void, c_void]),
    "dc_decoder_beam": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void,
                                       c_void, c_void]),
    "dc_decoder_v2_predict": (ctypes.c_int, [c_void, c_void, ctypes.c_int, c_void, ctypes.c_int,
                                             ctypes.c_int, c_void, c_void]),
    "dc_decoder_v2_greedy": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void,
                                            c_void]),
    "dc_decoder_v2_greedy_from": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void, c_void,
                                                 c_void, c_void]),
    "dc_gemm_f32": (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int, c_void, ctypes.c_int64, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void, c_void, ctypes.c_int64,
                                   ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int64, c_void]),
    "dc_gemm_bf16": (ctypes.c_int, [c_void, ctypes.c_int64, c_void, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, c_void, c_void, ctypes.c_int64, ctypes.c_int, c_void,
                                    ctypes.c_int64, c_void, ctypes.c_int64, c_void]),
    "dc_gemm_bf16_ex": (ctypes.c_int, [c_void, ctypes.c_int64, ctypes.c_int, c_void, ctypes.c_int64, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void, c_void, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int64, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int64, c_void, ctypes.c_int64,
                                       c_void]),
    "dc_gemm_bf16_argmax": (ctypes.c_int, [c_void, ctypes.c_int64, c_void, ctypes.c_int64, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, c_void, c_void, c_void, c_void]),
    "dc_gemm_bf16_topk": (ctypes.c_int, [c_void, ctypes.c_int64, c_void, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, c_void, ctypes.c_int, c_void, c_void, c_void]),
    "dc_gemm_bf16_lstm_cell": (ctypes.c_int, [c_void, ctypes.c_int64, c_void, ctypes.c_int64, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int64, c_void, c_void,
                                              c_void, c_void, ctypes.c_int64, c_void, ctypes.c_int64, c_void,
                                              ctypes.c_int64, c_void]),
    "dc_caption_rois": (ctypes.c_int, [c_void, c_void, c_void * 4, ctypes.c_int * 4, ctypes.c_int * 4] +
                        [ctypes.c_int] * 4 + [c_void, c_void]),
    "dc_caption_rois_host": (ctypes.c_int, [c_void, c_void, c_void * 4, ctypes.c_int * 4, ctypes.c_int * 4] +
                             [ctypes.c_int] * 4 + [c_void]),
    "dc_caption_rois_host_submit": (ctypes.c_int, [c_void, c_void, c_void * 4, ctypes.c_int * 4, ctypes.c_int * 4] +
                                    [ctypes.c_int] * 4 + [c_void]),
    "dc_caption_rois_host_wait": (ctypes.c_int, [c_void]),
    "dc_decoder_greedy_host": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void]),
    "dc_decoder_train_step": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void,
                                             ctypes.c_float, c_void, c_void]),
    "dc_decoder_train_step_ex": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void,
                                                ctypes.c_float, c_void, c_void, c_void]),
    "dc_decoder_v2_train_step": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int, c_void,
                                                ctypes.c_float, c_void, c_void]),
    "dc_decoder_teacher_forced": (ctypes.c_int, [c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_void, c_void]),
    "dc_adam_step": (ctypes.c_int, [c_void, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_int, ctypes.c_int64, ctypes.c_float, c_void]),
    "dc_adam_step_range": (ctypes.c_int, [c_void, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                          ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_int64, ctypes.c_int64, c_void]),
    "dc_decoder_params_updated": (ctypes.c_int, [c_void, c_void]),
    "dc_decoder_grad_buffer": (ctypes.c_int, [c_void, ctypes.POINTER(c_void), ctypes.POINTER(ctypes.c_int64)]),
    "dc_decoder_param_buffer": (ctypes.c_int, [c_void, ctypes.POINTER(c_void), ctypes.POINTER(ctypes.c_int64)]),
    "dc_decoder_grad_bucket": (ctypes.c_int, [c_void, ctypes.c_int, ctypes.POINTER(ctypes.c_int64),
                                              ctypes.POINTER(ctypes.c_int64)]),
    "dc_decoder_wait_grad_bucket": (ctypes.c_int, [c_void, ctypes.c_int, c_void]),
    "dc_decoder_weight_offset": (ctypes.c_int64, [c_void, ctypes.c_int]),
    "dc_decoder_get_grad": (ctypes.c_int, [c_void, ctypes.c_char_p, c_void, ctypes.c_int64]),
}


def load():
    """Load libdcap.so once; raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libdcap.so not found at %s -- build it with `python -c 'import __graft_entry__ as "
                "g; g.build()'` (there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().dc_last_error()
        raise DcapError(rc, msg.decode("utf-8", "replace") if msg else "")


def device_info():
    lib = load()
    sms, cc = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.dc_device_info(ctypes.byref(sms), ctypes.byref(cc)))
    return sms.value, cc.value
