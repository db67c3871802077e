"""ctypes wrapper of the C oracle (oracle/c/liboracle.so).  Test / bench infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "c", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = ctypes.CDLL(SO)
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _args(boxes, fms):
    boxes = np.ascontiguousarray(boxes, np.float32)
    fms = [np.ascontiguousarray(f, np.float32) for f in fms]
    ptrs = (ctypes.c_void_p * 4)(*[f.ctypes.data for f in fms])
    hs = (ctypes.c_int * 4)(*[f.shape[1] for f in fms])
    ws = (ctypes.c_int * 4)(*[f.shape[2] for f in fms])
    return boxes, fms, ptrs, hs, ws


def fpn_levels(boxes, image_shape):
    lib = load()
    b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    out = np.empty(b.shape[0], np.int32)
    lib.oracle_fpn_levels_f32(ctypes.c_void_p(b.ctypes.data), ctypes.c_int64(b.shape[0]),
                              int(image_shape[0]), int(image_shape[1]), ctypes.c_void_p(out.ctypes.data))
    return out.reshape(np.asarray(boxes).shape[:-1])


def pyramid_roi_align(boxes, fms, pool_shape, image_shape, literal=False, out=None, scratch=None):
    """Returns (out [B*N,ph,pw,C], levels [B,N])."""
    lib = load()
    boxes, fms, ptrs, hs, ws = _args(boxes, fms)
    B, N = boxes.shape[:2]
    C = fms[0].shape[-1]
    ph, pw = pool_shape
    if out is None:
        out = np.empty((B * N, ph, pw, C), np.float32)
    lv = np.empty((B, N), np.int32)
    if literal:
        if scratch is None:
            scratch = np.empty_like(out)
        rc = lib.oracle_pyramid_roi_align_literal_f32(
            ctypes.c_void_p(boxes.ctypes.data), ptrs, hs, ws, B, N, C, ph, pw, int(image_shape[0]),
            int(image_shape[1]), ctypes.c_void_p(scratch.ctypes.data), ctypes.c_void_p(out.ctypes.data),
            ctypes.c_void_p(lv.ctypes.data))
    else:
        rc = lib.oracle_pyramid_roi_align_f32(
            ctypes.c_void_p(boxes.ctypes.data), ptrs, hs, ws, B, N, C, ph, pw, int(image_shape[0]),
            int(image_shape[1]), ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(lv.ctypes.data))
    if rc != 0:
        raise ValueError("oracle rejected the call (rc=%d)" % rc)
    return out, lv


def num_threads():
    return load().oracle_num_threads()
