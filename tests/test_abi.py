"""The C-ABI library loads and exports every symbol include/dcap.h declares (no compute calls,
no GPU needed), and the ctypes signature table covers the same set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dcap.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import image_captioning_b200 as pkg
    names = _declared()
    assert len(names) >= 6
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libdcap.so does not export %s" % n


def test_ctypes_table_matches_header():
    import image_captioning_b200 as pkg
    assert sorted(pkg._lib.SIGNATURES) == _declared()


def test_last_error_is_thread_local_string():
    import image_captioning_b200 as pkg
    lib = pkg._lib.load()
    msg = lib.dc_last_error()
    assert isinstance(msg, bytes)


def test_argument_validation_without_gpu():
    """Shape validation happens before any CUDA call, so it can be exercised on CPU."""
    import image_captioning_b200 as pkg
    lib = pkg._lib.load()
    ptrs = (ctypes.c_void_p * 4)(16, 16, 16, 16)
    hs = (ctypes.c_int * 4)(8, 4, 2, 1)
    rc = lib.dc_pyramid_roi_align_f32(ctypes.c_void_p(16), ptrs, hs, hs, 1, 100001, 256, 7, 7,
                                      1024, 1024, ctypes.c_void_p(16), None, None)
    assert rc == -1 and b"100000" in lib.dc_last_error()
    rc = lib.dc_pyramid_roi_align_f32(ctypes.c_void_p(16), ptrs, hs, hs, 1, 10, 255, 7, 7,
                                      1024, 1024, ctypes.c_void_p(16), None, None)
    assert rc == -1 and b"multiple of 4" in lib.dc_last_error()
    rc = lib.dc_pyramid_roi_align_f32(ctypes.c_void_p(16), ptrs, hs, hs, 1, 10, 256, 0, 7,
                                      1024, 1024, ctypes.c_void_p(16), None, None)
    assert rc == -1
    # empty call is a no-op success
    rc = lib.dc_pyramid_roi_align_f32(None, ptrs, hs, hs, 0, 10, 256, 7, 7, 1024, 1024, None, None,
                                      None)
    assert rc == 0
