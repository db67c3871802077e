"""Thin torch-tensor wrappers over the exported GEMM primitives (dc_gemm_*); used by tests and
tuning scripts.  The decoder calls the same kernels from C++."""
import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _s(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gemm_f32(a, b, trans_a=False, trans_b=False, bias=None, addend=None, relu=False, out=None, accumulate=False):
    lib = _lib.load()
    M = a.shape[1] if trans_a else a.shape[0]
    K = a.shape[0] if trans_a else a.shape[1]
    N = b.shape[0] if trans_b else b.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_f32(_p(a), a.stride(0), int(trans_a), _p(b), b.stride(0), int(trans_b), M, N, K,
                                   _p(bias), _p(addend), addend.stride(0) if addend is not None else 0,
                                   int(relu), int(accumulate), _p(out), out.stride(0), _s(a.device)))
    return out


def gemm_bf16(a, bt, bias=None, addend=None, relu=False, out_dtype=torch.float32):
    """a [M,K] bf16, bt [N,K] bf16 (K contiguous) -> [M,N]."""
    lib = _lib.load()
    M, K = a.shape
    N = bt.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    f32 = out if out_dtype == torch.float32 else None
    b16 = out if out_dtype == torch.bfloat16 else None
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_bf16(_p(a), a.stride(0), _p(bt), bt.stride(0), M, N, K, _p(bias), _p(addend),
                                    addend.stride(0) if addend is not None else 0, int(relu),
                                    _p(f32), N if f32 is not None else 0, _p(b16), N if b16 is not None else 0,
                                    _s(a.device)))
    return out


def gemm_bf16_ex(a, b, M, N, K, a_mn=False, b_mn=False, bias=None, addend=None, addend_mod=0, relu=False,
                 mask_src=None, deint_units=0, atomic=False, split_k=1, out=None, out_dtype=torch.float32):
    """General tcgen05 GEMM (dc_gemm_bf16_ex).  K-major operand: [rows, K]; MN-major: [K, rows]."""
    lib = _lib.load()
    if out is None:
        out = (torch.zeros if atomic else torch.empty)((M, N), dtype=out_dtype, device=a.device)
    f32 = out if out.dtype == torch.float32 else None
    b16 = out if out.dtype == torch.bfloat16 else None
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_bf16_ex(
            _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), M, N, K, _p(bias), _p(addend),
            addend.stride(0) if addend is not None else 0, int(addend_mod), int(relu), _p(mask_src),
            mask_src.stride(0) if mask_src is not None else 0, int(deint_units), int(atomic), int(split_k),
            _p(f32), out.stride(0) if f32 is not None else 0, _p(b16), out.stride(0) if b16 is not None else 0,
            _s(a.device)))
    return out


def gemm_bf16_argmax(a, bt, bias, want_prob=False):
    lib = _lib.load()
    M, K = a.shape
    N = bt.shape[0]
    tok = torch.empty((M,), dtype=torch.int32, device=a.device)
    prob = torch.empty((M,), dtype=torch.float32, device=a.device) if want_prob else None
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_bf16_argmax(_p(a), a.stride(0), _p(bt), bt.stride(0), M, N, K, _p(bias), _p(tok),
                                           _p(prob), _s(a.device)))
    return (tok, prob) if want_prob else tok


def gemm_bf16_topk(a, bt, bias, k):
    """Fused vocabulary GEMM + softmax top-k (ascending): (idx [M,k] int32, prob [M,k] fp32)."""
    lib = _lib.load()
    M, K = a.shape
    N = bt.shape[0]
    idx = torch.empty((M, k), dtype=torch.int32, device=a.device)
    prob = torch.empty((M, k), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_bf16_topk(_p(a), a.stride(0), _p(bt), bt.stride(0), M, N, K, _p(bias), int(k), _p(idx),
                                         _p(prob), _s(a.device)))
    return idx, prob


def gemm_bf16_lstm_cell(a, bt_interleaved, units, c, h_prev, h_out, addend=None, bias=None, tok=None, h_out2=None):
    """Fused gates GEMM + Keras LSTM cell (gate-interleaved columns); c updated in place."""
    lib = _lib.load()
    M, K = a.shape
    with torch.cuda.device(a.device):
        _lib.check(lib.dc_gemm_bf16_lstm_cell(
            _p(a), a.stride(0), _p(bt_interleaved), bt_interleaved.stride(0), M, units, K, _p(addend),
            addend.stride(0) if addend is not None else 0, _p(bias), _p(tok), _p(c), _p(h_prev),
            h_prev.stride(0) if h_prev is not None else 0, _p(h_out), h_out.stride(0), _p(h_out2),
            h_out2.stride(0) if h_out2 is not None else 0, _s(a.device)))
    return h_out
