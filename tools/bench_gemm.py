"""Times the tcgen05 GEMM on the decoder's shapes (tuning aid; not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_captioning_b200 import gemm


def timeit(run, n=20):
    """GPU time per call: the n calls are captured in one CUDA graph so that host launch overhead
    (ctypes, tensor-map lookup) does not count."""
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                run()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best


which = sys.argv[1:] or ["plain", "cell"]
if "plain" in which:
    shapes = [("gates1", 8000, 2048, 832), ("dense1", 8000, 1024, 512), ("vocab", 8000, 10000, 1024),
              ("vocabT", 65536, 10000, 1024), ("head1", 8000, 1024, 12544), ("big", 8192, 8192, 8192)]
    for name, M, N, K in shapes:
        a = torch.randn((M, K), device="cuda").bfloat16()
        bt = torch.randn((N, K), device="cuda").bfloat16()
        bias = torch.randn((N,), device="cuda")
        out32 = torch.empty((M, N), device="cuda")
        out16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
        for mode in ("f32", "bf16", "argmax", "torch"):
            if mode == "argmax" and N < 10000:
                continue
            def run():
                if mode == "f32":
                    return gemm.gemm_bf16_ex(a, bt, M, N, K, bias=bias, out=out32)
                if mode == "bf16":
                    return gemm.gemm_bf16_ex(a, bt, M, N, K, bias=bias, out=out16)
                if mode == "argmax":
                    return gemm.gemm_bf16_argmax(a, bt, bias)
                return torch.nn.functional.linear(a, bt, out=None) if False else torch.matmul(a, bt.t(), out=out16)
            ms = timeit(run)
            print("%-7s %-6s M=%d N=%d K=%d  %.3f ms  %.0f TFLOP/s" % (name, mode, M, N, K, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
if "cell" in which:
    for M in (8000, 4096, 24576):
        for K in (832, 1024):
            U = 512
            a = torch.randn((M, K), device="cuda").bfloat16() * 0.1
            bt = torch.randn((4 * U, K), device="cuda").bfloat16() * 0.05
            addend = torch.randn((M, 4 * U), device="cuda")
            bias = torch.randn((4 * U,), device="cuda")
            c = torch.zeros((M, U), device="cuda")
            h_prev = torch.zeros((M, U), device="cuda", dtype=torch.bfloat16)
            h_out = torch.zeros((M, U), device="cuda", dtype=torch.bfloat16)
            tok = torch.ones((M,), device="cuda", dtype=torch.int32)
            out16 = torch.empty((M, 4 * U), device="cuda", dtype=torch.bfloat16)
            for mode in ("full", "no_addend", "store_bf16", "torch"):
                def run():
                    if mode == "full":
                        return gemm.gemm_bf16_lstm_cell(a, bt, U, c, h_prev, h_out, addend=addend, bias=None, tok=tok)
                    if mode == "no_addend":
                        return gemm.gemm_bf16_lstm_cell(a, bt, U, c, h_prev, h_out, addend=None, bias=bias, tok=tok)
                    if mode == "store_bf16":
                        return gemm.gemm_bf16_ex(a, bt, M, 4 * U, K, bias=bias, out=out16)
                    return torch.matmul(a, bt.t(), out=out16)
                ms = timeit(run)
                print("cell M=%d K=%d %-10s %.1f us  %.0f TFLOP/s" % (M, K, mode, ms * 1e3, 2.0 * M * 4 * U * K / ms / 1e9), flush=True)
