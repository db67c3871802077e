"""Times the fused arg-max / top-k vocabulary GEMMs at the beam-search chunk size (tuning aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_captioning_b200 import gemm
from tools.bench_gemm_common import timeit

M, N, K = 49152, 10000, 1024
a = torch.randn((M, K), device="cuda").bfloat16()
bt = torch.randn((N, K), device="cuda").bfloat16()
bias = torch.randn((N,), device="cuda")
for name, fn in (("argmax", lambda: gemm.gemm_bf16_argmax(a, bt, bias)),
                 ("argmax+prob", lambda: gemm.gemm_bf16_argmax(a, bt, bias, want_prob=True)),
                 ("topk1", lambda: gemm.gemm_bf16_topk(a, bt, bias, 1)),
                 ("topk3", lambda: gemm.gemm_bf16_topk(a, bt, bias, 3)),
                 ("topk8", lambda: gemm.gemm_bf16_topk(a, bt, bias, 8))):
    ms = timeit(fn, n=5)
    print("%-12s %.3f ms  %.0f TFLOP/s" % (name, ms, 2.0 * M * N * K / ms / 1e9), flush=True)
