"""Times the tcgen05 GEMM on the decoder's shapes (tuning aid; not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_captioning_b200 import gemm

shapes = [("gates1", 8000, 2048, 832), ("gates2", 8000, 2048, 1024), ("dense1", 8000, 1024, 512),
          ("vocab", 8000, 10000, 1024), ("head1", 8000, 1024, 12544), ("head2", 8000, 1024, 1024),
          ("big", 8192, 8192, 8192)]
for name, M, N, K in shapes:
    a = torch.randn((M, K), device="cuda").bfloat16()
    bt = torch.randn((N, K), device="cuda").bfloat16()
    bias = torch.randn((N,), device="cuda")
    for mode in ("f32", "bf16", "argmax", "torch"):
        def run():
            if mode == "f32":
                return gemm.gemm_bf16(a, bt, bias=bias)
            if mode == "bf16":
                return gemm.gemm_bf16(a, bt, bias=bias, out_dtype=torch.bfloat16)
            if mode == "argmax":
                return gemm.gemm_bf16_argmax(a, bt, bias)
            return torch.nn.functional.linear(a, bt)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 20
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print("%-7s %-6s M=%d N=%d K=%d  %.3f ms  %.0f TFLOP/s" % (name, mode, M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
