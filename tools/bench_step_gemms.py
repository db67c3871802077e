"""Times the four GEMMs of one greedy decode step in their real epilogue configurations, each alone (20 calls in one
CUDA graph, best of 5): the A/B harness for epilogue work.  usage: python tools/bench_step_gemms.py [M]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_captioning_b200 import gemm
from tools.bench_gemm_common import timeit

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
U, E, V, D = 512, 320, 10000, 1024
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
rb = lambda *s: (torch.randn(s, device=dev, generator=g) * 0.05).bfloat16()
rf = lambda *s: torch.randn(s, device=dev, generator=g)
tok = torch.ones((M,), device=dev, dtype=torch.int32)
res = []
# LSTM1: [emb | h1] x [W1e ; U1] + hoisted addend, fused cell, h to two destinations
x1, w1, add1 = rb(M, E + U), rb(4 * U, E + U), rf(M, 4 * U)
c1, hp1, ho1, ho1b = torch.zeros((M, U), device=dev), torch.zeros((M, U), device=dev, dtype=torch.bfloat16), torch.zeros((M, U), device=dev, dtype=torch.bfloat16), torch.zeros((M, U), device=dev, dtype=torch.bfloat16)
res.append(("LSTM1 gates+cell (K=832, addend)", 2.0 * M * 4 * U * (E + U), timeit(lambda: gemm.gemm_bf16_lstm_cell(x1, w1, U, c1, hp1, ho1, addend=add1, tok=tok, h_out2=ho1b))))
# LSTM2: [h1 | h2] x [W2 ; U2] + bias, fused cell
x2, w2, b2 = rb(M, 2 * U), rb(4 * U, 2 * U), rf(4 * U)
c2, hp2, ho2 = torch.zeros((M, U), device=dev), torch.zeros((M, U), device=dev, dtype=torch.bfloat16), torch.zeros((M, U), device=dev, dtype=torch.bfloat16)
res.append(("LSTM2 gates+cell (K=1024, bias)", 2.0 * M * 4 * U * 2 * U, timeit(lambda: gemm.gemm_bf16_lstm_cell(x2, w2, U, c2, hp2, ho2, bias=b2, tok=tok))))
# dense1: h2 x Wd1h + hoisted addend, relu, bf16 out
h2, wd1, addd = rb(M, U), rb(D, U), rf(M, D)
d = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
res.append(("dense1 (K=512, addend, relu, bf16)", 2.0 * M * D * U, timeit(lambda: gemm.gemm_bf16_ex(h2, wd1, M, D, U, addend=addd, relu=True, out=d))))
res.append(("dense1 without addend", 2.0 * M * D * U, timeit(lambda: gemm.gemm_bf16_ex(h2, wd1, M, D, U, relu=True, out=d))))
# vocabulary projection + arg-max
wd2, bv = rb(V, D), rf(V)
res.append(("vocabulary GEMM + arg-max (K=1024)", 2.0 * M * V * D, timeit(lambda: gemm.gemm_bf16_argmax(d, wd2, bv))))
tot = 0.0
for name, fl, ms in res:
    print("%-38s %7.1f us  %6.0f TFLOP/s" % (name, ms * 1e3, fl / ms / 1e9), flush=True)
    if "without" not in name:
        tot += ms
print("sum of the four: %.1f us per step" % (tot * 1e3))
