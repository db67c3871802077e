"""A/B timing of the decode-step GEMMs using only the API common to old and new trees."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from image_captioning_b200 import gemm


def timeit(run, n=20):
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
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best * 1e3


M, U = 8000, 512
res = {}
for K in (832, 1024):
    a = torch.randn((M, K), device="cuda").bfloat16() * 0.1
    bt = torch.randn((4 * U, K), device="cuda").bfloat16() * 0.05
    addend = torch.randn((M, 4 * U), device="cuda")
    c = torch.zeros((M, U), device="cuda")
    h_prev = torch.zeros((M, U), device="cuda", dtype=torch.bfloat16)
    h_out = torch.zeros((M, U), device="cuda", dtype=torch.bfloat16)
    tok = torch.ones((M,), device="cuda", dtype=torch.int32)
    res["cell K=%d" % K] = timeit(lambda: gemm.gemm_bf16_lstm_cell(a, bt, U, c, h_prev, h_out, addend=addend, tok=tok))
a = torch.randn((M, 512), device="cuda").bfloat16()
bt = torch.randn((1024, 512), device="cuda").bfloat16()
add = torch.randn((M, 1024), device="cuda")
res["dense1 bf16+addend+relu"] = timeit(lambda: gemm.gemm_bf16(a, bt, addend=add, relu=True, out_dtype=torch.bfloat16))
a = torch.randn((M, 1024), device="cuda").bfloat16()
bt = torch.randn((10000, 1024), device="cuda").bfloat16()
bias = torch.randn((10000,), device="cuda")
res["vocab argmax"] = timeit(lambda: gemm.gemm_bf16_argmax(a, bt, bias))
res["vocab f32"] = timeit(lambda: gemm.gemm_bf16(a, bt, bias=bias))
a = torch.randn((M, 12544), device="cuda").bfloat16()
bt = torch.randn((1024, 12544), device="cuda").bfloat16()
res["head1 bf16"] = timeit(lambda: gemm.gemm_bf16(a, bt, bias=bias[:1024].contiguous(), relu=True, out_dtype=torch.bfloat16))
print(" | ".join("%s %.1f us" % kv for kv in res.items()))
