"""GPU numerics of the GEMM primitives against plain PyTorch references of the same op:
fp32 FFMA GEMM vs torch fp64 matmul; bf16 tcgen05 GEMM vs torch matmul on the SAME bf16 inputs
with fp32 accumulation (tolerance: fp32 re-association only, 1e-3 relative of the row scale)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["default", "pair", "single"])
def kernel_family(request, monkeypatch):
    """Every tcgen05 test runs three times: with the dispatcher's own choice (launches of fewer than ~sms/8 pair tiles
    take the single-CTA kernels), with the CTA-pair kernels forced wherever N >= 256 (DCAP_2CTA=1) and with them off
    (DCAP_2CTA=0) -- csrc/gemm_tc.cu reads the variable per call."""
    if request.param == "pair":
        monkeypatch.setenv("DCAP_2CTA", "1")
    elif request.param == "single":
        monkeypatch.setenv("DCAP_2CTA", "0")
    else:
        monkeypatch.delenv("DCAP_2CTA", raising=False)
    return request.param


def _rand(shape, seed, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g).to(dtype)


@pytest.mark.parametrize("M,N,K,ta,tb", [(100, 2048, 812, False, False), (257, 130, 300, False, False),
                                         (64, 96, 40, True, False), (130, 70, 129, False, True),
                                         (33, 65, 17, True, True), (1, 10000, 1024, False, False)])
def test_sgemm_all_layouts(M, N, K, ta, tb):
    from image_captioning_b200 import gemm
    a = _rand((K, M) if ta else (M, K), 1)
    b = _rand((N, K) if tb else (K, N), 2)
    bias = _rand((N,), 3)
    add = _rand((M, N), 4)
    out = gemm.gemm_f32(a, b, ta, tb, bias=bias, addend=add, relu=True)
    A = (a.t() if ta else a).double()
    B = (b.t() if tb else b).double()
    want = torch.relu(A @ B + bias.double() + add.double()).float()
    torch.testing.assert_close(out, want, rtol=1e-4, atol=1e-4 * K ** 0.5)
    acc = gemm.gemm_f32(a, b, ta, tb, out=out.clone(), accumulate=True)
    torch.testing.assert_close(acc, (want.double() + A @ B).float(), rtol=1e-4, atol=2e-4 * K ** 0.5)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 128), (8000, 2048, 832), (100, 1024, 512),
                                   (777, 10000, 1024), (4096, 1024, 12544), (129, 72, 320), (5, 2048, 1024)])
def test_tcgen05_gemm_store(M, N, K):
    from image_captioning_b200 import gemm
    a = _rand((M, K), 11, torch.bfloat16)
    bt = _rand((N, K), 12, torch.bfloat16)
    bias = _rand((N,), 13)
    out = gemm.gemm_bf16(a, bt, bias=bias)
    want = a.float() @ bt.float().t() + bias
    scale = float(K) ** 0.5
    torch.testing.assert_close(out, want, rtol=1e-3, atol=1e-3 * scale)
    # relu + addend + bf16 output
    add = _rand((M, N), 14)
    out2 = gemm.gemm_bf16(a, bt, bias=bias, addend=add, relu=True, out_dtype=torch.bfloat16)
    want2 = torch.relu(want + add).to(torch.bfloat16)
    torch.testing.assert_close(out2.float(), want2.float(), rtol=2e-2, atol=2e-2 * scale)


def test_tcgen05_gemm_strided_operands():
    """A and the outputs are column slices of wider buffers (how the decoder's [emb|h] operand
    buffers are laid out)."""
    from image_captioning_b200 import gemm
    M, N, K = 300, 512, 512
    wide = _rand((M, 1024), 21, torch.bfloat16)
    a = wide[:, 512:]
    bt = _rand((N, K), 22, torch.bfloat16)
    out = gemm.gemm_bf16(a, bt)
    torch.testing.assert_close(out, a.float() @ bt.float().t(), rtol=1e-3, atol=3e-2)


@pytest.mark.parametrize("M,N,K", [(1000, 10000, 1024), (130, 300, 128), (64, 128, 64)])
def test_tcgen05_gemm_argmax_epilogue(M, N, K):
    from image_captioning_b200 import gemm
    a = _rand((M, K), 31, torch.bfloat16)
    bt = _rand((N, K), 32, torch.bfloat16)
    bias = _rand((N,), 33)
    tok, prob = gemm.gemm_bf16_argmax(a, bt, bias, want_prob=True)
    logits = gemm.gemm_bf16(a, bt, bias=bias)                   # same kernel, store epilogue
    want = logits.argmax(-1).to(torch.int32)
    # identical accumulation order in both epilogues -> identical arg-max
    assert torch.equal(tok, want)
    torch.testing.assert_close(prob, torch.softmax(logits, -1).max(-1).values, rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("M,N,K,k", [(1000, 10000, 1024, 3), (130, 300, 128, 5), (64, 128, 64, 8), (200, 520, 64, 1)])
def test_tcgen05_gemm_topk_epilogue(M, N, K, k):
    from image_captioning_b200 import gemm
    a = _rand((M, K), 34, torch.bfloat16)
    bt = _rand((N, K), 35, torch.bfloat16)
    bias = _rand((N,), 36)
    idx, prob = gemm.gemm_bf16_topk(a, bt, bias, k)
    logits = gemm.gemm_bf16(a, bt, bias=bias)                   # same kernel, store epilogue
    p = torch.softmax(logits, -1)
    want_p, want_i = torch.topk(logits, k, dim=-1)              # random logits: no ties
    assert torch.equal(idx, want_i.flip(-1).to(torch.int32))    # ascending order, best last
    torch.testing.assert_close(prob, torch.gather(p, 1, want_i.flip(-1)), rtol=2e-3, atol=1e-7)


def test_topk_larger_index_wins_ties():
    """np.argsort(p, stable)[-k:]: among equal probabilities the larger index is the better candidate."""
    from image_captioning_b200 import gemm
    M, N, K = 64, 600, 64
    a = torch.zeros((M, K), device="cuda", dtype=torch.bfloat16)
    bt = _rand((N, K), 41, torch.bfloat16)
    bias = torch.zeros((N,), device="cuda")
    bias[[17, 300, 301, 555]] = 1.0                              # four equal maxima across tiles / chunks
    idx, prob = gemm.gemm_bf16_topk(a, bt, bias, 3)
    assert bool((idx == torch.tensor([300, 301, 555], device="cuda", dtype=torch.int32)).all())
    idx5, _ = gemm.gemm_bf16_topk(a, bt, bias, 5)
    assert bool((idx5[:, 1:] == torch.tensor([17, 300, 301, 555], device="cuda", dtype=torch.int32)).all())
    assert bool((idx5[:, 0] == 599).all())                       # all remaining logits tie at 0: largest index


def test_argmax_first_index_on_ties():
    from image_captioning_b200 import gemm
    M, N, K = 64, 600, 64
    a = torch.zeros((M, K), device="cuda", dtype=torch.bfloat16)
    bt = _rand((N, K), 41, torch.bfloat16)
    bias = torch.zeros((N,), device="cuda")
    bias[[17, 300, 555]] = 1.0                                   # three equal maxima in different tiles
    tok = gemm.gemm_bf16_argmax(a, bt, bias)
    assert bool((tok == 17).all())


@pytest.mark.parametrize("M,U,K", [(300, 64, 128), (8000, 512, 832), (129, 512, 1024)])
def test_tcgen05_gemm_lstm_cell_epilogue(M, U, K):
    """Fused gates GEMM + cell against the same math in torch on the same bf16 operands."""
    from image_captioning_b200 import gemm
    a = _rand((M, K), 51, torch.bfloat16) * 0.5
    w = _rand((K, 4 * U), 52) * (1.0 / K ** 0.5)                  # Keras layout [K, 4U], blocks i|f|c|o
    addend = _rand((M, 4 * U), 53) * 0.5                          # Keras column order
    bias = _rand((4 * U,), 54) * 0.1
    c0 = _rand((M, U), 55)
    h_prev = _rand((M, U), 56, torch.bfloat16)
    tok = (torch.arange(M, device="cuda") % 5 != 0).to(torch.int32) * 7       # every 5th row masked
    # interleave: row 4u+g of Bt = column g*U+u of w
    perm = (torch.arange(4, device="cuda")[None, :] * U + torch.arange(U, device="cuda")[:, None]).reshape(-1)
    bt = w.t()[perm].contiguous().to(torch.bfloat16)
    add_i = addend[:, perm].contiguous()
    bias_i = bias[perm].contiguous()
    c = c0.clone()
    h_out = torch.zeros((M, 2 * U), device="cuda", dtype=torch.bfloat16)
    h_out2 = torch.zeros((M, U), device="cuda", dtype=torch.bfloat16)
    gemm.gemm_bf16_lstm_cell(a, bt, U, c, h_prev, h_out[:, U:], addend=add_i, bias=bias_i, tok=tok, h_out2=h_out2)
    z = a.float() @ w.to(torch.bfloat16).float() + addend + bias
    hs = lambda x: torch.clamp(0.2 * x + 0.5, 0, 1)
    i, f, g, o = hs(z[:, :U]), hs(z[:, U:2 * U]), torch.tanh(z[:, 2 * U:3 * U]), hs(z[:, 3 * U:])
    c_want = f * c0 + i * g
    h_want = o * torch.tanh(c_want)
    m = (tok != 0)[:, None]
    c_want = torch.where(m, c_want, c0)
    h_want = torch.where(m, h_want, h_prev.float())
    torch.testing.assert_close(c, c_want, rtol=2e-3, atol=2e-3)
    torch.testing.assert_close(h_out[:, U:].float(), h_want.to(torch.bfloat16).float(), rtol=2e-2, atol=1e-2)
    assert torch.equal(h_out[:, U:], h_out2)
    assert bool((h_out[:, :U] == 0).all())


@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("M,N,K", [(304, 2048, 4096), (512, 1024, 1000), (832, 256, 8192 + 64)])
def test_tcgen05_gemm_mn_major_operands(M, N, K, a_mn, b_mn):
    """MN-major UMMA operands: X^T * dY straight from the [R, in] / [R, out] activations."""
    from image_captioning_b200 import gemm
    a = _rand((M, K), 61).to(torch.bfloat16)
    b = _rand((N, K), 62).to(torch.bfloat16)
    want = a.float() @ b.float().t()
    a_op = a.t().contiguous() if a_mn else a
    b_op = b.t().contiguous() if b_mn else b
    got = gemm.gemm_bf16_ex(a_op, b_op, M, N, K, a_mn=a_mn, b_mn=b_mn)
    torch.testing.assert_close(got, want, rtol=2e-3, atol=2e-3 * K ** 0.5)


@pytest.mark.parametrize("split", [0, 3, 7])
def test_tcgen05_gemm_split_k_atomic_and_deinterleave(split):
    from image_captioning_b200 import gemm
    M, U, K = 304, 64, 9000
    N = 4 * U
    a = _rand((K, M), 63).to(torch.bfloat16)           # MN-major A: [K, M]
    b = _rand((K, N), 64).to(torch.bfloat16)           # MN-major B: [K, N], gate-interleaved columns
    want_i = a.float().t() @ b.float()                 # [M, N] interleaved
    perm = (torch.arange(4, device="cuda")[None, :] * U + torch.arange(U, device="cuda")[:, None]).reshape(-1)
    want = torch.empty_like(want_i)
    want[:, perm] = want_i                             # column 4u+g -> g*U+u
    base = _rand((M, N), 65)
    out = base.clone()
    gemm.gemm_bf16_ex(a, b, M, N, K, a_mn=True, b_mn=True, deint_units=U, atomic=True, split_k=split, out=out)
    torch.testing.assert_close(out - base, want, rtol=2e-3, atol=0.3)


def test_tcgen05_gemm_relu_mask_and_addend_mod():
    from image_captioning_b200 import gemm
    M, N, K, B = 700, 1024, 520, 100
    a = _rand((M, K), 66).to(torch.bfloat16)
    b = _rand((N, K), 67).to(torch.bfloat16)
    addend = _rand((B, N), 68)
    act = torch.relu(_rand((M, N), 69)).to(torch.bfloat16)
    want = a.float() @ b.float().t() + addend[torch.arange(M, device="cuda") % B]
    got = gemm.gemm_bf16_ex(a, b, M, N, K, addend=addend, addend_mod=B)
    torch.testing.assert_close(got, want, rtol=2e-3, atol=0.1)
    got = gemm.gemm_bf16_ex(a, b, M, N, K, mask_src=act, out_dtype=torch.bfloat16)
    want_m = torch.where(act.float() > 0, a.float() @ b.float().t(), torch.zeros((), device="cuda"))
    torch.testing.assert_close(got.float(), want_m.to(torch.bfloat16).float(), rtol=2e-2, atol=0.2)
