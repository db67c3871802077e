import torch, time
x = torch.empty(713159680 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for n_chunks in (1, 8, 32):
    parts = x.chunk(n_chunks); dparts = d.chunk(n_chunks)
    torch.cuda.synchronize()
    for _ in range(2):
        for a, b in zip(parts, dparts): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for a, b in zip(parts, dparts): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("chunks %2d: %.2f ms  %.1f GB/s" % (n_chunks, dt * 1e3, x.numel() * 4 / dt / 1e9))
