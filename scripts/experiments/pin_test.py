import time, torch
dev = torch.device("cuda", 0)
x = torch.randn(32, 100, 512, 257, 2, device=dev)  # 3.37 GB
torch.cuda.synchronize()
t = time.perf_counter(); y = x.cpu(); print("pageable .cpu(): %.2f s" % (time.perf_counter() - t))
t = time.perf_counter(); p = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); print("pinned alloc: %.2f s" % (time.perf_counter() - t))
t = time.perf_counter(); p.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("pinned copy: %.2f s" % (time.perf_counter() - t))
del p
t = time.perf_counter(); p = torch.empty(x.shape, dtype=x.dtype, pin_memory=True); print("pinned alloc again: %.2f s" % (time.perf_counter() - t))
t = time.perf_counter(); p.copy_(x, non_blocking=True); torch.cuda.synchronize(); print("pinned copy again: %.2f s" % (time.perf_counter() - t))
# staged through a 256 MB pinned ring into pageable memory
out = torch.empty(x.shape, dtype=x.dtype)
flat_d, flat_h = x.view(-1), out.view(-1)
CH = 64 * 1024 * 1024
ring = [torch.empty(CH, dtype=x.dtype, pin_memory=True) for _ in range(2)]
ev = [torch.cuda.Event() for _ in range(2)]
t = time.perf_counter()
n = flat_d.numel(); nchunks = (n + CH - 1) // CH
for i in range(nchunks + 1):
    if i < nchunks:
        a, b = i * CH, min(n, (i + 1) * CH)
        ring[i % 2][: b - a].copy_(flat_d[a:b], non_blocking=True); ev[i % 2].record()
    if i > 0:
        j = i - 1; a, b = j * CH, min(n, (j + 1) * CH)
        ev[j % 2].synchronize(); flat_h[a:b].copy_(ring[j % 2][: b - a])
print("staged ring -> pageable: %.2f s" % (time.perf_counter() - t), bool(torch.equal(out, y)))
