import torch, time
x = torch.empty(201326592 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(x, non_blocking=True)), ("D2H", lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(name, f"{201.3 / dt / 1e3:.1f} GB/s", f"{dt*1e3:.2f} ms")
