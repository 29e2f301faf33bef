import torch, time
x = torch.empty(512*513*938, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(x, non_blocking=True)), ("d2h", lambda: x.copy_(d, non_blocking=True))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
    print(name, x.numel()*4/dt/1e9, "GB/s", dt*1e3, "ms")
t=time.perf_counter(); y=torch.empty(512*239872, dtype=torch.float32, pin_memory=True); print("pinned alloc 491MB", (time.perf_counter()-t)*1e3,"ms")
del y
t=time.perf_counter(); y=torch.empty(512*239872, dtype=torch.float32, pin_memory=True); print("pinned alloc again", (time.perf_counter()-t)*1e3,"ms")
