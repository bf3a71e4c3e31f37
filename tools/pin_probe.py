import time, torch
shape = (1, 4320, 7680, 3)
dev = torch.rand(shape, device="cuda")
keep = None
for i in range(8):
    t0 = time.perf_counter()
    x = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    t1 = time.perf_counter()
    x.copy_(dev, non_blocking=True); torch.cuda.synchronize()
    t2 = time.perf_counter()
    keep = x
    print(i, "alloc ms %.2f  d2h ms %.2f  ptr %x" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, x.data_ptr()))
s = torch.cuda.Stream()
for i in range(6):
    t0 = time.perf_counter()
    x = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    t1 = time.perf_counter()
    with torch.cuda.stream(s):
        x[:, :2000].copy_(dev[:, :2000], non_blocking=True)
        x[:, 2000:].copy_(dev[:, 2000:], non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    keep = x
    print("side-stream", i, "alloc ms %.2f  d2h ms %.2f  ptr %x" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, x.data_ptr()))
