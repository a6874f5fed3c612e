import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
x = torch.randn(2, 256, 65, 65, requires_grad=True); w = torch.randn(256, 256, 3, 3, requires_grad=True)
for n in (8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    for _ in range(2):
        y = torch.nn.functional.conv2d(x, w, padding=2, dilation=2); y.sum().backward()
    t0 = time.perf_counter()
    for _ in range(5):
        y = torch.nn.functional.conv2d(x, w, padding=2, dilation=2); y.sum().backward()
    print("threads", n, "ms per conv fwd+bwd", (time.perf_counter() - t0) / 5 * 1e3, flush=True)
