import os, time, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
a = torch.randn(2048, 1024); b = torch.randn(1024, 4096)
for n in (4, 8, 16, 32, 64, 128):
    torch.set_num_threads(n)
    a @ b
    t0 = time.perf_counter()
    for _ in range(5): a @ b
    dt = (time.perf_counter() - t0) / 5
    print(f"threads {n}: {2*2048*1024*4096/dt/1e9:.1f} GFLOP/s")
