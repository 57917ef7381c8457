"""PCIe probe for the host-buffer path: H2D alone, D2H alone, both concurrently (two streams), at the bench's copy size."""
import time
import torch

n = 4096 * 1024
dev = torch.device("cuda", 0)
h_in = [torch.empty(n, dtype=torch.bfloat16).pin_memory() for _ in range(3)]
h_out = [torch.empty(n, dtype=torch.bfloat16).pin_memory() for _ in range(3)]
d_in = [torch.empty(n, dtype=torch.bfloat16, device=dev) for _ in range(3)]
d_out = [torch.empty(n, dtype=torch.bfloat16, device=dev) for _ in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
iters = 40


def run(do_h2d, do_d2h):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        if do_h2d:
            with torch.cuda.stream(s1):
                d_in[i % 3].copy_(h_in[i % 3], non_blocking=True)
        if do_d2h:
            with torch.cuda.stream(s2):
                h_out[i % 3].copy_(d_out[i % 3], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
mb = n * 2 / 1e6
print(f"copy size {mb:.1f} MB: H2D alone {a:.3f} ms ({mb / a:.1f} GB/s), D2H alone {b:.3f} ms ({mb / b:.1f} GB/s), "
      f"both concurrently {c:.3f} ms per pair ({2 * mb / c:.1f} GB/s aggregate)")


def run_split(nchunks):
    streams = [torch.cuda.Stream() for _ in range(nchunks)]
    step = n // nchunks
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(iters):
        for c, s in enumerate(streams):
            with torch.cuda.stream(s):
                d_in[i % 3][c * step:(c + 1) * step].copy_(h_in[i % 3][c * step:(c + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


for nch in (2, 4):
    run_split(nch)
    t = run_split(nch)
    print(f"H2D split into {nch} concurrent chunks: {t:.3f} ms ({mb / t:.1f} GB/s)")

# write-combined pinned memory through the CUDA runtime (torch has no flag for it)
import ctypes

rt = ctypes.CDLL("libcudart.so.12")
ptr = ctypes.c_void_p()
rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n * 2), ctypes.c_uint(4))  # cudaHostAllocWriteCombined
if rc == 0:
    ctypes.memset(ptr, 1, n * 2)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        for i in range(iters):
            rt.cudaMemcpyAsync(ctypes.c_void_p(d_in[i % 3].data_ptr()), ptr, ctypes.c_size_t(n * 2), ctypes.c_int(1),
                               ctypes.c_void_p(st.cuda_stream))
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) / iters * 1e3
    print(f"H2D from write-combined pinned memory: {t:.3f} ms ({mb / t:.1f} GB/s)")
    rt.cudaFreeHost(ptr)
import os
print("cpus allowed:", len(os.sched_getaffinity(0)), "numa nodes:", len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')]) if os.path.isdir('/sys/devices/system/node') else '?')
