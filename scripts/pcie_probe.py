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
