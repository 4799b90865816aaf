#!/usr/bin/env python
"""Read-only / write-only / copy bandwidth of this B200 with plain torch kernels (CUDA events, 1 GiB buffers, best of 5):
the write-only figure is the ceiling of the output-heavy GEMM epilogues."""
import torch

n = 1 << 29                      # 512 Mi bf16 = 1 GiB
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)


def best(fn, nbytes):
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return nbytes / min(ts[1:]) / 1e6


print("write-only  (fill)      %7.0f GB/s" % best(lambda: b.fill_(1.0), 2 * n))
print("write-only  (zero)      %7.0f GB/s" % best(lambda: b.zero_(), 2 * n))
print("read-only   (sum)       %7.0f GB/s" % best(lambda: a.float().sum() if False else torch.sum(a, dtype=torch.float32), 2 * n))
print("copy        (read+write)%7.0f GB/s" % best(lambda: b.copy_(a), 4 * n))
print("2 reads + 1 write (add) %7.0f GB/s" % best(lambda: torch.add(a, b, out=b), 6 * n))
