"""Pinned host -> device bandwidth on this box: one stream vs two / four concurrent streams (what bounds the e2e number)."""
import time
import torch

N = 1600 * (1 << 20)
h = torch.empty(N, dtype=torch.uint8).pin_memory()
d = torch.empty(N, dtype=torch.uint8, device="cuda")
for k in (1, 2, 4):
    ss = [torch.cuda.Stream() for _ in range(k)]
    per = N // k
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(ss):
            with torch.cuda.stream(s):
                d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{k} stream(s): {N / dt / 1e9:.1f} GB/s")
# chunked, as svb_scan_batch_v1_host does it: 8 x 200 MB alternating between two streams
ss = [torch.cuda.Stream() for _ in range(2)]
per = N // 8
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8):
    with torch.cuda.stream(ss[i & 1]):
        d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
torch.cuda.synchronize()
print(f"8 chunks on 2 streams: {N / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
