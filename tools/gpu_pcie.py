"""Development aid (GPU box): what the host link gives -- pinned H2D / D2H bandwidth with torch, one and two directions at once."""
import time
import torch

dev = torch.device("cuda", 0)
for mb in (16, 64, 256):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    up = 5 * n / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    for _ in range(5):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    down = 5 * n / (time.perf_counter() - t0) / 1e9
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    both = 10 * n / (time.perf_counter() - t0) / 1e9
    print(f"{mb:4d} MB  H2D {up:6.1f} GB/s  D2H {down:6.1f} GB/s  both directions {both:6.1f} GB/s", flush=True)
