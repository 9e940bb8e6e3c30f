"""Developer diagnostic (GPU box): pinned host<->device copy bandwidth at the bench's per-step sizes."""
import time
import torch
dev = torch.device("cuda", 0)
for mb in (2, 21, 25, 256):
    n = mb * (1 << 20)
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=20):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    run(True, True, 3)
    a, b, c = run(True, False), run(False, True), run(True, True)
    print("%4d MB: H2D %.1f GB/s  D2H %.1f GB/s  both %.1f + %.1f GB/s (%.3f ms)" % (mb, n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9, c * 1e3))
