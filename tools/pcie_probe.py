#!/usr/bin/env python
"""What the box's PCIe link gives the e2e path: pinned H2D alone, D2H alone, and both directions at once (two streams)."""
import torch, time
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=8):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    return reps * n * (int(h2d) + int(d2h)) / dt / 1e9
for _ in range(2): run(True, True, 2)
print(f"H2D alone {run(True, False):.1f} GB/s, D2H alone {run(False, True):.1f} GB/s, both at once {run(True, True):.1f} GB/s (sum of the two directions)")

# the same copies while the SMs stream through HBM (a device-to-device copy loop on a third stream): what the e2e leg sees
big_a = torch.empty(4 << 30, dtype=torch.uint8, device="cuda"); big_b = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
s3 = torch.cuda.Stream()
def run_loaded(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    with torch.cuda.stream(s3):
        for _ in range(40):
            big_b.copy_(big_a, non_blocking=True)
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    s1.synchronize(); s2.synchronize(); dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return reps * n * (int(h2d) + int(d2h)) / dt / 1e9
print(f"under an HBM-streaming kernel: H2D {run_loaded(True, False):.1f} GB/s, D2H {run_loaded(False, True):.1f} GB/s, both {run_loaded(True, True):.1f} GB/s")
