"""Pinned-memory copy bandwidth of the box: H2D alone, D2H alone, both at once (two streams)."""
import torch
n = 64 * 1024 * 1024 // 4
h1, h2 = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
d1, d2 = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=20):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record()
    torch.cuda.synchronize()
    return (up + down) * reps * n * 4 / (a.elapsed_time(b) * 1e-3) / 1e9


run(True, True, 3)
print("H2D %.1f GB/s  D2H %.1f GB/s  both %.1f GB/s (sum of directions)" % (run(True, False), run(False, True), run(True, True)))
