import torch
N = 134_400_000 // 4
x = torch.empty(N, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
def run(n, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        e0.record(s)
        for _ in range(reps):
            off = 0
            for c in x.chunk(n):
                d[off:off + c.numel()].copy_(c, non_blocking=True)
                off += c.numel()
        e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
run(1, 3)
for n in (1, 2, 4, 8, 16, 64, 1, 4):
    ms = run(n)
    print(f"chunks={n:3d} ({134.4 / n:6.1f} MB each): {ms:.3f} ms -> {0.1344 / ms * 1e3:.1f} GB/s")
