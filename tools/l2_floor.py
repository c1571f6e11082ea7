import torch, time
def bench(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for mib in (16, 256):
    n = mib * (1 << 20) // 16
    x = torch.randn(n, 2, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x); z = torch.randn_like(x)
    print(mib, "MiB copy us", round(bench(lambda: y.copy_(x)), 2), " y+=x us", round(bench(lambda: y.add_(x)), 2),
          " z=x+y", round(bench(lambda: torch.add(x, y, out=z)), 2))
