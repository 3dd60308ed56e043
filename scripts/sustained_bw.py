"""HBM copy / read bandwidth in a long loop (power-cap regime) vs a short burst."""
import subprocess, time, torch
x = torch.empty(1 << 30, dtype=torch.float32, device='cuda').normal_()
y = torch.empty_like(x)
def run(fn, bytes_per_call, seconds):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t_end = time.perf_counter() + seconds
    best = 0; last = 0
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        last = bytes_per_call * 20 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        best = max(best, last)
    clk = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw', '--format=csv,noheader'],
                         capture_output=True, text=True).stdout.strip()
    return best, last, clk
for name, fn, nbytes in (('copy', lambda: y.copy_(x), 2 * x.numel() * 4), ('sum', lambda: x.sum(), x.numel() * 4)):
    b, l, clk = run(fn, nbytes, 0.3)
    print(name, 'burst', round(b), 'GB/s', clk)
    b, l, clk = run(fn, nbytes, 5.0)
    print(name, 'sustained(last)', round(l), 'best', round(b), 'GB/s', clk)
