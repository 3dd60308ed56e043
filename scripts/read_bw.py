"""Pure-read HBM bandwidth probe (torch reduction kernels) for context next to the copy peak."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device='cuda').normal_()   # 4 GiB
for fn, name in ((lambda: x.sum(), 'sum f32'), (lambda: x.max(), 'max f32'),
                 (lambda: x.view(torch.int32).bitwise_and(1).sum() if False else x.amax(), 'amax')):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(name, f'{x.numel()*4/ms/1e6:.0f} GB/s')
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print('copy', f'{2*x.numel()*4/(e0.elapsed_time(e1)/10)/1e6:.0f} GB/s')
