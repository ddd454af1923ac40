import torch
x = torch.empty(2*1024**3, dtype=torch.int32, device="cuda")   # 8 GiB
for name, fn in (("zero_", lambda: x.zero_()), ("fill_(7)", lambda: x.fill_(7))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"{name}: {ms:.3f} ms for 8 GiB = {x.numel()*4/ms/1e6:.0f} GB/s")
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print(f"copy 8 GiB -> 8 GiB: {ms:.3f} ms = {2*x.numel()*4/ms/1e6:.0f} GB/s (read+write)")
