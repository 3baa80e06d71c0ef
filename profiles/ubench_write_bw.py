import torch, time
x = torch.empty(1<<30, dtype=torch.float64, device='cuda')   # 8 GB
y = torch.empty(1<<28, dtype=torch.float64, device='cuda')   # 2 GB
for t, name in ((x,'8GB'),):
    for _ in range(3): t.fill_(1.0)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): t.fill_(2.0)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print('fill', name, ms, 'ms', t.numel()*8/ms/1e6, 'GB/s')
# read-only: sum
for _ in range(2): x.sum()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.sum()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print('sum 8GB', ms, 'ms', x.numel()*8/ms/1e6, 'GB/s')
# 1:7 read:write mix: y read, x written (copy with broadcast)
z = x.view(4, 1<<28)
for _ in range(2): z.copy_(y.unsqueeze(0).expand(4, -1))
torch.cuda.synchronize()
e0.record()
for _ in range(5): z.copy_(y.unsqueeze(0).expand(4, -1))
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/5
print('broadcast copy 2GB->8GB', ms, 'ms', (x.numel()+y.numel())*8/ms/1e6, 'GB/s total')
