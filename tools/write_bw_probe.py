"""Write-only HBM bandwidth probe (torch memset / fill of 8 GiB): the roof k_compact_lut is compared with."""
import torch, time
x = torch.empty(8 * 1024**3 // 8, dtype=torch.int64, device='cuda')
for fn, name in ((lambda: x.zero_(), 'memset 8GiB'), (lambda: x.fill_(7), 'fill 8GiB')):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(name, ms, 'ms', x.numel() * 8 / ms / 1e9, 'TB/s... GB/ms')
