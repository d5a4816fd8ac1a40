"""Copy bandwidth of THIS box (same method as MEASURED_PEAKS.json): b.copy_(a), 1 Gi bf16, best of 10."""
import torch
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda")
b = torch.empty_like(a)
best = 0.0
for _ in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print("copy_GBps %.1f" % best)
