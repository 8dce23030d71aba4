# write-only / read-only / copy HBM bandwidth with torch ops (context for the kernel_conv GEMM's 2 GB of writes per launch)
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GiB
y = torch.empty(1 << 30, dtype=torch.float32, device="cuda")
def t(f, n=5):
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        a.record(); f(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print(f"fill  4 GiB: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s written")
ms = t(lambda: y.copy_(x)); print(f"copy  4 GiB: {ms:.3f} ms  {2*x.numel()*4/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"sum   4 GiB: {ms:.3f} ms  {x.numel()*4/ms/1e6:.0f} GB/s read")
