import torch, sys
dev = torch.device("cuda:0")
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
a = torch.empty(65536, 2048, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a); c = torch.empty_like(a)
ms = t(lambda: (a.zero_(), b.zero_())); print(f"memset 537 MB: {ms*1e3:.1f} us  {0.537/ms:.2f} TB/s write")
ms = t(lambda: b.copy_(a)); print(f"copy 268 -> 268 MB: {ms*1e3:.1f} us  {0.537/ms:.2f} TB/s r+w")
x = torch.empty(65536, 512, dtype=torch.bfloat16, device=dev)
ms = t(lambda: (a.fill_(1.0), c.fill_(2.0), x.sum())); print(f"fill 537 MB + read 67 MB: {ms*1e3:.1f} us")
