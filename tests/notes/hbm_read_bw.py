import torch
x = torch.randn(1 << 30, device="cuda", dtype=torch.bfloat16)
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.sum()); print(f"sum bf16 2 GiB: {ms:.3f} ms -> {2.147/ms:.2f} TB/s read-only")
xf = x.view(torch.float32)
ms = t(lambda: xf.sum()); print(f"sum f32 view 2 GiB: {ms:.3f} ms -> {2.147/ms:.2f} TB/s read-only")
ms = t(lambda: y.copy_(x)); print(f"copy 2+2 GiB: {ms:.3f} ms -> {4.295/ms:.2f} TB/s r+w")
ms = t(lambda: y.zero_()); print(f"memset 2 GiB: {ms:.3f} ms -> {2.147/ms:.2f} TB/s write-only")
ms = t(lambda: torch.add(x, x, out=y)); print(f"add x+x->y (1r,1w): {ms:.3f} ms -> {4.295/ms:.2f} TB/s")
z = torch.randn(1 << 30, device="cuda", dtype=torch.bfloat16)
ms = t(lambda: torch.add(x, z, out=y)); print(f"add x+z->y (2r,1w) 6 GiB: {ms:.3f} ms -> {6.442/ms:.2f} TB/s")
ms = t(lambda: torch.dot(x, z)); print(f"dot (2r) 4 GiB: {ms:.3f} ms -> {4.295/ms:.2f} TB/s read-only")
