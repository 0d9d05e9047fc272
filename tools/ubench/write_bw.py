import torch, time
dev = torch.device("cuda")
n = 64*1080*1920*3
t = torch.empty(n, dtype=torch.uint8, device=dev)
src = torch.empty(n, dtype=torch.uint8, device=dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
ms = timeit(lambda: t.zero_()); print("memset 398MB: %.3f ms  %.0f GB/s" % (ms, n/ms/1e6))
ms = timeit(lambda: t.fill_(7)); print("fill 398MB: %.3f ms  %.0f GB/s" % (ms, n/ms/1e6))
ms = timeit(lambda: t.copy_(src)); print("copy 398MB: %.3f ms  %.0f GB/s (r+w)" % (ms, 2*n/ms/1e6))
big = torch.empty(4*n, dtype=torch.uint8, device=dev)
ms = timeit(lambda: big.zero_()); print("memset 1.6GB: %.3f ms  %.0f GB/s" % (ms, 4*n/ms/1e6))
