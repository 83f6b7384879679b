import torch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3
for mb in (32, 65, 131, 262, 524, 1048, 2097):
    n = mb * 1000 * 1000 // 2
    a = torch.randn(n, device=dev).to(torch.float16); b = torch.empty_like(a)
    t = timeit(lambda: b.copy_(a))
    t2 = timeit(lambda: b.zero_())
    t3 = timeit(lambda: a.sum())
    print(f"{mb:5d} MB tensor: copy {t:8.1f} us -> {2*mb/t:6.2f} TB/s (r+w) | memset {t2:8.1f} us -> {mb/t2:6.2f} TB/s | read-sum {t3:8.1f} us -> {mb/t3:6.2f} TB/s")
