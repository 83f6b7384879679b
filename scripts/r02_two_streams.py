"""Does splitting the batch over two CUDA streams (two independent sets of nets, half the clips each) beat one stream?
The tails / ramps of one stream's launches can overlap the steady state of the other's (as Detector || Locator already do)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
dev = torch.device("cuda:0")
A = bench.make_models(dev); Bm = bench.make_models(dev)
x_np, msg_np, _ = bench.synth(64, 16000, 100)
x = torch.from_numpy(x_np).to(dev); msg = torch.from_numpy(msg_np).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, n=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return min(ts), sum(ts) / len(ts)
def g_one():
    A["generator"].embed_batch(x, msg, want_wm=False)
def g_two(parts=2):
    main = torch.cuda.current_stream()
    n = 64 // parts
    for i, (st, M) in enumerate(((s1, A), (s2, Bm))):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            M["generator"].embed_batch(x[i * 32:(i + 1) * 32], msg[i * 32:(i + 1) * 32], want_wm=False)
    main.wait_stream(s1); main.wait_stream(s2)
def full_one():
    _, y, _ = A["generator"].embed_batch(x, msg, want_wm=False)
    main = torch.cuda.current_stream(); s1.wait_stream(main)
    with torch.cuda.stream(s1):
        A["locator"].locate_batch(y)
    A["detector"].detect_batch(y)
    main.wait_stream(s1)
def full_two():
    main = torch.cuda.current_stream()
    for i, (st, M) in enumerate(((s1, A), (s2, Bm))):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            _, y, _ = M["generator"].embed_batch(x[i * 32:(i + 1) * 32], msg[i * 32:(i + 1) * 32], want_wm=False)
            M["detector"].detect_batch(y); M["locator"].locate_batch(y)
    main.wait_stream(s1); main.wait_stream(s2)
for name, fn in (("G one stream", g_one), ("G two streams x 32 clips", g_two), ("G+D+L one stream (L on side)", full_one), ("G+D+L two streams x 32 clips", full_two)):
    mn, av = timeit(fn)
    print(f"{name}: min {mn*1e3:.0f} us avg {av*1e3:.0f} us", flush=True)
