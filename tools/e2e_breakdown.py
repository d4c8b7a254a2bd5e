"""Where the end-to-end time goes: raw PCIe copy rates, host planning, pipelined device stage."""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bo_lz4_ada_b200 as lz
from tools import corpus

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
torch.cuda.set_device(0)
n = int(gib * (1 << 30))
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("%s pinned %.2f GB/s" % (name, n / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("H2D+D2H concurrent: %.2f GB/s each direction" % (n / dt / 1e9))
del h2, d2, d, h

c = corpus.build_corpus(n, 1 << 20, 4)
src = np.frombuffer(c["src"], dtype=np.uint8)
hs = torch.empty(len(src) + 64, dtype=torch.uint8).pin_memory(); hs[:len(src)].copy_(torch.from_numpy(src.copy()))
ctx = lz.DeviceContext(0)
t = time.perf_counter(); b = lz.Batch(ctx, hs.data_ptr(), c["items"]); print("plan %.1f ms" % (1e3 * (time.perf_counter() - t)))
need = b.output_bytes
hd = torch.empty(need + 64, dtype=torch.uint8).pin_memory()
ds = torch.empty(len(src) + 256, dtype=torch.uint8, device="cuda"); dd = torch.empty(need + 256, dtype=torch.uint8, device="cuda")
t = time.perf_counter(); assert lz.lib().lz4ada_batch_upload(b._h, None, None) == 0; print("tables %.1f ms" % (1e3 * (time.perf_counter() - t)))
for chunks in (1, 4, 8, 16):
    for rep in range(2):
        t = time.perf_counter()
        assert lz.lib().lz4ada_batch_run_pipelined(b._h, hs.data_ptr(), hd.data_ptr(), ds.data_ptr(), dd.data_ptr(), chunks) == 0
        dt = time.perf_counter() - t
    print("pipelined chunks=%d: %.1f ms -> %.1f GB/s" % (chunks, 1e3 * dt, c["plain_bytes"] / dt / 1e9))
