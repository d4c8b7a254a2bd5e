import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import bo_lz4_ada_b200 as lz, oracle_binding
from tools import corpus
from test_gpu_parity import _mutations
o = oracle_binding.load()
ctx = lz.DeviceContext(0)
g = -1
rng = np.random.default_rng(100 + g)
text = corpus.text_like(60000, seed=19)
mix = text[:20000] + bytes(5000) + corpus.random_bytes(3000, seed=1) + text[20000:45000]
bases = [corpus.build_frame(mix, 4, False, True, True), corpus.build_frame(mix, 4, True, True, False),
         corpus.build_frame(text, 4, False, False, False, block_size=7000)]
bad = []
for base in bases:
    bad += _mutations(base, rng, 40)
data = bad[117]
base = bases[2]
diff = [i for i in range(min(len(data), len(base))) if data[i] != base[i]]
print("stream len", len(data), len(base), "mutated bytes at", diff[:10], [hex(data[i]) for i in diff[:10]], [hex(base[i]) for i in diff[:10]])
oexc, oout, oeof, omsg = o.decode_stream(data, chunk=0, out_cap=1 << 21)
for trial in range(3):
    exc, out, eof, msg = lz.batch_decompress(ctx, [data])[0]
    dpos = [i for i in range(min(len(out), len(oout))) if out[i] != oout[i]]
    print(trial, exc == oexc, msg == omsg, len(out), len(oout), "diff positions", dpos[:8], [(chr(out[i]), chr(oout[i])) for i in dpos[:8]])
print(omsg)
# where do blocks start
import struct
pos = 7; blk = 0; outpos = 0
while pos < len(data):
    w = struct.unpack_from("<I", data, pos)[0]; pos += 4
    if w == 0: break
    n = w & 0x7ffffff
    print("block", blk, "src", pos, "len", n, "stored", bool(w >> 31))
    pos += n; blk += 1
