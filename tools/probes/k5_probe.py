#!/usr/bin/env python3
"""K5 (size_blocks_kernel) on a batch worth profiling: 16384 text blocks of 64 KiB, exact sizing on, so that every
block is sized by the pre-pass before placement."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bo_lz4_ada_b200 as lz  # noqa: E402
from tools import corpus  # noqa: E402

c = corpus.build_corpus(1 << 30, 1 << 20, 4, kinds=("text",))
ctx = lz.DeviceContext(0)
src = bytes(c["src"])
b = lz.Batch(ctx, src, c["items"])
b.exact_sizing()
d_src, d_dst = ctx.alloc(len(src) + 64), ctx.alloc(b.output_bytes + 64)
b.upload(d_src)
b.run(d_src, d_dst)
res = b.results()
assert all(r["exception"] == "OK" for r in res) and sum(r["out_len"] for r in res) == c["plain_bytes"]
print("exact sizing ok: %d blocks, retried %d" % (b.block_count, b.retried_streams()))
b.close()
ctx.close()
