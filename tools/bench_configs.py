#!/usr/bin/env python3
"""Secondary configurations of BASELINE.json (configs[2..4]) on one GPU: correctness against the
plain data plus device-resident throughput.  Not the headline bench (that is bench.py); the numbers
go to profiles/ as context for DESIGN.md."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bo_lz4_ada_b200 as lz  # noqa: E402
from tools import corpus  # noqa: E402


def run_case(ctx, name, streams, plains, steps=3):
    src = b"".join(streams)
    offs, pos = [], 0
    for s in streams:
        offs.append((pos, len(s)))
        pos += len(s)
    b = lz.Batch(ctx, src, offs)
    d_src = torch.empty(len(src) + 256, dtype=torch.uint8, device="cuda")
    d_dst = torch.empty(b.output_bytes + 256, dtype=torch.uint8, device="cuda")
    b.upload(d_src.data_ptr())
    b.run(d_src.data_ptr(), d_dst.data_ptr())
    res = b.results()
    ok = all(r["exception"] == "OK" for r in res)
    total = 0
    for r, p in zip(res, plains):
        got = bytes(d_dst[r["dst_off"]:r["dst_off"] + r["out_len"]].cpu().numpy())
        ok = ok and corpus.xxh32(got) == corpus.xxh32(p) and len(got) == len(p)
        total += r["out_len"]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kms = []
    for _ in range(steps):
        b.run(d_src.data_ptr(), d_dst.data_ptr())
        kms.append(b.kernel_ms())
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    tr = b.traffic()
    out = {"config": name, "streams": len(streams), "blocks": int(b.block_count), "plain_bytes": total,
           "compressed_bytes": len(src), "bit_exact": bool(ok), "ms_per_step": 1e3 * dt,
           "decompressed_GBps": total / dt / 1e9,
           "kernel_ms": {k: float(np.mean([m[k] for m in kms])) for k in kms[0]},
           "algorithmic_GBps": (tr["compressed_read"] + tr["decompressed_written"] + tr["checksum_reread"]) / dt / 1e9}
    b.close()
    print(json.dumps(out), flush=True)
    return out


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    only = sys.argv[2] if len(sys.argv) > 2 else ""   # substring filter on the config name
    torch.cuda.set_device(0)
    ctx = lz.DeviceContext(0, torch.cuda.current_stream().cuda_stream)
    results = []
    # configs[2]: 4 MiB independent blocks, thirds RLE / text / random, one-block frames (lz4 CLI defaults:
    # content checksum, no block checksum)
    for kinds in (("rle",), ("random",), ("text",), ("rle", "text", "random")):
        if only and only not in "4MiB-blocks/" + "+".join(kinds):
            continue
        n = int(1024 * scale) if kinds != ("text",) else int(256 * scale)
        c = corpus.build_corpus(n * (4 << 20), 4 << 20, 7, kinds=kinds, block_checksum=False, keep_plain=True)
        streams = [bytes(c["src"][o:o + l]) for o, l in c["items"]]
        results.append(run_case(ctx, "4MiB-blocks/" + "+".join(kinds), streams, c["plain"]))
        del c, streams
    # the headline corpus as ONE frame (64 KiB blocks, block + content checksum): the content
    # checksum is then a single serial XXH32 chain (K3 runs one quad), the decode is as parallel as before
    if only and only in "one-frame":
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(16) as ex:
            data = b"".join(ex.map(lambda i: corpus.text_like(64 << 20, seed=77 + i), range(max(1, int(16 * scale)))))
        for cc in (True, False):
            s = corpus.build_frame(data, 4, True, cc)
            results.append(run_case(ctx, "one-frame/64KiB-blocks/%s" % ("content-checksum" if cc else "no-content-checksum"), [s], [data],
                                    steps=2))
    # configs[4], second case: linked frames with 64 KiB blocks (every block leans on the whole previous one)
    if only and only in "linked-64KiB-blocks/256-frames":
        text64 = corpus.text_like(6 << 20, seed=5)
        streams, plains = [], []
        for i in range(int(256 * scale)):
            p = text64[(i * 10007) % (2 << 20):][:2 << 20]
            streams.append(corpus.build_frame(p, 4, True, True, True, independent=False))
            plains.append(p)
        results.append(run_case(ctx, "linked-64KiB-blocks/256-frames", streams, plains))
    if only:
        return
    # configs[3]: legacy frames, concatenated modern frames, skippable frames: batch of 1024 streams
    text = corpus.text_like(6 << 20, seed=5)
    rle = corpus.rle_like(2 << 20, seed=6)
    streams, plains = [], []
    for i in range(int(1024 * scale)):
        a = text[(i * 4099) % (5 << 20):][:200000 + (i % 7) * 30000]
        z = rle[(i * 7919) % (1 << 20):][:100000 + (i % 5) * 50000]
        kind = i % 4
        if kind == 0:
            s = corpus.build_legacy_frame(a + z)
            p = a + z
        elif kind == 1:
            s = corpus.build_frame(a, 4, True, True) + corpus.build_frame(z, 4, False, True, True)
            p = a + z
        elif kind == 2:
            s = corpus.skippable_frame(b"meta" * (i % 9), i % 16) + corpus.build_frame(a, 4) + corpus.skippable_frame(b"", 1)
            p = a
        else:
            s = corpus.build_legacy_frame(z) + corpus.build_frame(a, 5, True, True)
            p = z + a
        streams.append(s)
        plains.append(p)
    results.append(run_case(ctx, "legacy+concatenated+skippable/1024-streams", streams, plains))
    # configs[4]: linked-block frames, 256 concurrent frames (256 KiB blocks), K4 one warp per frame
    streams, plains = [], []
    for i in range(int(256 * scale)):
        p = text[(i * 10007) % (2 << 20):][:2 << 20]
        streams.append(corpus.build_frame(p, 5, True, True, True, independent=False))
        plains.append(p)
    results.append(run_case(ctx, "linked-256KiB-blocks/256-frames", streams, plains))
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                           "bench_configs.json"), "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
