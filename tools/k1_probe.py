#!/usr/bin/env python3
"""K1 probe: decode one synthetic corpus with several K1 tunings, print kernel times.

    python tools/k1_probe.py --size-mib 1024 --tunings 64,16 [--kinds text] [--block 64k]

With LZ4B200_PROF=1 the v3 kernel's phase counters are printed when the context closes.
Development aid (device-resident timing only); bench.py is the measurement of record."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size-mib", type=int, default=1024)
    ap.add_argument("--frame-mib", type=float, default=1.0)
    ap.add_argument("--block", default="64k", choices=["64k", "256k", "1m", "4m"])
    ap.add_argument("--kinds", default="text")
    ap.add_argument("--tunings", default="64,16")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-block-checksum", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--linked", action="store_true", help="block-dependent frames (chains: K4 / K6)")
    args = ap.parse_args()
    import torch
    import bo_lz4_ada_b200 as lz
    from tools import corpus

    code = {"64k": 4, "256k": 5, "1m": 6, "4m": 7}[args.block]
    c = corpus.build_corpus(args.size_mib << 20, int(args.frame_mib * (1 << 20)), code, kinds=tuple(args.kinds.split(",")),
                            block_checksum=not args.no_block_checksum, seed=args.seed, independent=not args.linked)
    src_np = np.frombuffer(c["src"], dtype=np.uint8)
    n_src = len(src_np)
    stream = torch.cuda.current_stream()
    ctx = lz.DeviceContext(0, stream.cuda_stream)
    h_src = torch.empty(n_src + 64, dtype=torch.uint8).pin_memory()
    h_src[:n_src].copy_(torch.from_numpy(src_np.copy()))
    batch = lz.Batch(ctx, h_src.data_ptr(), c["items"])
    batch.src_bytes = n_src
    d_src = torch.empty(n_src + 256, dtype=torch.uint8, device="cuda")
    d_dst = torch.empty(batch.output_bytes + 256, dtype=torch.uint8, device="cuda")
    batch.upload(d_src.data_ptr())
    torch.cuda.synchronize()
    for t in [int(x) for x in args.tunings.split(",")]:
        ctx.set_tuning(t)
        ms = []
        for _ in range(args.reps + 1):
            batch.run(d_src.data_ptr(), d_dst.data_ptr())
            ms.append(batch.kernel_ms())
        res = batch.results()
        bad = [r for r in res if r["exception"] != "OK"]
        plain = sum(r["out_len"] for r in res)
        k1 = float(np.mean([m["k1_decode_blocks"] for m in ms[1:]]))
        k4 = float(np.mean([m.get("k4_decode_linked", 0.0) for m in ms[1:]]))
        k3 = float(np.mean([m["k3_xxh32_frames"] for m in ms[1:]]))
        if bad:
            print("bad streams: %d of %d; first: %s" % (len(bad), len(res), [(i, r["exception"], r["message"][:100]) for i, r in enumerate(res) if r["exception"] != "OK"][:3]), file=sys.stderr)
        print(json.dumps({"tuning": t, "fallbacks": ctx.k1_fallbacks() if t in (60, 61) else None, "kernel": batch.k1_kernel_name(), "blocks": int(batch.block_count), "plain_bytes": plain, "ok": not bad and plain == c["plain_bytes"],
                          "k1_ms": k1, "k4_ms": k4, "k3_ms": k3, "k1_GBps_out": plain / (k1 / 1e3) / 1e9 if k1 else 0,
                          "k1_alg_GBps": (plain + n_src) / (k1 / 1e3) / 1e9 if k1 else 0}), flush=True)
    ctx.set_tuning(0)
    batch.close()
    ctx.close()


if __name__ == "__main__":
    main()
