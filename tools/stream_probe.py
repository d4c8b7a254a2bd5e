#!/usr/bin/env python3
"""Throughput of the streaming API (LZ4Ada.Update, one block per call) on one frame.

    python tools/stream_probe.py [MiB] [block-code 4..7] [linked 0/1]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bo_lz4_ada_b200 as lz  # noqa: E402
from tools import corpus  # noqa: E402


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    code = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    linked = len(sys.argv) > 3 and sys.argv[3] == "1"
    plain = corpus.text_like(mib << 20, seed=9)
    frame = corpus.build_frame(plain, code, True, True, False, not linked)
    ctx = lz.DeviceContext(0)
    ctx.make_default()
    import numpy as np
    arr = np.frombuffer(bytearray(frame), dtype=np.uint8)   # writable: slices reach the library without a copy
    for feed in (1 << 16, 1 << 20, 8 << 20, len(frame)):
        best = None
        for rep in range(2):
            dec = lz.Init()
            total = 0
            pos = 0
            t0 = time.perf_counter()
            first = b""
            while pos < len(frame):
                c, o, _, _ = dec.Update(arr[pos:pos + feed])
                if not first:
                    first = o
                total += len(o)
                pos += c
            dt = time.perf_counter() - t0
            assert total == len(plain) and first == plain[:len(first)]
            best = dt if best is None or dt < best else best
            dec.close()
        print("Update loop: %d MiB, block code %d, %s, Input pieces of %d KiB: %.1f ms -> %.1f MB/s decompressed" % (
            mib, code, "linked" if linked else "independent", feed >> 10, 1e3 * best, len(plain) / best / 1e6), flush=True)
    lz.lib().lz4ada_set_device_context(None)
    ctx.close()


if __name__ == "__main__":
    main()
