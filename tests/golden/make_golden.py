#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference's own test vectors.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

What it does
  * copies every .lz4 / .err / .eds of /root/reference/test_vectors_lz4 verbatim
    (these are the reference's test *data*, not source code);
  * copies .bin files up to 512 KiB; larger ones are represented by their
    size + sha256 + XXH32 in manifest.json (tests compare digests);
  * z9m.bin is absent from the reference checkout (.MISSING_LARGE_BLOBS); it is
    9 437 166 zero bytes and its XXH32 must equal the content
    checksum stored in z9m.lz4 -- checked here;
  * records the inline known-answer cases of test_suite/lz4test.adb (XXH32 KAT
    :131-140, two legacy frames :153-170, raw block :217-221, reservation
    prefix :355-362, Single_Frame violation :387-407) in inline_cases.json.

The GPU box has no /root/reference; nothing under tests/ reads it at run time.
"""
import hashlib
import json
import os
import shutil
import sys

import xxhash

SRC = "/root/reference/test_vectors_lz4"
DST = os.path.dirname(os.path.abspath(__file__))
BIN_COPY_LIMIT = 512 * 1024


def digest(b):
    return {"size": len(b), "sha256": hashlib.sha256(b).hexdigest(),
            "xxh32": "%08x" % xxhash.xxh32(b, seed=0).intdigest()}


def main():
    if not os.path.isdir(SRC):
        sys.exit("reference vectors not found at " + SRC)
    manifest = {"good": {}, "error": {}}
    for name in sorted(os.listdir(SRC)):
        stem, ext = os.path.splitext(name)
        path = os.path.join(SRC, name)
        if ext in (".lz4", ".err", ".eds"):
            shutil.copyfile(path, os.path.join(DST, name))
            os.chmod(os.path.join(DST, name), 0o644)
    for name in sorted(os.listdir(SRC)):
        stem, ext = os.path.splitext(name)
        if ext == ".lz4":
            binp = os.path.join(SRC, stem + ".bin")
            if os.path.exists(binp):
                data = open(binp, "rb").read()
            elif stem == "z9m":
                data = bytes(9437166)
                stored = int.from_bytes(open(os.path.join(SRC, name), "rb").read()[-4:], "little")
                assert stored == xxhash.xxh32(data, seed=0).intdigest(), "z9m content checksum"
            else:
                raise SystemExit("no .bin for " + name)
            entry = digest(data)
            entry["lz4"] = digest(open(os.path.join(SRC, name), "rb").read())
            entry["bin_in_tree"] = len(data) <= BIN_COPY_LIMIT
            if entry["bin_in_tree"]:
                with open(os.path.join(DST, stem + ".bin"), "wb") as f:
                    f.write(data)
            manifest["good"][stem] = entry
        elif ext == ".err":
            eds = open(os.path.join(SRC, stem + ".eds"), "r").readline().rstrip("\n")
            manifest["error"][stem] = {"eds": eds,
                                       "err": digest(open(os.path.join(SRC, name), "rb").read())}
    with open(os.path.join(DST, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)

    two_legacy = bytes([
        0x02, 0x21, 0x4c, 0x18, 0x30, 0x00, 0x00, 0x00, 0xf0, 0x1f, 0x3c, 0x3f, 0x78, 0x6d,
        0x6c, 0x20, 0x76, 0x65, 0x72, 0x73, 0x69, 0x6f, 0x6e, 0x3d, 0x22, 0x31, 0x2e, 0x30,
        0x22, 0x20, 0x65, 0x6e, 0x63, 0x6f, 0x64, 0x69, 0x6e, 0x67, 0x3d, 0x22, 0x55, 0x54,
        0x46, 0x2d, 0x38, 0x22, 0x3f, 0x3e, 0x3c, 0x74, 0x65, 0x73, 0x74, 0x2f, 0x3e, 0x0a,
        0x02, 0x21, 0x4c, 0x18, 0x0e, 0x00, 0x00, 0x00, 0xd0, 0x48, 0x65, 0x6c, 0x6c, 0x6f,
        0x20, 0x77, 0x6f, 0x72, 0x6c, 0x64, 0x2e, 0x0a])
    minilegacy = open(os.path.join(SRC, "minilegacy.lz4"), "rb").read()
    inline = {
        "xxh32_kat": {"input_hex": (bytes([0x1a] * 14) + bytes([0x11, 0x10])).hex(),
                      "expect": "f994ef8a", "cite": "test_suite/lz4test.adb:129-147"},
        "two_legacy_frames": {"input_hex": two_legacy.hex(),
                              "expect_hex": (b'<?xml version="1.0" encoding="UTF-8"?><test/>\nHello world.\n').hex(),
                              "cite": "test_suite/lz4test.adb:149-214"},
        "hello_block": {"input_hex": bytes([0xd0, 0x48, 0x65, 0x6c, 0x6c, 0x6f, 0x2c, 0x20, 0x77,
                                            0x6f, 0x72, 0x6c, 0x64, 0x2e]).hex(),
                        "expect_hex": b"Hello, world.".hex(), "cite": "test_suite/lz4test.adb:216-248"},
        "reservation_exceeded": {"input_hex": open(os.path.join(SRC, "z2841.lz4"), "rb").read()[:36].hex(),
                                 "reservation": "SZ_64_KiB", "expect": "TOO_LITTLE_MEMORY",
                                 "cite": "test_suite/lz4test.adb:353-382"},
        "unexpected_multi_frame": {"input_hex": (minilegacy + minilegacy).hex(),
                                   "expect": "DATA_CORRUPTION", "cite": "test_suite/lz4test.adb:384-430"},
    }
    assert len(minilegacy + minilegacy) == 112
    with open(os.path.join(DST, "inline_cases.json"), "w") as f:
        json.dump(inline, f, indent=1, sort_keys=True)
    print("golden: %d good, %d error vectors" % (len(manifest["good"]), len(manifest["error"])))


if __name__ == "__main__":
    main()
